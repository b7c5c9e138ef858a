#!/usr/bin/env python
"""Kernel A/B micro-bench on the bench index (development aid; bench.py is the measurement of record).
    python tools/kbench.py --index BASE [--reads PREFIX] [--policy n2k1] [--B 2000000] [--steps 5] [--streams 4] [--tag name]
BOWTIE_B200_LIB selects the library variant; BT_RARE_PERIOD / BT_RARE_THRESH / BT_MAIN_BUDGET are read by the library.
Reads are generated once and cached as PREFIX.<policy>.<B>.npz so that every variant sees the same batches."""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--index", required=True); ap.add_argument("--reads", default="/dev/shm/kbench_reads")
ap.add_argument("--policy", default="n2k1"); ap.add_argument("--B", type=int, default=2_000_000)
ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--streams", type=int, default=4); ap.add_argument("--tag", default="")
ap.add_argument("--single", action="store_true", help="also time one synchronous batch (latency)")
a = ap.parse_args()
import torch, bowtie_b200  # noqa: E402
base = Path(a.index)
pd = bench.POLICIES[a.policy]; R = 2 if pd["paired"] else 1
cache = Path(f"{a.reads}.{a.policy}.{a.B}.npz")
if cache.exists():
    z = np.load(cache); hs = [(z[f"c{k}"], z[f"q{k}"], z[f"o{k}"], z[f"s{k}"]) for k in range(2)]
else:
    g = bench.load_genome(base)
    gen = bench.make_pairs if R == 2 else bench.make_reads
    hs = [gen(g, a.B, seed=999 + k)[:4] for k in range(2)]
    np.savez(cache, **{f"{n}{k}": hs[k][i] for k in range(2) for i, n in enumerate("cqos")})
    del g
t0 = time.time()
ix = bowtie_b200.Index(str(base), need_mirror=True, device=0)
t_load = time.time() - t0
pol = bench.lib_policy(a.policy)
slots, mm_cap = R, 7; rw = bowtie_b200.BT_HIT_HDR_WORDS + mm_cap
dev = [tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (h[0], h[1], h[2].view(np.int64), h[3].view(np.int32))) for h in hs]
NS = a.streams
ctxs = [bowtie_b200.Context(ix) for _ in range(NS)]; streams = [torch.cuda.Stream() for _ in range(NS)]
outs = [(torch.zeros(a.B, dtype=torch.int32, device="cuda"), torch.zeros(a.B, dtype=torch.int32, device="cuda"), torch.zeros(a.B * slots * rw, dtype=torch.int32, device="cuda")) for _ in range(NS)]
main = torch.cuda.current_stream()
def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in streams: st.wait_event(e0)
    for k in range(n):
        s, q, o, sd = dev[k & 1]; f, g, h = outs[k % NS]
        ctxs[k % NS].align_device(s.data_ptr(), q.data_ptr(), o.data_ptr(), sd.data_ptr(), a.B * R, bench.READ_LEN, pol, f.data_ptr(), g.data_ptr(), h.data_ptr(), slots, mm_cap, streams[k % NS].cuda_stream)
    for cx, st in zip(ctxs, streams):
        cx.join(st.cuda_stream); ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
run(a.warmup); ix.stats(reset=True)
ms = run(a.steps); st = ix.stats(reset=True)
res = {"tag": a.tag, "lib": os.environ.get("BOWTIE_B200_LIB", "default"), "policy": a.policy, "units_per_s": a.B * a.steps / (ms / 1e3), "ms_per_step": ms / a.steps,
       "env": {k: os.environ[k] for k in os.environ if k.startswith("BT_") and k not in ("BT_BUILD_VERBOSE",)},
       "side_fetches_per_unit": st.side_fetches / (a.B * a.steps), "alg_GBs": st.algorithmic_bytes / (ms / 1e3) / 1e9, "iters_per_unit": st.iters / (a.B * a.steps),
       "aligned_frac": float((outs[(a.steps - 1) % NS][0] > 0).float().mean().item()), "index_load_s": round(t_load, 1)}
# order-independent checksum of the results of the last step (variants must agree)
f, g, h = outs[(a.steps - 1) % NS]
res["checksum"] = int((h.view(a.B, -1)[:, :5].to(torch.int64) * torch.arange(1, 6, device="cuda")).sum().item() % (1 << 61))
if a.single:
    torch.cuda.synchronize(); t0 = time.perf_counter(); ms1 = run(1); res["single_batch_ms"] = ms1
print(json.dumps(res))
