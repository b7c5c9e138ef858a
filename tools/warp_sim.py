"""Warp-level replay of bt_search_kernel's scheduling on the host emulation (tests/host_emu/emu.cpp: emu_warp_sim): how full a warp is
when it executes each kind of code, for a given deferral period / threshold / budget — the quantity ncu calls "threads active per
instruction", estimated without a GPU.  Development aid for choosing what to measure; not a measurement.
Usage: python tools/warp_sim.py [n_reads=20000] [period=16] [thresh=24] [budget=8000] [nwarps=16]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from helpers import HostEmu, Policy, dev_policy  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
period = int(sys.argv[2]) if len(sys.argv) > 2 else 16
thresh = int(sys.argv[3]) if len(sys.argv) > 3 else 24
budget = int(sys.argv[4]) if len(sys.argv) > 4 else 8000
nwarps = int(sys.argv[5]) if len(sys.argv) > 5 else 16
base = bench.Path(os.environ.get("BT_BENCH_INDEX", str(bench.REF_DIR / "cache" / "benchs_24_256000000_1_10_5_1")))
genome = bench.load_genome(base)
codes, quals, offs, seeds, _ = bench.make_reads(genome, n, seed=12345)
emu = HostEmu()
L = emu.L
L.emu_warp_sim.restype = C.c_int
pol = Policy(mode=1, mms=2, khits=1)
cp = dev_policy(pol)
out = np.zeros(9 + 33, np.float64)
fw, bw = emu.index(base, False), emu.index(base, True)
rc = L.emu_warp_sim(C.c_void_p(fw), C.c_void_p(bw), C.byref(cp), C.c_uint32(n), C.c_void_p(codes.ctypes.data), C.c_void_p(quals.ctypes.data),
                    C.c_void_p(offs.ctypes.data), C.c_void_p(seeds.ctypes.data), C.c_uint32(nwarps), C.c_uint32(period), C.c_uint32(thresh), C.c_uint32(budget),
                    C.c_uint32(800), C.c_uint32(16), C.c_uint32(256), C.c_void_p(out.ctypes.data))
assert rc == 0
it, fpaths, flanes, rpass, rpaths, rlanes, idle, live, budgeted = out[:9]
print(f"period {period} thresh {thresh} budget {budget}: {it / n:.1f} warp iterations per read x 32 lanes")
print(f"  lanes holding a read per iteration      {live / it:5.1f} of 32")
print(f"  fast lanes per iteration                {flanes / it:5.1f}   (code paths per fast pass: {fpaths / max(1, (out[9 + 1:9 + 33].sum())):.2f} -> {flanes / max(fpaths, 1):.1f} lanes per executed path)")
print(f"  rare pass in {100 * rpass / it:4.1f} % of the iterations: {rlanes / max(rpass, 1):.1f} lanes, {rpaths / max(rpass, 1):.1f} distinct states -> {rlanes / max(rpaths, 1):.1f} lanes per executed path")
print(f"  lanes waiting for a rare pass           {idle / it:5.1f} per iteration")
print(f"  reads moved to the heavy pass           {int(budgeted)} of {n}")
h = out[9:]
cf, cr = float(os.environ.get("BT_SIM_CF", 110)), float(os.environ.get("BT_SIM_CR", 600))
print(f"  cost model ({cf:.0f} instr per fast path, {cr:.0f} per rare path; calibrated on the measured period/threshold sweep of profiles/README.md): "
      f"{(fpaths * cf + rpaths * cr) / n:.0f} warp instructions per read")
print("  fast-lane histogram (iterations with k fast lanes), k = 0..32 in eighths:", [int(h[i:i + 4].sum()) for i in range(1, 33, 4)])
