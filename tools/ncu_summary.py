#!/usr/bin/env python
"""Summarise an Nsight Compute report of one kernel launch for profiles/: the headline metrics (JSON) and, when the library build that was
profiled is given, where the executed instructions and the stall samples fall by source region (SASS lines of the report joined with
nvdisasm's line table of the same cubin).
    python tools/ncu_summary.py REPORT.ncu-rep OUT_PREFIX [LIB.so KERNEL_SYMBOL REGIONS.py]
REGIONS.py: a Python literal list of (name, file, first_line, last_line)."""
import collections, csv, json, re, subprocess, sys, tempfile
rep, outp = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
m = {}
for i, h in enumerate(hdr):
    if h in keep:
        m[h] = vals[i] + (" " + units[i] if units[i] else "")
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v >= 0.1:
            m.setdefault("stall_cycles_per_issued_instruction", {})[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 2)
json.dump(m, open(outp + ".metrics.json", "w"), indent=1)
print(json.dumps(m, indent=1))
if len(sys.argv) > 5:
    import os
    so, sym, regions = os.path.abspath(sys.argv[3]), sys.argv[4], eval(open(sys.argv[5]).read())
    td = tempfile.mkdtemp()
    subprocess.run(f"cd {td} && cuobjdump -xelf all {so} >/dev/null && for f in *.cubin; do nvdisasm -g -c $f > $f.txt 2>/dev/null; done", shell=True, check=True)
    lines = None
    import glob
    for f in glob.glob(td + "/*.txt"):
        L = open(f).read().splitlines()
        if any(l.startswith(".text." + sym) for l in L):
            lines = L
            break
    st = [i for i, l in enumerate(lines) if l.startswith(".text." + sym)][0]
    en = [i for i, l in enumerate(lines) if l.startswith(".text.") and i > st][0]
    cur, instrs = None, []
    for l in lines[st:en]:
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm:
            cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l):
            instrs.append(cur)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines())); h2 = rows[1]; data = rows[2:]
    ii, it, isamp = h2.index("Instructions Executed"), h2.index("Thread Instructions Executed"), h2.index("# Samples")
    if len(data) != len(instrs):
        print(f"WARNING: report has {len(data)} SASS lines, cubin {len(instrs)} — is this the build that was profiled?")
    def reg(key):
        if not key: return "?"
        for name, ff, a, b in regions:
            if key[0] == ff and a <= key[1] <= b: return name
        return key[0]
    bi, bt, bs, sc = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for k in range(min(len(data), len(instrs))):
        r = reg(instrs[k]); bi[r] += int(data[k][ii]); bt[r] += int(data[k][it]); bs[r] += int(data[k][isamp]); sc[r] += 1
    tot, tots = sum(bi.values()), sum(bs.values())
    with open(outp + ".regions.txt", "w") as f:
        f.write(f"# {rep}: executed warp instructions {tot:.4g}; share of instructions / of stall samples / active threads per instruction / static SASS instructions, by source region\n")
        for r, v in bi.most_common():
            f.write(f"{100 * v / tot:5.1f}% inst {100 * bs[r] / max(tots, 1):5.1f}% samples  {bt[r] / max(v, 1):5.1f} thr/inst  {sc[r]:5d} static  {r}\n")
    print(open(outp + ".regions.txt").read())
