#!/usr/bin/env python
"""Per-source-line view of one kernel launch of an Nsight Compute report: executed warp instructions, active threads per instruction and
stall samples for every source line (SASS lines of the report joined with nvdisasm's line table of the cubin that was profiled).
    python tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SYMBOL [launch_index=0] [min_share_pct=0.2]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, so, sym = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.2
td = tempfile.mkdtemp()
subprocess.run(f"cd {td} && cuobjdump -xelf all {so} >/dev/null && for f in *.cubin; do nvdisasm -g -c $f > $f.txt 2>/dev/null; done", shell=True, check=True)
lines = None
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
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h2 = rows[hi]; data = rows[hi + 1:]
ii, it, isamp = h2.index("Instructions Executed"), h2.index("Thread Instructions Executed"), h2.index("# Samples")
if len(data) != len(instrs):
    print(f"WARNING: report has {len(data)} SASS lines, cubin {len(instrs)}")
bi, bt, bs, sc = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for k in range(min(len(data), len(instrs))):
    r = instrs[k] or ("?", 0); bi[r] += int(data[k][ii]); bt[r] += int(data[k][it]); bs[r] += int(data[k][isamp]); sc[r] += 1
tot, tots = sum(bi.values()), sum(bs.values())
print(f"# launch {launch}: {tot:.4g} warp instructions, {sum(bt.values()) / max(tot, 1):.2f} threads/inst")
for r in sorted(bi):
    if 100 * bi[r] / tot >= minpct:
        print(f"{r[0]}:{r[1]:5d}  {100 * bi[r] / tot:5.2f}% inst {100 * bs[r] / max(tots, 1):5.2f}% samples {bt[r] / max(bi[r], 1):5.1f} thr/inst {sc[r]:4d} static")
