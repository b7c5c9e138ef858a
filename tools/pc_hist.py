"""Where a lane's iterations go: runs the host emulation of the search state machine (tests/host_emu/emu.cpp, BT_EMU_PC_HIST=1) on the
bench workload (-n 2 -k 1, 100-bp synthetic reads, the bench index) and prints transitions per read by state and LF kind, plus the
per-read distribution.  Development aid for the kernel work; not a measurement.  Usage: python tools/pc_hist.py [n_reads=20000]"""
import sys, os, types
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
os.environ['BT_EMU_PC_HIST'] = '1'; os.environ['BT_EMU_PROFILE'] = '1'
import numpy as np
from helpers import HostEmu, Policy
import bench
base, name = bench.fallback_index()
genome = bench.load_genome(base)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
codes, quals, offs, seeds, names = bench.make_reads(genome, n, seed=12345)
class B:
    def __len__(self): return n
batch = B()
batch.seq_codes, batch.qual_cat, batch.offs, batch.seeds = codes, quals, offs, seeds
emu = HostEmu()
pol = Policy(mode=1, mms=2, khits=1)
res, flags = emu.align(base, batch, pol)
print("aligned", int((res.nhits_per_read > 0).sum()) if hasattr(res, 'nhits_per_read') else '?', "of", n)
ipr = emu.iters_per_read
print("iters/read: mean %.1f p50 %d p90 %d p99 %d max %d" % (ipr.mean(), np.percentile(ipr, 50), np.percentile(ipr, 90), np.percentile(ipr, 99), ipr.max()))
import ctypes
emu.L.emu_prof.restype = ctypes.POINTER(ctypes.c_ulonglong)
pr = emu.L.emu_prof()
print("per read: BTLOOP %.2f (scans %.2f, scan positions %.2f) CHILD_RET %.2f (rescans %.2f, rescan positions %.2f)" % tuple(pr[i] / n for i in range(6)))
print("per read: scan positions with live alternatives %.2f; rescan positions at <= lowest quality %.2f, of which live %.2f" % tuple(pr[i] / n for i in (6, 7, 8)))
