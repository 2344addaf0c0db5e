"""Diagnostic: per-step scan-kernel time over a long back-to-back loop + nvidia-smi clocks/power (is the sustained
number power/clock limited?).  python tools/diag_sustained.py [steps]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densephrases_b200.sharded import ShardedIvfPq
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
wl = bench.workload(1)
ix = ShardedIvfPq(wl["nlist"]); ix.build_synthetic(bench.opq_matrix(bench.SEED_INDEX), wl["lens"], bench.SEED_INDEX); ix.nprobe = 256
dev = torch.device("cuda", 0)
Q = bench.make_queries(ix, wl, 16, 0, 1, dev)
for s in range(3): ix.search_device(Q[s], 10)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
ix.local.set_profile(True)
t0 = time.time()
allms = []
for rep in range(steps // 60 + 1):
    for s in range(60): ix.search_device(Q[s % 16], 10)
    torch.cuda.synchronize()
    allms += [float(v) for v in ix.local.profile_scan_ms()][-60:]
t1 = time.time()
time.sleep(0.3); p.terminate(); out = p.communicate()[0]
a = np.array(allms)
print("steps", len(a), "wall %.3fs" % (t1 - t0), "scan ms: first10 %.3f  mid %.3f  last10 %.3f  min %.3f max %.3f" % (a[:10].mean(), a[len(a)//2-5:len(a)//2+5].mean(), a[-10:].mean(), a.min(), a.max()))
print("GB/s first10 %.0f last10 %.0f" % (38.4 / a[:10].mean() * 1000, 38.4 / a[-10:].mean() * 1000))
lines = out.strip().splitlines()
print("\n".join(lines[::max(1, len(lines)//25)]))
