"""Per-CTA wall-clock trace of the numeric BA kernels of one c2 call (globaltimer): when each CTA entered, passed
pdl_wait and left.  Needs a -DPGBA_LIN_TIMING build:  PGBA_LIB=cdv-slam_b200/lib/libpgba_lintiming.so python profiles/cta_trace.py"""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import torch
import bench
from cdvslam_b200 import native
dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
arm = bench.GpuArm(bench.make_workload(wl, 0, 64), dev)
L = native.lib()
g = arm.capture()
rows = []
for it in range(6):
    arm.restore(); arm.flush_l2(); torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(); g.replay(); eb.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (4 * 3 * 512))()
    L.pgba_debug_cta_timestamps(buf)
    buf2 = (ctypes.c_ulonglong * (2 * 3 * 512))()
    L.pgba_debug_plan_cta_timestamps(buf2)
    rows.append(np.concatenate([np.array(list(buf2), dtype=np.float64).reshape(2, 3, 512),
                                np.array(list(buf), dtype=np.float64).reshape(4, 3, 512)]))
    ev_us = 1e3 * ea.elapsed_time(eb)
t = rows[-1]
names = {0: "plan_cluster", 1: "plan_cells", 2: "linearize #1", 3: "solve (last)", 4: "update", 5: "linearize #2 (fused update)"}
valid = t > 0
t0 = t[valid].min()
print("event-timed graph replay of this call: %.1f us" % ev_us)
print("all times in us relative to the first traced CTA entry; n = CTAs traced (<= 512)")
for k in (0, 1, 2, 3, 5, 4):
    m = valid[k, 0] & valid[k, 2]
    if not m.any():
        continue
    e, w, x = (t[k, 0][m] - t0) / 1e3, (t[k, 1][m] - t0) / 1e3, (t[k, 2][m] - t0) / 1e3
    print("%-30s n=%3d  entry %.1f..%.1f  wait passed %.1f..%.1f  exit %.1f..%.1f  | CTA body (exit - wait) min %.1f med %.1f max %.1f"
          % (names[k], m.sum(), e.min(), e.max(), w.min(), w.max(), x.min(), x.max(), (x - w).min(), np.median(x - w), (x - w).max()))
    busy = (x - w) > 2.0                     # CTAs that owned a chunk (the grid is sized for the worst-case chunk count)
    if busy.any() and busy.sum() < m.sum():
        b = (x - w)[busy]
        print("%-30s    busy CTAs: n=%3d  body min %.1f p25 %.1f med %.1f p75 %.1f max %.1f us" % ("", busy.sum(), b.min(), np.percentile(b, 25), np.median(b), np.percentile(b, 75), b.max()))
