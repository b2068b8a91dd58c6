"""Phase timestamps (globaltimer, ns) of plan_cluster_kernel on the c2 window; needs a -DPGBA_PLAN_TIMING build:
    PGBA_LIB=cdv-slam_b200/lib/libpgba_timing.so python profiles/plan_timing.py"""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import torch
import bench
from cdvslam_b200 import native

dev = torch.device("cuda", 0)
arm = bench.GpuArm(bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2", 0, 64), dev)
L = native.lib()
acc = []
for it in range(8):
    arm.restore(); native.invalidate_plan_cache(); arm.flush_l2(); torch.cuda.synchronize()
    arm.call(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    L.pgba_debug_plan_timestamps(buf)
    t = np.array(list(buf)[:16], dtype=np.float64)
    acc.append(t - t[0])
acc = np.median(np.array(acc[2:]), axis=0)
direct = os.environ.get("PGBA_PLAN_DIRECT", "1") != "0"
names_direct = ["start", "D0 zero+load+pack", "sync B1", "D1 frame min/max", "sync B2", "D2 chunk table", "D3 zero bitmaps", "D3 presence bits",
                "sync B3", "D4 kx/slots/fill", "sync B4", "D5 stores", "D4 OR+counts+scans", "D5 verify", "-", "-"]
names = names_direct if direct else ["start", "zero+load issued", "sync0", "P1 frames", "sync1", "P2 table", "sync2", "P3 count", "sync3", "P4 scan", "sync4", "P5 scatter", "P2a loads", "P2b scan", "P2c table", "-"]
order = np.argsort(acc)
prev = 0.0
for i in order:
    if names[i] == "-" or acc[i] < 0:
        continue
    print("%-22s t=%8.0f ns  (+%6.0f)" % (names[i], acc[i], acc[i] - prev))
    prev = acc[i]
