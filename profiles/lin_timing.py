"""Phase clocks of linearize_kernel (CTA 0) on the c2 window; needs a -DPGBA_LIN_TIMING build passed through PGBA_LIB."""
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
    arm.restore(); arm.flush_l2(); torch.cuda.synchronize()
    arm.call(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    L.pgba_debug_lin_timestamps(buf)
    t = np.array(list(buf)[:18], dtype=np.float64)
    acc.append(t - t[0])
acc = np.median(np.array(acc[2:]), axis=0)
names = ["start (after pdl_wait)", "chunk loaded, sync", "rel poses, sync", "patches staged, sync", "edge loop + H partials, sync",
         "H reduced, sync", "Q per patch, sync", "E copy + Schur + y", "sync", "B1 (AH), sync", "B2 (Bii), sync", "B3 scatter",
         "umma: corner done", "umma: staged values read", "umma: operands written, fenced", "umma: MMAs issued (thread 0)",
         "umma: accumulator complete", "umma: epilogue done"]
for n, a, d in zip(names, acc, np.diff(np.concatenate([[0], acc]))):
    print("%-36s t=%8.0f cyc  (+%6.0f)" % (n, a, d))
