"""Short program for ncu captures: a few direct (no CUDA graph) calls of the BA hot path.
    python profiles/ncu_target.py [c2|c5|corr] [calls]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import torch
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if what in ("c2", "c5"):
    arm = bench.GpuArm(bench.make_workload(what, 0, 64), dev)
    for _ in range(calls):
        arm.restore()
        arm.flush_l2()
        arm.call()
    torch.cuda.synchronize()
else:
    import numpy as np
    from cdvslam_b200 import synth, fastba, altcorr
    from tests.helpers import to_dev
    p = synth.config_c2()
    d = to_dev(p)
    C, dt = (24, torch.float16) if what == "corr" else ((128, torch.float16) if what == "corr128h" else (128, torch.float32))
    gmap, pyr = synth.make_fmaps(p, C=C)
    g = torch.as_tensor(gmap, device=dev)[None].to(dt)
    f0 = torch.as_tensor(pyr[0], device=dev)[None].to(dt)
    f1 = torch.as_tensor(pyr[1], device=dev)[None].to(dt)
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    for _ in range(calls):
        altcorr.corr_pyramid2(g, [f0, f1], coords, d["kk"], d["jj"], 3)
        altcorr.corr(g, f0, coords, d["kk"], d["jj"], 3)
    torch.cuda.synchronize()
print("done", what)
