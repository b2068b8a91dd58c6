import sys, os, time
sys.path[:0] = ['/root/repo', '/root/repo/cdv-slam_b200']
import numpy as np, torch
from cdvslam_b200 import synth, fastba
from oracle import ba_oracle
from tests.helpers import to_dev, f32_problem, rel_err
p = synth.config_c4()
q = f32_problem(p)
d = to_dev(p)
p0, q0 = d["poses"].clone(), d["patches"].clone()
def run(iters):
    d["poses"].copy_(p0); d["patches"].copy_(q0)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, M=p.M, iterations=iters, eff_impl=True)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
for it in (1, 2, 2, 2):
    print("iterations", it, "ms", run(it))
for iters in (1, 2):
    run(iters)
    o_poses, o_patches = ba_oracle.ba(q["poses"], q["patches"], q["intrinsics"], q["target"], q["weight"], q["lmbda"], p.ii, p.jj, p.kk, p.t0, p.t1, iters)
    poses = d["poses"][0].cpu().numpy().astype(np.float64); patches = d["patches"][0].cpu().numpy().astype(np.float64)
    e = np.abs(patches[:, 2, 0, 0] - o_patches[:, 2, 0, 0]) / np.abs(o_patches[:, 2, 0, 0])
    print("iters", iters, "pose rel", rel_err(poses, o_poses), "depth rel err pct 50/99/99.9/max", np.percentile(e, [50, 99, 99.9, 100]), "n>1e-3:", int((e > 1e-3).sum()))
    bad = np.argsort(e)[-5:]
    print(" worst patches", bad, e[bad], "depth oracle", o_patches[bad, 2, 0, 0], "gpu", patches[bad, 2, 0, 0], "init", np.asarray(p.patches)[bad, 2, 0, 0])
