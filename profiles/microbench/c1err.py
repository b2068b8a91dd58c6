import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/cdv-slam_b200']
import numpy as np, torch
from cdvslam_b200 import synth, fastba
from oracle import ba_oracle
from tests.helpers import to_dev, f32_problem, rel_err
p = synth.config_c1()
q = f32_problem(p)
_, _, dbg = ba_oracle.ba(q["poses"], q["patches"], q["intrinsics"], q["target"], q["weight"], q["lmbda"], p.ii, p.jj, p.kk, p.t0, p.t1, 1, debug=True)
o = dbg[0]
d = to_dev(p)
for it in range(6):
    g = fastba.linearize_debug(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, with_schur=True)
    S = g["S"].cpu().numpy().astype(np.float64); y = g["y"].cpu().numpy().astype(np.float64)
    Sd = S + np.diag(1e-4 * np.diag(S) + 1.0)
    dx64 = np.linalg.solve(Sd, y)
    print("pc=%s S %.2e y %.2e dX %.2e dZ %.2e | dX from fp64 solve of GPU S,y: %.2e | cond %.1e" % (os.environ.get("PGBA_PC"), rel_err(S, o["S"]), rel_err(y, o["y"]), rel_err(g["dX"].cpu().numpy(), o["dX"]), rel_err(g["dZ"].cpu().numpy(), o["dZ"]), rel_err(dx64, o["dX"]), np.linalg.cond(Sd)))
