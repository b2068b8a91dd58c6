"""Accuracy of the two Schur implementations (PGBA_SCHUR_UMMA=0: FFMA2, 1: tcgen05 3xTF32 with the accumulator in TMEM) at a forced chunk size:
relative errors of S, y, dX, dZ against the float64 oracle on c1 (ill-conditioned), c2, one c5 window.  Run with the env set."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
from cdvslam_b200 import synth, fastba
from oracle import ba_oracle
from tests.helpers import to_dev, f32_problem, rel_err
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("PGBA_")}}
for name, maker in (("c1", synth.config_c1), ("c2", synth.config_c2), ("c5w0", lambda: synth.config_c5_window(0))):
    p = maker(); d = to_dev(p); q = f32_problem(p)
    _, _, dbg = ba_oracle.ba(q["poses"], q["patches"], q["intrinsics"], q["target"], q["weight"], q["lmbda"], p.ii, p.jj, p.kk, p.t0, p.t1, 1, debug=True)
    o = dbg[0]
    g = fastba.linearize_debug(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, with_schur=True)
    S = g["S"].cpu().numpy().astype(np.float64)
    out[name] = {k: float(rel_err(g[k].cpu().numpy(), o[k])) for k in ("S", "y", "dX", "dZ")}
    out[name]["S_maxabs_rel"] = float(np.abs(np.tril(S) - np.tril(o["S"])).max() / np.abs(o["S"]).max())
    out[name]["cond"] = float(np.linalg.cond(S + np.diag(1e-4 * np.diag(S) + 1.0)))
print(json.dumps(out))
