"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np, torch
from cdvslam_b200 import synth, fastba, altcorr
from tests.helpers import to_dev

def run(p, iters=2, eff=False):
    d = to_dev(p)
    fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, M=p.M, iterations=iters, eff_impl=eff)
    torch.cuda.synchronize()
    return d

p = synth.small_problem(seed=3, F=8, M=32, t0=3, lifetime=5)
d = run(p)
fastba.linearize_debug(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1)
run(synth.config_c2())
rng = np.random.default_rng(5)
g = synth.make_problem("global", 75, synth.global_edges(75, 12, 10, rng), 1, 75, 5, 12, eff_impl=True)
run(g, 2, True)
q = synth.small_problem(seed=4, F=3, M=40, t0=3, lifetime=3); q.t0 = q.t1 = 3
run(q, 2)
# batched
probs = [synth.small_problem(seed=20 + s, F=8, M=16, t0=3, lifetime=5) for s in range(3)]
ds = [to_dev(x) for x in probs]
cat = lambda k: torch.cat([x[k] for x in ds], 0).contiguous()
idx = lambda k: torch.stack([x[k] for x in ds], 0).contiguous()
fastba.BA_batched(cat("poses"), cat("patches"), cat("intrinsics"), cat("target"), cat("weight"), ds[0]["lmbda"], idx("ii"), idx("jj"), idx("kk"), probs[0].t0, probs[0].t1, M=16, iterations=2)
# corr: staged f32/f16, tiled f16, generic radius, backward, patchify
for C, dt in ((24, np.float32), (24, np.float16), (8, np.float16)):
    gm, pyr = synth.make_fmaps(p, C=C, n_mem=8, dtype=dt)
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    coords[0, :10] -= 70.0; coords[0, 10:20] += 60.0
    G = torch.as_tensor(gm, device="cuda")[None]; M0 = torch.as_tensor(pyr[0], device="cuda")[None]; M1 = torch.as_tensor(pyr[1], device="cuda")[None]
    for tiled in ("0", "1"):
        os.environ["PCORR_TILED"] = tiled
        altcorr.corr(G, M0, coords, d["kk"], d["jj"], 3)
        altcorr.corr_pyramid2(G, [M0, M1], coords, d["kk"], d["jj"], 3)
    altcorr.corr(G, M0, coords, d["kk"], d["jj"], 1)
Gf = torch.as_tensor(synth.make_fmaps(p, C=8, n_mem=8)[0], device="cuda")[None].requires_grad_(True)
Mf = torch.as_tensor(synth.make_fmaps(p, C=8, n_mem=8)[1][0], device="cuda")[None].requires_grad_(True)
altcorr.corr(Gf, Mf, coords, d["kk"], d["jj"], 3).sum().backward()
net = torch.randn(2, 12, 30, 40, device="cuda", requires_grad=True)
pc = torch.rand(2, 50, 2, device="cuda") * 45 - 3
altcorr.patchify(net, pc, 1).sum().backward()
fastba.neighbors(d["kk"], d["jj"])
torch.cuda.synchronize()
print("sanitizer target done")
