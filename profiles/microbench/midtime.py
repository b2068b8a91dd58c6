import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/cdv-slam_b200']
import numpy as np, torch
from cdvslam_b200 import synth, fastba
from tests.helpers import to_dev
for F in (60, 120, 200, 250):
    p = synth.make_problem("g", F, synth.global_edges(F, 96, F // 5, np.random.default_rng(3)), 1, F, 5, 96, eff_impl=True)
    d = to_dev(p)
    p0, q0 = d["poses"].clone(), d["patches"].clone()
    def call():
        fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, M=p.M, iterations=2, eff_impl=True)
    call(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): call()
    ts = []
    for _ in range(5):
        d["poses"].copy_(p0); d["patches"].copy_(q0); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("F", F, "ms %.3f" % np.median(ts), os.environ.get("PGBA_BIG_ND", "1"))
