import sys, os, time
sys.path[:0] = ['/root/repo', '/root/repo/cdv-slam_b200']
import numpy as np, torch
from cdvslam_b200 import synth, fastba
from tests.helpers import to_dev
p = synth.config_c4()
d = to_dev(p)
p0, q0 = d["poses"].clone(), d["patches"].clone()
def call(iters):
    fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, M=p.M, iterations=iters, eff_impl=True)
def run(iters, graph=None):
    d["poses"].copy_(p0); d["patches"].copy_(q0)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    a.record()
    if graph is None: call(iters)
    else: graph.replay()
    b.record(); tc = time.perf_counter() - t; torch.cuda.synchronize()
    return a.elapsed_time(b), tc * 1e3
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
if mode == "time":
    for it in (1, 1, 2, 2, 4):
        print("eager iterations", it, "gpu ms %.3f cpu enqueue ms %.3f" % run(it))
    for it in (1, 2):
        call(it); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): call(it)
        for _ in range(3): print("graph iterations", it, "gpu ms %.3f cpu ms %.3f" % run(it, g))
else:
    call(1); torch.cuda.synchronize()
    print("done")
