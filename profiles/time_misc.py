"""Timing probes (CUDA events, graph replays, L2 flushed): fastba.neighbors on the c2 / c4 edge lists; fastba.BA on c2 with the
plan tables rebuilt (cold) or reused (warm), with the reuse count of every call."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import torch
from cdvslam_b200 import synth, fastba, native
import bench

dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("PGBA_")}}


def graph_time(fn, n=30, before=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return bench.timed_events(g.replay, n, before=before)


for name, p in (("c2", synth.config_c2()), ("c4", synth.config_c4())):
    kk = torch.as_tensor(p.kk, device=dev); jj = torch.as_tensor(p.jj, device=dev)
    out["neighbors_%s_ms" % name] = graph_time(lambda: fastba.neighbors(kk, jj), before=flush_buf.zero_)

p = synth.config_c2()
d = synth.to_torch(p, dev)
p0, q0 = d["poses"].clone(), d["patches"].clone()
call = lambda: fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"],
                         d["kk"], p.t0, p.t1, M=p.M, iterations=2)
for _ in range(3):
    call()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    call()


def reset(cold):
    d["poses"].copy_(p0); d["patches"].copy_(q0)
    if cold:
        native.invalidate_plan_cache()
    flush_buf.zero_()


out["ba_c2_cold_ms"] = bench.timed_events(g.replay, 50, before=lambda: reset(True))
out["ba_c2_warm_ms"] = bench.timed_events(g.replay, 50, before=lambda: reset(False))
hits = []
for cold in (True, False, False, True, False):
    reset(cold); call(); torch.cuda.synchronize(); hits.append(fastba.last_plan_hits())
out["eager_hits_cold_warm_warm_cold_warm"] = hits
print(json.dumps(out))
