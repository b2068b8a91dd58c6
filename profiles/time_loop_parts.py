"""Components of the c3 update loop (slam.py:316-337, 470-496) on the c2 graph, each as its own CUDA graph, warm (no L2
flush: inside the 12-update loop the data of the previous update is resident) and cold (L2 flushed before every replay)."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import torch
from cdvslam_b200 import synth, fastba, altcorr, native
import bench

dev = torch.device("cuda", 0)
p = synth.config_c2()
d = synth.to_torch(p, dev)
gmap, pyr = synth.make_fmaps(p, C=24)
g = torch.as_tensor(gmap, device=dev)[None].half()
f0 = torch.as_tensor(pyr[0], device=dev)[None].half().contiguous()
f1 = torch.as_tensor(pyr[1], device=dev)[None].half().contiguous()
ring = altcorr.PyramidRing([f0, f1])
coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"], clamp_depth=True)
delta = torch.randn((1, p.E, 2), device=dev)
weight = torch.rand((1, p.E, 2), device=dev)
target = coords[:, :, :, 1, 1] + delta
p0, q0 = d["poses"].clone(), d["patches"].clone()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
parts = {
    "reproject": lambda: fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"], clamp_depth=True),
    "ring_update_one_slot": lambda: ring.update(21),
    "ring_lookup": lambda: ring.lookup(g, coords, d["kk"], d["jj"], 3),
    "corr_pyramid2_api": lambda: altcorr.corr_pyramid2(g, [f0, f1], coords, d["kk"], d["jj"], 3),
    "neighbors": lambda: fastba.neighbors(d["kk"], d["jj"]),
    "target_add": lambda: coords[:, :, :, 1, 1] + delta,
    "BA_2it_plan_reused": lambda: fastba.BA(d["poses"], d["patches"], d["intrinsics"], target, weight, d["lmbda"], d["ii"], d["jj"],
                                            d["kk"], p.t0, p.t1, M=p.M, iterations=2),
}
out = {}
for name, fn in parts.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    reset = (lambda: (d["poses"].copy_(p0), d["patches"].copy_(q0))) if name.startswith("BA") else (lambda: None)
    out[name] = {"warm_us": 1e3 * bench.timed_events(gr.replay, 30, before=reset),
                 "cold_us": 1e3 * bench.timed_events(gr.replay, 30, before=lambda: (reset(), flush.zero_()))}
print(json.dumps(out, indent=1))
