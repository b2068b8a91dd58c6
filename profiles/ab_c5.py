"""A/B timing of the batched (c5) and single-window (c2) BA call for the library given by PGBA_LIB / env switches:
CUDA-graph replays, L2 flushed before every replay, per-stage events from pgba_ba_solve_profiled."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import torch
import bench
dev = torch.device("cuda", 0)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("PGBA_")}}
for wl in (sys.argv[1:] or ["c5", "c2"]):
    arm = bench.GpuArm(bench.make_workload(wl, 0, 64), dev)
    g = arm.capture()
    arm.timed_resident(g, 5)
    ms = arm.timed_resident(g, 50 if wl == "c5" else 200)
    out[wl] = {"ms_median": float(np.median(ms)), "ms_mean": float(np.mean(ms)),
               "stages": {k: round(v, 5) for k, v in arm.profiled(20).items()}}
print(json.dumps(out))
