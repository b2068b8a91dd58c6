"""Batched BA call (c5 windows) for several window counts: CUDA-graph replays, L2 flushed, per-stage events.
    python profiles/ab_windows.py 8 16 32 64      (env switches: PGBA_PC, PGBA_BATCH_GROUPS, ...)"""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import torch
import bench
dev = torch.device("cuda", 0)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("PGBA_")}}
for n in [int(a) for a in sys.argv[1:]] or [64]:
    arm = bench.GpuArm(bench.make_workload("c5", 0, n), dev)
    g = arm.capture()
    arm.timed_resident(g, 5)
    ms = arm.timed_resident(g, 40)
    out[str(n)] = {"us": round(1e3 * float(np.median(ms)), 1), "stages_us": {k: round(1e3 * v, 1) for k, v in arm.profiled(10).items()}}
print(json.dumps(out))
