"""One eager global-BA call (config c4) for ncu launch lists:  ncu --metrics gpu__time_duration.sum ... python profiles/ncu_target_c4.py"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import torch
from cdvslam_b200 import synth, fastba
p = synth.config_c4()
d = synth.to_torch(p, torch.device("cuda", 0))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"],
              p.t0, p.t1, M=p.M, iterations=2, eff_impl=True)
torch.cuda.synchronize()
print("done c4")
