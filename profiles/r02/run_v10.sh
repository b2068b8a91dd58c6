mkdir -p gpurun_out
python -m pytest tests/test_ba_gpu.py tests/test_pgo_gpu.py tests/test_diffba.py tests/test_parity_r2_gpu.py -m gpu -x -q -k "global or pgo or solve_system or on_the_gpu" 2>&1 | tail -3
python - <<'PY'
import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "cdv-slam_b200")]
import torch, bench
dev = torch.device("cuda", 0)
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("c4 single-launch back-substitution:", bench.extra_c4(dev, fl.zero_)["ms_per_call"], "ms")
PY
PGBA_BIG_BACK_SPLIT=1 python - <<'PY'
import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "cdv-slam_b200")]
import torch, bench
dev = torch.device("cuda", 0)
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("c4 per-panel back-substitution:", bench.extra_c4(dev, fl.zero_)["ms_per_call"], "ms")
PY
python profiles/time_pgo.py 2>&1 | tail -3
