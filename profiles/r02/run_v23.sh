mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -x -q) > gpurun_out/pytest_ba_v23.log 2>&1
tail -5 gpurun_out/pytest_ba_v23.log
python profiles/ab_c5.py c2 c5
PGBA_WARP_SOLVE=0 python profiles/ab_c5.py c2 c5
