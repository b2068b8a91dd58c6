mkdir -p gpurun_out
(python -m pytest tests/test_corr_gpu.py tests/test_parity_r2_gpu.py -m gpu -x -q) > gpurun_out/pytest_corr_v9.log 2>&1
tail -4 gpurun_out/pytest_corr_v9.log
python profiles/time_corr.py 2>&1 | tail -12
