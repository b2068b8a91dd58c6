python profiles/ab_windows.py 2 4 8 12 16 24 32 64
PGBA_BATCH_GROUPS=2 python profiles/ab_windows.py 8 12 16
PGBA_BATCH_GROUPS=4 python profiles/ab_windows.py 16 24
(timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_parity_r2_gpu.py -m gpu -x -q) 2>&1 | tail -2
