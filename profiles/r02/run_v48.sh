python profiles/ab_windows.py 2 3 4 6 8 12 16 32 64
(timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "batch or window or group or plan") 2>&1 | tail -2
