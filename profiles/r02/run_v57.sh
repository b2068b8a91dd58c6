PGBA_BATCH_GROUPS=2 python profiles/ab_windows.py 32 64
PGBA_BATCH_GROUPS=3 python profiles/ab_windows.py 32 64
python profiles/ab_windows.py 32 48
