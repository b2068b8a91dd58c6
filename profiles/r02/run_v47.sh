PGBA_PLAN_DIRECT=1 python profiles/ab_windows.py 2 4 6 8
PGBA_PLAN_DIRECT=1 PGBA_BATCH_GROUPS=2 python profiles/ab_windows.py 4 6 8
PGBA_PLAN_DIRECT=1 PGBA_BATCH_GROUPS=2 PGBA_PC=128 python profiles/ab_windows.py 6 8
python profiles/ab_windows.py 6
