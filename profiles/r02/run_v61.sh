python profiles/ab_windows.py 16 32 64
PGBA_GROUPS_IT_MAJOR=1 python profiles/ab_windows.py 16 32 64
python profiles/ab_windows.py 64
PGBA_GROUPS_IT_MAJOR=1 python profiles/ab_windows.py 64
