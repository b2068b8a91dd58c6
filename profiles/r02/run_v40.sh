python profiles/ab_c5.py c2 c5
PGBA_LIB=cdv-slam_b200/lib/libpgba_prev.so python profiles/ab_c5.py c2 c5
python profiles/ab_c5.py c2
PGBA_LIB=cdv-slam_b200/lib/libpgba_prev.so python profiles/ab_c5.py c2
