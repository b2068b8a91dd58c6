mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "global") > gpurun_out/pytest_nd_v34.log 2>&1
tail -3 gpurun_out/pytest_nd_v34.log
timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -2
PGBA_LIB=cdv-slam_b200/lib/libpgba_nd64.so timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -1
PGBA_LIB=cdv-slam_b200/lib/libpgba_ndts.so python profiles/nd_timeline.py > gpurun_out/nd_timeline_v34.txt 2>&1
sed -n 24,28p gpurun_out/nd_timeline_v34.txt; tail -3 gpurun_out/nd_timeline_v34.txt
