mkdir -p gpurun_out
(time python -m pytest tests/test_ba_gpu.py -m gpu -x -q) > gpurun_out/pytest_ba_v3.log 2>&1
tail -5 gpurun_out/pytest_ba_v3.log
export PGBA_LIB=cdv-slam_b200/lib/libpgba_timing.so
echo "== plan_timing c2 direct"; python profiles/plan_timing.py c2 2>&1 | tail -16
echo "== plan_timing c5 direct"; python profiles/plan_timing.py c5 2>&1 | tail -16
echo "== plan_timing c2 old"; PGBA_PLAN_DIRECT=0 python profiles/plan_timing.py c2 2>&1 | tail -16
echo "== cta_trace c2 direct"; python profiles/cta_trace.py c2 2>&1 | tail -14
echo "== cta_trace c2 old"; PGBA_PLAN_DIRECT=0 python profiles/cta_trace.py c2 2>&1 | tail -14
