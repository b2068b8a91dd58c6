mkdir -p gpurun_out
export PGBA_LIB=cdv-slam_b200/lib/libpgba_timing.so
echo "== plan_timing c5 direct"; python profiles/plan_timing.py c5 2>&1 | tail -16
echo "== plan_timing c5 direct, no cache (no hash)"; PGBA_PLAN_CACHE=0 python profiles/plan_timing.py c5 2>&1 | tail -16
echo "== cta_trace c5"; python profiles/cta_trace.py c5 2>&1 | tail -14
