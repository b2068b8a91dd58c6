mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v7.log 2>&1
tail -4 gpurun_out/pytest_gpu_v7.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo bench rc=$?
c5() { python bench.py --steps 20 --warmup 5 --workload c5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $1', round(b['ms_per_step']*1e3,1), 'us; reuse', round(b['plan_reuse']['ms_per_step']*1e3,1), {k: round(v*1e3,1) for k,v in b['stages_ms'].items()})"; }
for g in 1 3 4; do PGBA_BATCH_GROUPS=$g c5 "groups=$g"; done
export PGBA_LIB=cdv-slam_b200/lib/libpgba_timing.so
echo "== plan_timing c5 direct"; python profiles/plan_timing.py c5 2>&1 | tail -16
