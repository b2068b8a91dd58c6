mkdir -p gpurun_out
python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "arena or host_buffer" 2>&1 | tail -3
echo "== solve_bench"; profiles/microbench/bin/solve_bench 10 1; profiles/microbench/bin/solve_bench 10 64 | head -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo bench rc=$?
PGBA_BATCH_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:'linearize|solve_small|update_large|plan_direct' -c 7 -o gpurun_out/ncu_c5_v8 python profiles/ncu_target.py c5 1 > gpurun_out/ncu_c5_v8.log 2>&1; echo ncu c5 rc=$?
ncu --set full --clock-control none --import-source on -k regex:'linearize|solve_small|update_kernel|plan_cluster|plan_cells' -c 7 -o gpurun_out/ncu_c2_v8 python profiles/ncu_target.py c2 1 > gpurun_out/ncu_c2_v8.log 2>&1; echo ncu c2 rc=$?
ls -la gpurun_out/*.ncu-rep
