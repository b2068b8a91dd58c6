#!/usr/bin/env bash
# Round-2 baseline on the GPU box: GPU parity suite, the driver's bench commands (both arms), launch list of bench.py.
# Usage (from the dev container): gpurun --timeout 1500 -- 'bash profiles/r02/run_baseline.sh v1'
tag="${1:-v1}"
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu rc=$?"
