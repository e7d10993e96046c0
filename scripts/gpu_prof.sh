#!/bin/bash
# GPU pass: tensor-path parity tests, bench (cfg2), ncu launch list, ncu --set full of the tcgen05 kernels.
# usage: bash scripts/gpu_prof.sh <tag> [kernel-regex]
TAG=${1:-r1}
REGEX=${2:-k_tc_}
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_path" > gpurun_out/pytest_tc_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc_$TAG.log
tail -3 gpurun_out/pytest_tc_$TAG.log
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 8 -c 4 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
