#!/bin/bash
# First GPU pass: parity tests (generic path), reference goldens, bench, ncu launch list.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not tensor_path" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python oracle/gen_ref_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1
echo "golden rc=$?" >> gpurun_out/golden.log
tail -3 gpurun_out/golden.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_generic.json 2> gpurun_out/bench_generic.err
echo "bench rc=$?"
cat gpurun_out/bench_generic.json
tail -5 gpurun_out/bench_generic.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_generic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"
