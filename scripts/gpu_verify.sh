#!/bin/bash
# Verification pass: full GPU suite, smoke, default bench (all legs), reference arm, ncu launch list.
# usage: bash scripts/gpu_verify.sh <tag>
TAG=${1:-v}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc > gpurun_out/nproc_$TAG.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_cfg2_$TAG.json; tail -3 gpurun_out/bench_cfg2_$TAG.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_refarm_$TAG.json 2> gpurun_out/bench_refarm_$TAG.err; echo "refarm rc=$?"
cat gpurun_out/bench_refarm_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
