#!/bin/bash
# Current-state pass: GPU parity tests, bench (cfg2), ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_state.json 2> gpurun_out/bench_state.err; echo "bench rc=$?"
cat gpurun_out/bench_state.json
tail -3 gpurun_out/bench_state.err
