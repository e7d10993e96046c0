#!/bin/bash
# quick GPU pass: tensor-path parity tests + bench (cfg2).  usage: bash scripts/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_path" > gpurun_out/pytest_tc_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc_$TAG.log
tail -4 gpurun_out/pytest_tc_$TAG.log
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"])
for k,v in d["kernels_ms_per_step"].items(): print(f"  {k:24s} {v*1e3:8.1f} us")
PY
tail -3 gpurun_out/bench_$TAG.err
