#!/bin/bash
# full GPU suite + the three benchmark workloads.  usage: bash scripts/gpu_all.sh <tag>
TAG=${1:-all}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
for w in cfg2 cfg4 cfg3; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_$TAG.json"))
print("$w ms_per_step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],3), "ref_cuda", d.get("reference_cuda"))
print("   ", {k: round(v*1e3,1) for k,v in d["kernels_ms_per_step"].items()})
PY
  tail -2 gpurun_out/bench_${w}_$TAG.err
done
