#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over a -k selection of the GPU tests.
# usage: bash scripts/gpu_sanitize.sh <tag> '<pytest -k expression>'
TAG=${1:-san}; SEL=${2:-grad_input_paths}
OUT=gpurun_out/sanitizer_$TAG.txt
echo "pytest -k '$SEL'" > $OUT
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> $OUT
  timeout 600 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_robustness.py -m gpu -q -x -k "$SEL" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 >> $OUT
done
cat $OUT
