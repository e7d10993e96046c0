#!/bin/bash
# quick GPU pass: tensor-path parity tests + per-kernel times (cfg2 by default).  usage: bash scripts/gpu_q.sh [workloads...]
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_path" 2>&1 | tail -3
for w in ${@:-cfg2}; do
  echo "== $w"; timeout 90 python scripts/kernel_times.py $w 10 2>&1 | tail -16
done | tee gpurun_out/kernel_times.log
