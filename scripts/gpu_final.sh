#!/bin/bash
# Round-end pass: full suite, smoke, default bench (all legs), reference arm, other workloads, ncu launch list,
# ncu --set full of the tcgen05 kernels.  usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
bash scripts/gpu_verify.sh $TAG
for w in cfg4 cfg3; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 8 -c 4 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
