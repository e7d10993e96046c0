#!/bin/bash
# Round-end pass: gpu_verify.sh + the other workloads (cfg3 fp32, cfg4 3D, cfg2 in bf16, cfg2 with channels-last
# activations), a 2000-step sustained line, the smoke() launch list, and one `ncu --set full` capture of the tcgen05
# kernels.  usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
bash scripts/gpu_verify.sh $TAG
for w in cfg4 cfg3 cfg2_bf16; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --channels-last --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_cfg2_cl_$TAG.json 2> gpurun_out/bench_cfg2_cl_$TAG.err; echo "bench channels-last rc=$?"
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_cfg2_sustained_$TAG.json 2> gpurun_out/bench_cfg2_sustained_$TAG.err; echo "bench sustained rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_smoke_$TAG.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke_$TAG.log 2>&1; echo "ncu smoke rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_|k_col2im" -s 8 -c 4 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tc_|k_col2im" -s 8 -c 4 -f -o gpurun_out/prof_cfg4_$TAG python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_cfg4_$TAG.log 2>&1; echo "ncu full cfg4 rc=$?"
