#!/bin/bash
# CPU box: copy the artefacts of a `gpu_final.sh <tag>` run from gpurun_out/ into profiles/ (r2_* names).
# usage: bash scripts/collect_profiles.sh <tag>
T=${1:?tag}; G=gpurun_out
cp $G/bench_cfg2_$T.json profiles/r2_bench_cfg2.json
cp $G/bench_refarm_$T.json profiles/r2_bench_refarm.json
cp $G/bench_cfg3_$T.json profiles/r2_bench_cfg3.json
cp $G/bench_cfg4_$T.json profiles/r2_bench_cfg4.json
cp $G/bench_cfg2_bf16_$T.json profiles/r2_bench_cfg2_bf16.json
cp $G/bench_cfg2_cl_$T.json profiles/r2_bench_cfg2_channels_last.json
cp $G/bench_cfg2_sustained_$T.json profiles/r2_bench_cfg2_sustained.json
cp $G/launches_$T.csv profiles/r2_launches_cfg2.csv
python scripts/launch_summary.py profiles/r2_launches_cfg2.csv > profiles/r2_launches_cfg2_summary.txt
cp $G/launches_smoke_$T.csv profiles/r2_launches_smoke.csv
python scripts/launch_summary.py profiles/r2_launches_smoke.csv > profiles/r2_launches_smoke_summary.txt
python scripts/ncu_summary.py $G/prof_$T.ncu-rep > profiles/r2_ncu_full_cfg2_kernels.txt 2>/dev/null
python scripts/ncu_summary.py $G/prof_cfg4_$T.ncu-rep > profiles/r2_ncu_full_cfg4_3d_kernels.txt 2>/dev/null
python scripts/ncu_traffic.py $G/prof_$T.ncu-rep cfg2
python scripts/ncu_traffic.py $G/prof_cfg4_$T.ncu-rep cfg4
python scripts/sass_counts.py > profiles/r2_sass_opcounts.txt
cp $G/r2_parity_margins.json profiles/r2_parity_margins.json
(tail -5 $G/pytest_gpu_$T.log; tail -1 $G/smoke_$T.log) > profiles/r2_pytest_gpu.txt
