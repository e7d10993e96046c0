#!/bin/bash
# full GPU test suite with a hard time limit.  usage: bash scripts/gpu_tests.sh <tag>
TAG=${1:-t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -25 gpurun_out/pytest_gpu_$TAG.log
