#!/bin/bash
# ncu --set full of one launch of each stage kernel for one MMF_STAGE_CFG:  bash tools/gpu_ncu_one.sh <tag> <cfg>
set -u
TAG=$1; CFG=$2
mkdir -p gpurun_out
MMF_STAGE_CFG=$CFG timeout 600 ncu --set full --clock-control none --import-source on -k regex:uniform_stage_kernel -s 9 -c 3 -f -o gpurun_out/${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --repeats 1 > gpurun_out/${TAG}.log 2>&1
echo "exit code: $?"; ls -la gpurun_out | grep "$TAG"
