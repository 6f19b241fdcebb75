#!/bin/bash
# ncu --set full of one launch per stage for a few kernel forms (source-level stall samples)
set -u
mkdir -p gpurun_out
for cfg in ${CFGS:-p16:p16:h12:h12 p16:p16:d12:d12}; do
  tag=$(echo $cfg | tr ':' '_')
  MMF_STAGE_CFG=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:uniform_stage_kernel -s 9 -c 3 -f -o gpurun_out/r02b_$tag \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02b_$tag.log 2>&1
  echo "$cfg exit code: $?"
done
ls -la gpurun_out | grep r02b
