#!/bin/bash
# One GPU call that brings back what profiles/ is made of, for the default kernels or for a chosen mix:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_profile_call.sh r02a [p16:p16:h12:h12]'
# then, here:  python tools/make_profiles.py --tag r02a --rep gpurun_out/r02a_stage.ncu-rep --launches gpurun_out/r02a_launches.csv
# 1. the bench line of the configuration (never taken under a profiler), 2. the launch list of the same command
# (per-launch times are cold-cache and serialised: only the SHARES are compared with the bench), 3. ncu --set full of
# one launch of each stage kernel.  One GPU only: ncu replays every kernel about 40 times.
set -u
TAG=${1:-r02a}
CFG=${2:-}
mkdir -p gpurun_out
if [ -n "$CFG" ]; then export MMF_STAGE_CFG="$CFG"; fi
echo "profile call $TAG, MMF_STAGE_CFG='${MMF_STAGE_CFG:-}'" | tee gpurun_out/${TAG}_info.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv >> gpurun_out/${TAG}_info.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit code: $?"; tail -c 2000 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bodies > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list exit code: $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_stage_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_stage \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-bodies > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu --set full exit code: $?"; ls -la gpurun_out/ | grep "$TAG"
