#!/bin/bash
# GPU call 2 of round 2: the new default parity tests (benchmark configuration against the oracle at 64^3/128^3/256^3,
# run-to-tMax-then-continue), the FP64 / SHFL / LDS peaks of this chip, and the profile set of the default kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_uniform_gpu.py -m gpu -x -q -k "bench_configuration or run_to_tmax" > gpurun_out/call2_parity.log 2>&1
echo "parity exit code: $?"; tail -5 gpurun_out/call2_parity.log
timeout 120 ./tools/ubench_fp64 > gpurun_out/ubench_fp64.txt 2>&1
echo "ubench exit code: $?"; tail -40 gpurun_out/ubench_fp64.txt
bash tools/gpu_profile_call.sh r02a
