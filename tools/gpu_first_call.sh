#!/bin/bash
# First GPU call for the stage-kernel forms that were developed on the CPU emulator (tools/emu) and have
# never run on a GPU: parity first, then the sweep that decides whether they become the default.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_first_call.sh'
# (every step has its own, shorter timeout: a form that hangs on real hardware costs its own group only)
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export MMF_TEST_EXPERIMENTAL=1
# 1. parity of the experimental forms (d = plane-decoupled, h = + merged halo warp, w = + two y rows per warp, b / c = box with bodies), bounded
: > gpurun_out/experimental_parity.log
for group in "fused_steps and (d12 or d16 or d8)" "fused_steps and (h12 or h16 or h8)" "fused_steps and w8" "bodies" "primitives"; do
  echo "== $group" >> gpurun_out/experimental_parity.log
  timeout 300 python -m pytest tests/test_uniform_gpu.py -m gpu -x -q -k "$group" >> gpurun_out/experimental_parity.log 2>&1
  echo "parity [$group] exit code: $?" | tee -a gpurun_out/experimental_parity.log
done
echo "== generic fused stages" >> gpurun_out/experimental_parity.log
timeout 300 python -m pytest tests/test_generic_gpu.py -m gpu -x -q -k generic_fused >> gpurun_out/experimental_parity.log 2>&1
echo "parity [generic fused stages] exit code: $?" | tee -a gpurun_out/experimental_parity.log
grep -E "passed|failed|error|exit code" gpurun_out/experimental_parity.log | tail -14
# 1b. the reference's own main on the new pieces: golden strings and bitwise reference fields (bodies included)
#     with the body cases on the fused path and the writer's primitives from the device
MMF_UNIFORM_BODIES=1 MMF_DEVICE_PRIMITIVES=1 timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_reference_fields_gpu.py -m gpu -x -q > gpurun_out/experimental_dropin.log 2>&1
echo "drop-in exit code: $?" | tee -a gpurun_out/experimental_dropin.log
tail -3 gpurun_out/experimental_dropin.log
# 2. sweep at the benchmark size: per-stage kernel times, bitwise equality with the default mix
timeout 1200 python tools/stage_sweep.py --size 256 --steps 6 > gpurun_out/stage_sweep_256.jsonl 2> gpurun_out/stage_sweep_256.err
cat gpurun_out/stage_sweep_256.jsonl
# 3. z-chunk sensitivity of the two most promising mixes
for lz in 26 32 43 52 64; do
  timeout 300 python tools/stage_sweep.py --size 256 --steps 6 --variants "p16:p16:h12:h12@$lz,p16:p16:w8:w8@$lz,p16:p16:d12:d12@$lz,p16:p16:r12:r12@$lz" >> gpurun_out/stage_sweep_lz.jsonl 2>> gpurun_out/stage_sweep_256.err
done
cat gpurun_out/stage_sweep_lz.jsonl
timeout 600 python tools/generic_bench.py --size 128 --steps 6 > gpurun_out/generic_bench_128.jsonl 2>> gpurun_out/stage_sweep_256.err
# the same with stages 2 and 3 of the generic path fused (generic_stage_kernel)
MMF_GENERIC_FUSED=1 timeout 600 python tools/generic_bench.py --size 128 --steps 6 >> gpurun_out/generic_bench_128.jsonl 2>> gpurun_out/stage_sweep_256.err
cat gpurun_out/generic_bench_128.jsonl
# 4. a box with bodies: fused uniform path (form b, MMF_UNIFORM_BODIES=1) against the generic path
for mode in 1 2; do
  MMF_UNIFORM_BODIES=$mode timeout 600 python tools/generic_bench.py --size 128 --steps 6 --bodies >> gpurun_out/body_bench_128.jsonl 2>> gpurun_out/stage_sweep_256.err
done
cat gpurun_out/body_bench_128.jsonl
