#!/bin/bash
# a box with bodies on the fused path: parity of kernel form 'b', then its step times next to the
# same box without bodies and to the generic path (one GPU)
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02k}
timeout 600 python -m pytest tests/test_uniform_gpu.py tests/test_reference_fields_gpu.py tests/test_dropin_gpu.py -x -q -m gpu -k "bodies or body or reference_fields or dropin" > gpurun_out/${TAG}_bodies_pytest.log 2>&1
echo "pytest exit code: $?"; tail -6 gpurun_out/${TAG}_bodies_pytest.log
O=gpurun_out/${TAG}_bodies.jsonl
: > $O
for n in 192 256; do
  timeout 300 python tools/generic_bench.py --size $n --steps 20 --no-generic >> $O 2>gpurun_out/${TAG}_bodies.err
  timeout 300 python tools/generic_bench.py --size $n --steps 20 --bodies --no-generic >> $O 2>>gpurun_out/${TAG}_bodies.err
done
timeout 600 python tools/generic_bench.py --size 192 --steps 10 --bodies >> $O 2>>gpurun_out/${TAG}_bodies.err
cat $O; tail -3 gpurun_out/${TAG}_bodies.err
