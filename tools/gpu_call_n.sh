#!/bin/bash
# one N-GPU call: multi-GPU parity tests (N <= 2 only: they take a while), then the weak-scaling bench line
#   N=2 TAG=r02n tools/gpu_call_n.sh
set -u
N=${N:-2}; TAG=${TAG:-r02n}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
if [ "${TESTS:-1}" = 1 ]; then timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -4; fi
MMF_TRACE=1 timeout 300 $TR 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
python - ${TAG}_n${N} <<'P'
import json, sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    print(f, "value %.4e ms/step %.4f" % (d['value'], d['ms_per_step']), [round(x,4) for x in d['repeats']['ms_per_step']], "parity", d['parity'].get('max_ulp'), d['roofline']['stage_ms'], d['config'])
except Exception as e:
    print(f, "FAILED", e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
P
grep "mmf trace" gpurun_out/${TAG}_n${N}.err | grep x20 | head -2
