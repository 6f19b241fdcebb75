#!/bin/bash
# one 2-GPU call: multi-GPU parity, then the weak-scaling bench with and without the fused exchange
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -4
MMF_TRACE=1 timeout 300 $TR 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_n2_dma.json 2> gpurun_out/r02i_n2_dma.err
MMF_TRACE=1 MMF_DMA_PUSH=0 timeout 300 $TR 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r02i_n2_push.json 2> gpurun_out/r02i_n2_push.err
for f in r02i_n2_dma r02i_n2_push; do python - $f <<'P'
import json, sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    print(f, "value %.4e ms/step %.4f" % (d['value'], d['ms_per_step']), [round(x,4) for x in d['repeats']['ms_per_step']], "parity", d['parity'].get('max_ulp'), d['roofline']['stage_ms'])
except Exception as e:
    print(f, "FAILED", e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
P
grep "mmf trace" gpurun_out/$f.err | grep x20 | head -2
done
