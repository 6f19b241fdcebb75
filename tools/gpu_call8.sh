#!/bin/bash
# one 8-GPU call: weak scaling at 8 (copy-engine exchange vs push kernel), strong scaling of 512^3 at 8, weak at 4
set -u
mkdir -p gpurun_out
T=${TAG:-r02f}
run() { # name nproc port extra-env... -- bench args
  name=$1; np=$2; port=$3; shift 3
  env MMF_TRACE=1 "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline $ARGS > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python - ${T}_$name <<'P'
import json, sys, re
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    print(f, "value %.4e ms/step %.4f" % (d['value'], d['ms_per_step']), [round(x,4) for x in d['repeats']['ms_per_step']], "parity", d['parity'].get('max_ulp'), d['config']['decomposition'][:12])
    t=open(f'gpurun_out/{f}.err').read(); vals={}
    for m in re.finditer(r'(\S+?)=([0-9.]+)\(x(\d+)\)', t):
        if m.group(3) in ('20','19','60'): vals.setdefault(m.group(1),[]).append(float(m.group(2)))
    print("   ", "  ".join("%s %.3f-%.3f" % (k,min(v),max(v)) for k,v in vals.items()))
except Exception as e:
    print(f, "FAILED", e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
P
}
ARGS="" run n8_weak_dma 8 29541 MMF_X=0
ARGS="--no-parity" run n8_weak_pushk 8 29542 MMF_DMA_PUSH=0
ARGS="--no-parity --scaling strong --size 512" run n8_strong512_dma 8 29543 MMF_X=0
ARGS="--no-parity" run n4_weak_dma 4 29544 MMF_X=0
