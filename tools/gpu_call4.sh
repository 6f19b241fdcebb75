#!/bin/bash
# parity of the TMA-load form, then the sweep (every variant in its own process with a timeout)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_uniform_gpu.py -q -x -m gpu -k "fused_steps and (t12 or t16 or t8)" 2>&1 | tail -5
timeout 900 python tools/stage_sweep.py --size 256 --steps 6 --variants "${VARIANTS:-r12,t12,t16,t8,r16}" --variant-timeout 60 > gpurun_out/sweep_t.jsonl 2> gpurun_out/sweep_t.err
python - <<'P'
import json
for l in open('gpurun_out/sweep_t.jsonl'):
    d=json.loads(l)
    if 'variant' in d: print(d['variant'], round(d.get('ms_per_step',0),4), [round(x,4) for x in d.get('stage_ms',[])], d.get('bitwise_equal_to_first'), d.get('error','')[-300:])
P
