#!/bin/bash
# how the step time and the clocks behave when the load lasts: long runs with nvidia-smi sampling every 20 ms
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 20 > gpurun_out/sustained_clocks.csv &
SMI=$!
python bench.py --steps ${STEPS:-200} --warmup 5 --repeats 5 --no-cpu-baseline --no-parity > gpurun_out/sustained_bench.json 2>&1
kill $SMI
python - <<'P'
import json
d=json.loads(open('gpurun_out/sustained_bench.json').read().strip().splitlines()[-1])
print("ms/step per repeat:", [round(x,4) for x in d['repeats']['ms_per_step']])
rows=[l.strip().split(', ') for l in open('gpurun_out/sustained_clocks.csv').readlines()[1:]]
import collections
sm=[int(r[1].split()[0]) for r in rows if len(r)>4]
pw=[float(r[3].split()[0]) for r in rows if len(r)>4]
print("samples", len(sm), "sm MHz min/median/max", min(sm), sorted(sm)[len(sm)//2], max(sm), "power max", max(pw))
busy=[(s,p,r[5],r[6]) for s,p,r in zip(sm,pw,rows) if p>400]
print("under load (>400 W):", len(busy), "sm MHz", collections.Counter(b[0] for b in busy).most_common(6), "reasons", collections.Counter((b[2],b[3]) for b in busy).most_common(4))
print("power under load min/median/max", min(b[1] for b in busy), sorted(b[1] for b in busy)[len(busy)//2], max(b[1] for b in busy))
P
