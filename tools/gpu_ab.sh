#!/bin/bash
# A/B timing that survives the power cap and box-to-box differences: the (library build, MMF_STAGE_CFG) pairs are run
# round-robin, ROUNDS times each, every run in its own process; the table shows the median and the best run of each.
#   bash tools/gpu_ab.sh "base:t16 base:r12 skipdead0:t16" [rounds] [steps]
# "base" = the product build, any other name = minimmerflow_b200/lib_<name> (make OUT=../lib_<name> EXTRA=...)
set -u
PAIRS=$1; ROUNDS=${2:-3}; STEPS=${3:-20}
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
for r in $(seq $ROUNDS); do
  for pair in $PAIRS; do
    name=${pair%%:*}; cfg=${pair#*:}
    if [ "$name" = base ]; then unset MMF_LIB_PATH; else export MMF_LIB_PATH=$PWD/minimmerflow_b200/lib_$name/libmmf_b200.so; fi
    timeout 120 python tools/stage_sweep.py --size ${SIZE:-256} --steps $STEPS --one "$cfg" 2>/dev/null | grep '^{' | sed "s/^{/{\"lib\": \"$name\", \"round\": $r, /" >> gpurun_out/ab.jsonl
  done
done
python - <<'P'
import json, collections, statistics
rows=collections.OrderedDict()
for l in open('gpurun_out/ab.jsonl'):
    d=json.loads(l); rows.setdefault((d['lib'], d['variant']), []).append(d)
for (lib, var), ds in rows.items():
    ms=[d['ms_per_step'] for d in ds]
    st=[statistics.median(d['stage_ms'][i] for d in ds) for i in range(3)]
    print(f"{lib:12s} {var:14s} median {statistics.median(ms):.4f} best {min(ms):.4f} runs {[round(m,3) for m in ms]} stages {[round(x,4) for x in st]} sha {ds[0]['state_sha256'][:8]}")
P
