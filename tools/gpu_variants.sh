#!/bin/bash
# times experiment builds of the library (make OUT=../lib_<name> EXTRA=...) next to the product build:
#   bash tools/gpu_variants.sh "<variants for stage_sweep>" name1 name2 ...
set -u
VARS=$1; shift
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for name in base "$@"; do
  if [ "$name" = base ]; then unset MMF_LIB_PATH; else export MMF_LIB_PATH=$PWD/minimmerflow_b200/lib_$name/libmmf_b200.so; fi
  timeout 300 python tools/stage_sweep.py --size 256 --steps 6 --variants "$VARS" --variant-timeout 60 2>/dev/null | grep -v summary | sed "s/^{/{\"lib\": \"$name\", /" >> gpurun_out/variants.jsonl
done
python - <<'P'
import json
for l in open('gpurun_out/variants.jsonl'):
    d=json.loads(l)
    print(f"{d['lib']:10s} {d['variant']:18s}", round(d.get('ms_per_step',0),4), [round(x,4) for x in d.get('stage_ms',[])], d.get('state_sha256','')[:8], d.get('error','')[-200:])
P
