#!/bin/bash
# generic path: library builds (minimmerflow_b200/lib_<name>, "base" = the product build) round-robin on three meshes
#   bash tools/gpu_generic_ab.sh "base pf1_5 pf1_4" [rounds]
set -u
LIBS=$1; ROUNDS=${2:-2}
mkdir -p gpurun_out
O=gpurun_out/${TAG:-gen_ab}.jsonl
: > $O
for r in $(seq $ROUNDS); do
  for name in $LIBS; do
    if [ "$name" = base ]; then unset MMF_LIB_PATH; else export MMF_LIB_PATH=$PWD/minimmerflow_b200/lib_$name/libmmf_b200.so; fi
    for args in "--dim 2 --size 1024" "--two-level 48" "--size 128 --generic-only"; do
      timeout 120 python tools/generic_bench.py $args --steps 10 2>>gpurun_out/${TAG:-gen_ab}.err | grep '^{' | sed "s/^{/{\"lib\": \"$name\", \"round\": $r, /" >> $O
    done
  done
done
python - <<P
import json, collections, statistics
rows = collections.OrderedDict()
for l in open('$O'):
    d = json.loads(l)
    if 'ms_per_step' not in d: continue
    rows.setdefault((d['lib'], d.get('mesh') or ('3-D %d cells' % d['cells'])), []).append(d['ms_per_step'])
for (lib, mesh), ms in rows.items():
    print(f"{lib:8s} {mesh[:44]:44s} median {statistics.median(ms):.4f} runs {[round(m, 4) for m in ms]}")
P
tail -3 gpurun_out/${TAG:-gen_ab}.err
