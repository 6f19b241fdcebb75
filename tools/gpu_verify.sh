#!/bin/bash
# the round-end checks in one GPU call: the whole `-m gpu` suite, smoke(), the default bench line
set -u
mkdir -p gpurun_out
TAG=${TAG:-verify}
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit code: $?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit code: $?"
python - <<P
import json
d = json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity'))
print('stage_ms', r.get('stage_ms'), 'frac', r['frac'], 'whole', r['whole_step']['frac'], 'clocks', d.get('clocks'))
print('repeats', d.get('repeats_ms_per_step') or d.get('repeats'))
P
