#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short --durations=8 2>&1 | tail -24) > gpurun_out/r02fe_tests.log
tail -16 gpurun_out/r02fe_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02fe_bench.json 2> gpurun_out/r02fe_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02fe_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'], d['gpu_launches'], d['roofline']['frac'])
print({k: d['roofline']['iteration'][k].get('frac') for k in d['roofline'].get('iteration', {})})
PY
