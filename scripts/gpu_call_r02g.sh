#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02g_tests.log
timeout 900 python bench.py --no-cpu > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -5 gpurun_out/r02g_tests.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
print(d['e2e_stages_ms'][-1])
PY
