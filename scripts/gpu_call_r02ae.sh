#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_large.py -q --tb=short 2>&1 | tail -5)
timeout 900 python bench.py > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ae_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
PY
