#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_field.py tests/test_gpu_large.py -q --tb=short 2>&1 | tail -15) > gpurun_out/r02w_tests.log
tail -8 gpurun_out/r02w_tests.log
timeout 900 python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
PY
MK_RECYCLE_FUSED=0 timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r02w_bench_unfused.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench_unfused.json').read().strip().splitlines()[-1])
print('unfused', d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
PY
