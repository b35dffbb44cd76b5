#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_setup.py tests/test_gpu_field.py tests/test_gpu_igm.py -q --tb=short -x 2>&1 | tail -30) > gpurun_out/r02o_tests.log
tail -15 gpurun_out/r02o_tests.log
timeout 900 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
PY
MK_AMG_WEIGHTED=0 timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r02o_bench_unweighted.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench_unweighted.json').read().strip().splitlines()[-1])
print('unweighted', d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
PY
