#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(MK_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_barrier.py -x -q 2>&1 | grep -v "inverse iteration\|igm solve\|mis2\|build_amg\|igm_create\|layout" | tail -80) > gpurun_out/r02h_barrier.log
tail -40 gpurun_out/r02h_barrier.log
(timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_barrier.py 2>&1 | tail -15) > gpurun_out/r02h_tests.log
tail -5 gpurun_out/r02h_tests.log
timeout 900 python bench.py --no-cpu > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
print(d['e2e_stages_ms'][-1])
PY
