#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests/test_gpu_igm.py tests/test_gpu_large.py tests/test_gpu_barrier.py tests/test_gpu_integ.py tests/test_shim.py -m gpu -q --tb=short 2>&1 | tail -15) > gpurun_out/r02x_tests.log
tail -8 gpurun_out/r02x_tests.log
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'], d['gpu_launches'])
PY
