#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_setup.py tests/test_gpu_field.py tests/test_gpu_igm.py tests/test_gpu_integ.py tests/test_gpu_large.py -q --tb=short -x 2>&1 | tail -8)
MK_VERBOSE=1 timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02af_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
PY
grep -E "build_amg_device: (P = S T|A P|R \(A P\)|R = P)" gpurun_out/r02af_bench.err | head -12
