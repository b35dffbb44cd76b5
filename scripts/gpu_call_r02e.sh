#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r02e_tests.log
timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/r02e_ref.json 2> gpurun_out/r02e_ref.err
timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02e_probe.json 2> gpurun_out/r02e_probe.err
tail -8 gpurun_out/r02e_tests.log
tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'], d['stages'])
print(d['roofline'].get('iteration'))
print(d.get('reference_sample'), d.get('cpu_baseline'))
print(d['e2e_stages_ms'][-1])
PY
cut -c1-600 gpurun_out/r02e_ref.json
cut -c1-900 gpurun_out/r02e_probe.json
