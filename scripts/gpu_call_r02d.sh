#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r02d_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
MK_VERBOSE=1 timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02d_probe.json 2> gpurun_out/r02d_probe.err
tail -8 gpurun_out/r02d_tests.log
python -c "
import json
d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])"
grep "inverse iteration" gpurun_out/r02d_probe.err | awk '{s+=\$(NF-1); n+=\$5} END {print \"field iteration batches ms\", s, \"iterations\", n}'
grep "igm solve" gpurun_out/r02d_probe.err | awk '{s+=\$(NF-1)} END {print \"igm iteration batches ms\", s}'
