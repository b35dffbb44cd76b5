#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tests/tools/b2_penalty_probe.py torus_24x16 6 > gpurun_out/r02t_b2.jsonl 2> gpurun_out/r02t_b2.err
timeout 300 python tests/tools/b2_penalty_probe.py sphere_ico3 4 >> gpurun_out/r02t_b2.jsonl 2>> gpurun_out/r02t_b2.err
cut -c1-700 gpurun_out/r02t_b2.jsonl; tail -3 gpurun_out/r02t_b2.err
(timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r02t_tests.log
tail -12 gpurun_out/r02t_tests.log
timeout 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'])
print(d['roofline'].get('iteration',{}).get('field',{}).get('frac'), d['roofline'].get('iteration',{}).get('igm',{}).get('frac'))
PY
timeout 900 python scripts/bench_partitioned.py --kind 4 --levels 11 --reps 50 --outer 3 > gpurun_out/r02t_part_1gpu_8M.json 2> gpurun_out/r02t_part_1gpu_8M.err
tail -2 gpurun_out/r02t_part_1gpu_8M.err | cut -c1-300; cut -c1-1500 gpurun_out/r02t_part_1gpu_8M.json
