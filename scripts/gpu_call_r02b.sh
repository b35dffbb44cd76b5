#!/bin/bash
# round 2, call b: tail kernel reshaped for memory-level parallelism; large goldens
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_field.py tests/test_gpu_igm.py tests/test_gpu_large.py -x -q 2>&1 | tail -40) > gpurun_out/r02b_tests.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
MK_TAIL_ROWS=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02b_bench_notail.json 2> gpurun_out/r02b_bench_notail.err
MK_TAIL_ROWS=16384 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02b_bench_tail16k.json 2> gpurun_out/r02b_bench_tail16k.err
tail -5 gpurun_out/r02b_tests.log
for f in gpurun_out/r02b_bench*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['stages']['field_solve_ms'], d['stages']['field_inner'], d['stages']['igm_solve_ms'], d['stages']['igm_inner'])"; done
