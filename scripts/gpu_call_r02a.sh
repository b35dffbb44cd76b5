#!/bin/bash
# round 2, call a: first run of the restructured V-cycle (fused smoother epilogues, cluster tail kernel, fused PCG body)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
(timeout 1200 python -m pytest tests/test_gpu_field.py tests/test_gpu_igm.py tests/test_gpu_setup.py -x -q 2>&1 | tail -60) > gpurun_out/r02a_tests.log
(MK_TAIL_ROWS=0 timeout 600 python -m pytest tests/test_gpu_field.py -x -q 2>&1 | tail -30) > gpurun_out/r02a_tests_notail.log
MK_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
MK_TAIL_ROWS=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02a_bench_notail.json 2> gpurun_out/r02a_bench_notail.err
tail -5 gpurun_out/r02a_tests.log; cat gpurun_out/r02a_bench.json | cut -c1-600; tail -3 gpurun_out/r02a_bench.err
