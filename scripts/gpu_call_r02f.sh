#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/trace_iterations.py field 9 2 > gpurun_out/r02f_trace_field.log 2>&1
timeout 300 python scripts/trace_iterations.py igm 9 > gpurun_out/r02f_trace_igm.log 2>&1
tail -3 gpurun_out/r02f_trace_field.log gpurun_out/r02f_trace_igm.log
(timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_integ.py -x -q 2>&1 | tail -15) > gpurun_out/r02f_tests.log
tail -5 gpurun_out/r02f_tests.log
ls -la gpurun_out/trace_*.csv
