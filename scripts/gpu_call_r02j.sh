#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(MK_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_barrier.py -q --tb=short 2>&1 | grep -E "LM direction|FAILED|passed|failed|Error|assert|^E " | tail -40) > gpurun_out/r02j_barrier.log
tail -30 gpurun_out/r02j_barrier.log
timeout 400 python scripts/reorder_probe.py 9 > gpurun_out/r02j_reorder.json 2> gpurun_out/r02j_reorder.err
cat gpurun_out/r02j_reorder.json; tail -3 gpurun_out/r02j_reorder.err
