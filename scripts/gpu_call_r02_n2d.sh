#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(MK_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --tb=short 2>&1 | grep -E "distributed level 1|passed|failed|Error|assert|^E " | tail -14) > gpurun_out/r02n2d_dist_tests.log
tail -12 gpurun_out/r02n2d_dist_tests.log
