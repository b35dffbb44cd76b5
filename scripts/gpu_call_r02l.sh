#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r02l_tests.log
tail -25 gpurun_out/r02l_tests.log
