#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MK_VERBOSE=1 timeout 900 python tests/tools/fullx_tol_probe.py octa9 1e-10 1e-11 > gpurun_out/r02v_fullx.jsonl 2> gpurun_out/r02v_fullx.err
grep -E "first-order" gpurun_out/r02v_fullx.err | tail; cut -c1-420 gpurun_out/r02v_fullx.jsonl | grep -v poisson
(timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r02v_tests.log
tail -12 gpurun_out/r02v_tests.log
