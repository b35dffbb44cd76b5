#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(MK_VERBOSE=1 timeout 600 python -m pytest tests/test_gpu_large.py -q --tb=short -k "octa7" 2>&1 | grep -E "candidate|slack|passed|failed|assert|^E " | tail -80) > gpurun_out/r02n_octa7.log
tail -60 gpurun_out/r02n_octa7.log
(timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r02n_tests.log
tail -25 gpurun_out/r02n_tests.log
timeout 600 python tests/tools/fullx_tol_probe.py octa9 1e-10 > gpurun_out/r02n_fullx.jsonl 2> gpurun_out/r02n_fullx.err
cut -c1-600 gpurun_out/r02n_fullx.jsonl
