#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python tests/tools/fullx_tol_probe.py octa9 1e-10 > gpurun_out/r02m_fullx_tol2.jsonl 2> gpurun_out/r02m_fullx_tol.err
cat gpurun_out/r02m_fullx_tol2.jsonl; tail -5 gpurun_out/r02m_fullx_tol.err
timeout 600 python tests/tools/fullx_tol_probe.py octa8 1e-10 >> gpurun_out/r02m_fullx_tol2.jsonl 2>> gpurun_out/r02m_fullx_tol.err; cat gpurun_out/r02m_fullx_tol2.jsonl
