#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MK_VERBOSE=1 timeout 900 python tests/tools/fullx_tol_probe.py octa9 1e-10 1e-11 3e-12 > gpurun_out/r02u_fullx.jsonl 2> gpurun_out/r02u_fullx.err
grep -E "candidate|slack update" gpurun_out/r02u_fullx.err | tail -60
cut -c1-500 gpurun_out/r02u_fullx.jsonl
