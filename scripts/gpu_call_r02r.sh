#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MK_VERBOSE=2 timeout 300 python scripts/recycle_probe.py one > gpurun_out/r02r_one.out 2> gpurun_out/r02r_one.err
MK_VERBOSE=2 timeout 300 python scripts/recycle_probe.py two > gpurun_out/r02r_two.out 2> gpurun_out/r02r_two.err
tail -1 gpurun_out/r02r_one.out | cut -c1-300; tail -1 gpurun_out/r02r_two.out | cut -c1-300
sed -n '/=== timed/,$p' gpurun_out/r02r_two.err | grep -E "inverse iteration|recycle slot" | tail -100
