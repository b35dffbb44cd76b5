#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python scripts/bench_partitioned.py --kind 4 --levels 11 --reps 50 --outer 3 > gpurun_out/r02ac_part_1gpu.json 2> gpurun_out/r02ac_part_1gpu.err
tail -2 gpurun_out/r02ac_part_1gpu.err | cut -c1-200; cut -c1-1500 gpurun_out/r02ac_part_1gpu.json
