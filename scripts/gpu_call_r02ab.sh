#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PART_TRACE=gpurun_out/r02ab_part_trace_1gpu_2M.txt timeout 600 python scripts/bench_partitioned.py --kind 4 --levels 10 --reps 20 > gpurun_out/r02ab_part.json 2> gpurun_out/r02ab_part.err
tail -2 gpurun_out/r02ab_part.err | cut -c1-300
cat gpurun_out/r02ab_part_trace_1gpu_2M.txt
cut -c1-900 gpurun_out/r02ab_part.json
