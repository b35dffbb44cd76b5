#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python scripts/trace_iterations.py field 9 50 > gpurun_out/r02ad_trace_field.log 2>&1
python scripts/trace_summary.py gpurun_out/trace_field.csv 40 > gpurun_out/r02ad_trace_field_summary.txt; cat gpurun_out/r02ad_trace_field_summary.txt
rm -f gpurun_out/trace_field.csv
