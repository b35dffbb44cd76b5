#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python scripts/trace_iterations.py field 9 50 > gpurun_out/r02q_trace_field.log 2>&1
python scripts/trace_summary.py gpurun_out/trace_field.csv 40 > gpurun_out/r02q_trace_field_summary.txt; cat gpurun_out/r02q_trace_field_summary.txt
timeout 600 python scripts/trace_iterations.py igm_round 9 > gpurun_out/r02q_trace_igm.log 2>&1
mv gpurun_out/trace_igm_round.csv gpurun_out/trace_igm_round.csv 2>/dev/null
python scripts/trace_summary.py gpurun_out/trace_igm_round.csv 40 > gpurun_out/r02q_trace_igm_summary.txt; cat gpurun_out/r02q_trace_igm_summary.txt
MK_VERBOSE=1 timeout 600 python bench.py --steps 1 --warmup 1 2> gpurun_out/r02q_verbose.err > /dev/null
grep -E "igm_create|build_amg_device|field_amg|spgemm:" gpurun_out/r02q_verbose.err | tail -80
timeout 900 python scripts/tune_amg.py 9 \
  '{"MK_AMG_THETA_FIELD":0.02}' \
  '{"MK_AMG_THETA_FIELD":0.05}' \
  '{"MK_AMG_THETA_FIELD":0.08}' \
  '{"MK_AMG_THETA_FIELD":0.1}' \
  '{"MK_AMG_COARSE_FIELD":1000}' \
  '{"MK_AMG_COARSE_FIELD":1000,"MK_AMG_THETA_FIELD":0.05}' \
  '{"MK_AMG_COARSE_FIELD":1400,"MK_AMG_THETA_FIELD":0.05}' \
  '{"MK_AMG_COARSE_REAL":1400}' \
  '{"MK_AMG_THETA_REAL":0.08}' \
  '{"MK_AMG_THETA_REAL":0.12}' \
  '{"MK_AMG_THETA_FIELD":0.05,"MK_AMG_DEG_COARSE":3}' \
  '{"MK_AMG_THETA_FIELD":0.05,"MK_AMG_RATIO":6}' \
  > gpurun_out/r02q_tune.log 2>&1
cat gpurun_out/r02q_tune.log | cut -c1-260
