#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python scripts/tune_amg.py 9 \
  '{"MK_AMG_DEG_COARSE":1}' \
  '{"MK_AMG_DEG_COARSE":3}' \
  '{"MK_AMG_DEG":2}' \
  '{"MK_AMG_DEG":2,"MK_AMG_DEG_COARSE":3}' \
  '{"MK_AMG_RATIO":4}' \
  '{"MK_AMG_RATIO":6}' \
  '{"MK_AMG_RATIO":12}' \
  '{"MK_AMG_RATIO":16}' \
  '{"MK_AMG_OMEGA":1.0}' \
  '{"MK_AMG_OMEGA":1.2}' \
  '{"MK_AMG_OMEGA":1.5}' \
  '{"MK_AMG_THETA":0.05}' \
  '{"MK_AMG_THETA":0.15}' \
  '{"MK_AMG_THETA":0.2}' \
  '{"MK_AMG_COARSE":1000}' \
  '{"MK_AMG_COARSE":100}' \
  '{"MK_PRECOND_FP32":0}' \
  '{"MK_AMG_TYPED":0}' \
  '{"MK_AMG_INJECT":0}' \
  > gpurun_out/r02p_tune.log 2>&1
cat gpurun_out/r02p_tune.log | cut -c1-260
