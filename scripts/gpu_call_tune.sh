#!/bin/bash
# One gpurun call: multigrid / precision variants of the 1M-vertex step on one workload (scripts/tune_amg.py).
mkdir -p gpurun_out
timeout 280 python scripts/tune_amg.py 9 \
  '{"MK_AMG_TYPED":1}' \
  '{"MK_AMG_TYPED":1,"MK_FP64_TAIL":-1}' \
  '{"MK_AMG_TYPED":1,"MK_FP64_TAIL":-2}' \
  '{"MK_AMG_TYPED":1,"MK_FP64_TAIL":2}' \
  '{"MK_AMG_TYPED":1,"MK_PRECOND_FP32":0}' \
  '{"MK_AMG_TYPED":1,"MK_AMG_COARSE":500}' \
  '{"MK_AMG_TYPED":1,"MK_AMG_COARSE":500,"MK_FP64_TAIL":-1}' \
  '{"MK_FP64_TAIL":-1}' \
  '{"MK_PRECOND_FP32":0}' \
  > gpurun_out/tune_tail.log 2>&1
tail -12 gpurun_out/tune_tail.log
