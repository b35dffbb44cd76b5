#!/bin/bash
# One gpurun call: GPU tests, then the multigrid variants of the 1M-vertex step on one workload (scripts/tune_amg.py), then smoke().
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; tail -3 gpurun_out/pytest_gpu17.log
timeout 240 python scripts/tune_amg.py 9 '{"MK_AMG_TYPED":0}' '{"MK_AMG_DEG_COARSE":2}' '{"MK_AMG_DEG_COARSE":3}' '{"MK_AMG_TYPED":0,"MK_AMG_DEG_COARSE":2}' > gpurun_out/tune_typed.log 2>&1
cat gpurun_out/tune_typed.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
