#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --tb=short 2>&1 | tail -25) > gpurun_out/r02n2b_dist_tests.log
tail -6 gpurun_out/r02n2b_dist_tests.log
for fp in 1 0; do
MK_PRECOND_FP32=$fp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$fp scripts/bench_partitioned.py --kind 4 --levels 11 --reps 50 --outer 3 > gpurun_out/r02n2b_part_fp32_$fp.json 2> gpurun_out/r02n2b_part_fp32_$fp.err
tail -2 gpurun_out/r02n2b_part_fp32_$fp.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n2b_part_fp32_$fp.json').read().strip().splitlines()[-1])
    print('fp32=$fp', d['n_gpus'], 'spmv', d['spmv_ms_overlap'], 'pcg', d['pcg_precond1'], 'invit', d.get('inverse_iteration'))
except Exception as e:
    print('no line', e)
PY
done
