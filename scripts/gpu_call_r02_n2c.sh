#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(MK_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --tb=short 2>&1 | grep -E "distributed level 1|passed|failed|Error|assert" | tail -12) > gpurun_out/r02n2c_dist_tests.log
tail -8 gpurun_out/r02n2c_dist_tests.log
for l1 in 1 0; do
MK_DIST_L1=$l1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$l1 scripts/bench_partitioned.py --kind 4 --levels 11 --reps 50 --outer 3 > gpurun_out/r02n2c_part_l1_$l1.json 2> gpurun_out/r02n2c_part_l1_$l1.err
tail -2 gpurun_out/r02n2c_part_l1_$l1.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n2c_part_l1_$l1.json').read().strip().splitlines()[-1])
    print('dist_l1=$l1', d['n_gpus'], 'pcg', d['pcg_precond1'], 'resident', d.get('pcg_precond1_resident'), d.get('pcg_precond1_resident_error'), 'invit', d.get('inverse_iteration'))
except Exception as e:
    print('no line', e)
PY
done
