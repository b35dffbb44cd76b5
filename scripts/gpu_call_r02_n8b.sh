#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for l1 in 1 0; do
MK_DIST_L1=$l1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$l1 scripts/bench_partitioned.py --kind 4 --levels 11 --reps 50 --outer 3 > gpurun_out/r02n8b_part_n${N}_l1_$l1.json 2> gpurun_out/r02n8b_part_n${N}_l1_$l1.err
tail -2 gpurun_out/r02n8b_part_n${N}_l1_$l1.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n8b_part_n${N}_l1_$l1.json').read().strip().splitlines()[-1])
    print('dist_l1=$l1', d['n_gpus'], 'spmv', round(d['spmv_ms_overlap']*1e3,1), 'pcg', d['pcg_precond1'], 'resident', d.get('pcg_precond1_resident'), d.get('pcg_precond1_resident_error'), 'invit', d.get('inverse_iteration'))
except Exception as e:
    print('no line', e)
PY
done
