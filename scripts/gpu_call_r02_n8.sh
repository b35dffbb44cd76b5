#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02n8_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02n8_bench.json 2> gpurun_out/r02n8_bench.err
tail -3 gpurun_out/r02n8_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02n8_bench.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
    print(json.dumps(d.get('partitioned'))[:1800])
except Exception as e:
    print("no bench line", e)
PY
