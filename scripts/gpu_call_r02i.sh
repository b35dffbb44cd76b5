#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(MK_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_barrier.py -q --tb=short 2>&1 | grep -E "LM direction|FAILED|passed|failed|Error|assert|^E " | tail -40) > gpurun_out/r02i_barrier.log
tail -30 gpurun_out/r02i_barrier.log
MK_VERBOSE=1 timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02i_probe.json 2> gpurun_out/r02i_probe.err
grep "inverse iteration" gpurun_out/r02i_probe.err | head -50 | awk '{printf "%s/%s/%s ", $5, $9, $(NF-1)}'; echo
grep "inverse iteration" gpurun_out/r02i_probe.err | head -50 | awk '{s+=$(NF-1); n+=$5} END {print "field iteration batches ms", s, "iterations", n}'
python -c "
import json; d=json.load(open('gpurun_out/r02i_probe.json')); print(d['field']['solve_ms'], d['field']['inner'], d['igm']['solve_ms'], d['igm']['inner'])"
