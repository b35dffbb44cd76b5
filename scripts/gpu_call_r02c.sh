#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MK_TAIL_ROWS=0 timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02c_probe_notail.json 2> gpurun_out/r02c_probe_notail.err
timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02c_probe_tail.json 2> gpurun_out/r02c_probe_tail.err
MK_TAIL_ROWS=16384 timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02c_probe_tail16k.json 2> gpurun_out/r02c_probe_tail16k.err
MK_TAIL_CLUSTER=8 timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02c_probe_tail_c8.json 2> gpurun_out/r02c_probe_tail_c8.err
cat gpurun_out/r02c_probe_*.json | cut -c1-1500
tail -3 gpurun_out/r02c_probe_tail.err
