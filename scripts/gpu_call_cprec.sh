#!/bin/bash
mkdir -p gpurun_out
MK_VERBOSE=1 timeout 68 python scripts/cprec_probe.py 5 9 > gpurun_out/cprec_probe.log 2> gpurun_out/cprec_probe.err
grep -v "^\[mk\]" gpurun_out/cprec_probe.log | tail -8
grep "igm solve" gpurun_out/cprec_probe.err | tail -13 | cut -c1-90
