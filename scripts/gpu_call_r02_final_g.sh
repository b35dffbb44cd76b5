#!/bin/bash
# ncu --set full on the real-fp32 SELL SpMV of the IGM V-cycle after the move to 8 CTAs per SM (32 registers); Poisson solve only, 12 launches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python scripts/profile_kernels.py prep /tmp/wl.npz 9 > gpurun_out/r02fg_prep.log 2>&1
timeout 110 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_spmv_sell<float" --launch-skip 60 --launch-count 12 -f -o /tmp/ncu_r02_fg python scripts/profile_kernels.py run /tmp/wl.npz igm > gpurun_out/r02fg_ncu.log 2>&1
ncu -i /tmp/ncu_r02_fg.ncu-rep --page raw --csv > gpurun_out/ncu_full_r02_fg_sell_float_raw.csv 2> gpurun_out/r02fg_export.err
ls -la gpurun_out/ncu_full_r02_fg_sell_float_raw.csv
