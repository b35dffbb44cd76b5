#!/bin/bash
# ncu --set full on the kernels of an IGM PCG iteration (the field iteration is in ncu_full_r02_raw.csv); exported to CSV on the box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/profile_kernels.py prep /tmp/wl.npz 9 > gpurun_out/r02ff_prep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_sell<float|k_spmv_sell<double|k_update_xr_handover|k_cprec_finish|k_update_p<|k_project" --launch-skip 400 --launch-count 44 -f -o /tmp/ncu_r02_igm python scripts/profile_kernels.py run /tmp/wl.npz igm,round > gpurun_out/r02ff_ncu_igm.log 2>&1
ls -la /tmp/ncu_r02_igm.ncu-rep
ncu -i /tmp/ncu_r02_igm.ncu-rep --page raw --csv > gpurun_out/ncu_full_r02_igm_raw.csv 2> gpurun_out/r02ff_export.err
ls -la gpurun_out/ncu_full_r02_igm_raw.csv; du -sh gpurun_out
