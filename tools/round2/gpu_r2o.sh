#!/usr/bin/env bash
# round 2, 1-GPU visit: ncu --set full of the 30-gas pass kernel (split mode) vs the fused kernel, and of Config E on the tiled kernel
set -u
out=gpurun_out; mkdir -p $out
R="python tools/gpu_time.py"
NPK=1 WITH_E=0 WITH_R=1 timeout 600 ncu --set full --clock-control none -k regex:ega_fast_kernel -s 1 -c 1 -o $out/prof_refspec_split_r2o -f $R > $out/ncu_refspec_split_r2o.log 2>&1; echo "ncu split rc=$?"
JRB_NO_SPLIT=1 NPK=1 WITH_E=0 WITH_R=1 timeout 600 ncu --set full --clock-control none -k regex:ega_fast_kernel -s 1 -c 1 -o $out/prof_refspec_fused_r2o -f $R > $out/ncu_refspec_fused_r2o.log 2>&1; echo "ncu fused rc=$?"
NPK=1 NPK_E=58 timeout 600 ncu --set full --clock-control none -k regex:ega_tiled_kernel -s 4 -c 1 -o $out/prof_e_tiled_r2o -f $R > $out/ncu_e_tiled_r2o.log 2>&1; echo "ncu E rc=$?"
ls -la $out/*r2o.ncu-rep
