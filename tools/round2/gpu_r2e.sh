#!/usr/bin/env bash
# round 2, 5th GPU visit (1 GPU): tiled kernel A/B
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_split_mode.py -m gpu -q -k "tiled or cooperative" > $out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2e.log
tail -12 $out/pytest_gpu_r2e.log
{
echo "== fused 115";  JRB_EGA_TILED=0 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== tiled 115";  JRB_EGA_TILED=1 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== tiled 115, 512 threads";  JRB_EGA_THREADS=512 JRB_EGA_TILED=1 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== tiled 115, 640 threads";  JRB_EGA_THREADS=640 JRB_EGA_TILED=1 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
} > $out/variants_r2e.log 2>&1
grep -E "^==|^\[" $out/variants_r2e.log
SHORT="python tools/gpu_time.py"
JRB_EGA_TILED=1 NPK=115 WITH_E=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ega_tiled -s 1 -c 1 -o $out/prof_ega_tiled_r2e -f $SHORT > $out/ncu_tiled_r2e.log 2>&1
echo "ncu rc=$?"
