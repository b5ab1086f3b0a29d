#!/usr/bin/env bash
# round 2, 1-GPU visit: latency mode with the continuum/source kernel on a side stream; full tests
set -u
out=gpurun_out; mkdir -p $out
{
echo "== single package (side stream)"; timeout 120 python tools/gpu_single.py
echo "== single package (no side stream)"; JRB_NO_SIDE_STREAM=1 timeout 120 python tools/gpu_single.py
} > $out/variants_r2q.log 2>&1
grep -E "^==|^\[|Error" $out/variants_r2q.log
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2q.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu_r2q.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $out/launches_single_r2q.csv python tools/gpu_single.py > $out/ncu_single_r2q.log 2>&1
grep -E "ega_|ray_step|los_fin|stage_k|tail_sort" $out/launches_single_r2q.csv | tail -9 | cut -d'"' -f10,28-30
