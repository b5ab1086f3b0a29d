#!/usr/bin/env bash
# round 2, 1-GPU visit: latency mode with all rays longest first
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_split_mode.py tests/test_gpu_io_and_lanes.py tests/test_gpu_dropin.py -m gpu -q > $out/pytest_gpu_r2p.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu_r2p.log
{
echo "== single package (sorted)"; timeout 120 python tools/gpu_single.py
echo "== single package (not sorted)"; JRB_NO_TAIL_SORT=1 timeout 120 python tools/gpu_single.py
echo "== 4 packages (sorted)"; NPK=4 WITH_E=0 timeout 120 python tools/gpu_time.py
echo "== 4 packages (not sorted)"; JRB_NO_TAIL_SORT=1 NPK=4 WITH_E=0 timeout 120 python tools/gpu_time.py
} > $out/variants_r2p.log 2>&1
grep -E "^==|^\[|Error" $out/variants_r2p.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $out/launches_single_r2p.csv python tools/gpu_single.py > $out/ncu_single_r2p.log 2>&1
grep -E "ega_|ray_step|los_fin|stage_k|tail_sort" $out/launches_single_r2p.csv | tail -9 | cut -d'"' -f10,28-30
