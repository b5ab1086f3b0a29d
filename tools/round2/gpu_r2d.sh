#!/usr/bin/env bash
# round 2, 4th GPU visit (1 GPU): full tests, single-package breakdown, PERCH throughput, ncu captures
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2d.log
tail -12 $out/pytest_gpu_r2d.log
{
echo "== single package";             timeout 120 python tools/gpu_single.py
echo "== single package, no coop tracer"; JRB_NO_COOP_TRACER=1 timeout 120 python tools/gpu_single.py
echo "== jitter D 115 (per-channel axes)"; JITTER=1 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== jitter D 115, generic kernel";    JITTER=1 GENERIC=1 NPK=16 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== 30 gases x 100 channels, 66-ray packages x 64"; WITH_R=1 NPK=1 WITH_E=0 timeout 600 python tools/gpu_time.py
} > $out/variants_r2d.log 2>&1
grep -E "^==|^\[" $out/variants_r2d.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/launches_single_r2d.csv python tools/gpu_single.py > $out/ncu_single_r2d.log 2>&1
echo "ncu single rc=$?"; grep -E "ega_|ray_step|los_fin|stage_k" $out/launches_single_r2d.csv | tail -14 | cut -d'"' -f10,28-30
# full-size EGA launch: DRAM traffic + secondary pipe numbers for roofline.traffic
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config-e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ega_fast_kernel -s 2 -c 1 -o $out/prof_ega_full_r2d -f $CMD > $out/ncu_full_r2d.log 2>&1
echo "ncu full rc=$?"; ls -la $out/*.ncu-rep 2>/dev/null | tail -3
