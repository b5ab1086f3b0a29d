#!/usr/bin/env bash
# round 2, second GPU visit: tests, full bench line, A/B variants, single-package launch list
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2b.log
tail -25 $out/pytest_gpu_r2b.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_r2b.json 2> $out/bench_r2b.err; echo "bench rc=$?"
tail -c 1200 $out/bench_r2b.err; head -c 7000 $out/bench_r2b.json; echo
{
echo "== baseline 115";            NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== no host rows";            JRB_NO_HOST_ROWS=1 NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== LOS evict first";         JRB_LOS_EVICT_FIRST=1 NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== L2 persist";              JRB_L2_PERSIST=1 NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== L2 persist + evict first"; JRB_L2_PERSIST=1 JRB_LOS_EVICT_FIRST=1 NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== 30 gases: split (default)";  NPK=1 WITH_E=0 WITH_C=1 NPK_C=8 timeout 300 python tools/gpu_time.py
echo "== 30 gases: fused";            JRB_NO_SPLIT=1 NPK=1 WITH_E=0 WITH_C=1 NPK_C=8 timeout 300 python tools/gpu_time.py
echo "== single package";             timeout 120 python tools/gpu_single.py
echo "== single package, fused";      JRB_NO_SPLIT=1 timeout 120 python tools/gpu_single.py
} > $out/variants_r2b.log 2>&1
grep -E "^==|^\[" $out/variants_r2b.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_single_r2b.csv python tools/gpu_single.py > $out/ncu_single_r2b.log 2>&1
echo "ncu single rc=$?"; tail -40 $out/launches_single_r2b.csv | cut -c1-220
