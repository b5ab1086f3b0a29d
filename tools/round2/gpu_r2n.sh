#!/usr/bin/env bash
# round 2, 1-GPU visit: final build -- tests, smoke, refspec shape timing, launch list of the default bench command
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2n.log 2>&1; echo "pytest all rc=$?"; tail -4 $out/pytest_gpu_r2n.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2n.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke_r2n.log
{
echo "== 30 gases x 100 channels (final)"; NPK=1 WITH_E=0 WITH_R=1 timeout 600 python tools/gpu_time.py
echo "== jitter D 115 (per-channel axes, final)"; JITTER=1 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
} > $out/variants_r2n.log 2>&1
grep -E "^==|^\[|Error" $out/variants_r2n.log
SHORT="python bench.py --steps 2 --warmup 1 --packages 115 --no-cpu-baseline --no-config-e"
timeout 300 $SHORT > $out/plain_r2n.json 2> $out/plain_r2n.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_r2n.csv $SHORT > $out/ncu_launches_r2n.log 2>&1
echo "ncu launches rc=$?"; grep -E "ega_|ray_step|los_fin|stage_k|tail_sort" $out/launches_r2n.csv | head -8 | cut -d'"' -f10,28-30
