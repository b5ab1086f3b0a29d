#!/usr/bin/env bash
# round 2, 7th GPU visit (1 GPU): final-code tests + smoke, full-size ncu capture of the tiled kernel, ncu launch list of the bench command
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2g.log
tail -6 $out/pytest_gpu_r2g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2g.log 2>&1; echo "smoke rc=$?"; tail -4 $out/smoke_r2g.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config-e"
timeout 600 $CMD > $out/plain_r2g.json 2> $out/plain_r2g.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ega_tiled_kernel -s 2 -c 1 -o $out/prof_ega_tiled_full_r2g -f $CMD > $out/ncu_full_r2g.log 2>&1
echo "ncu full rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --packages 115 --no-cpu-baseline --no-config-e"
timeout 300 $SHORT > $out/plain2_r2g.json 2> $out/plain2_r2g.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_r2g.csv $SHORT > $out/ncu_launches_r2g.log 2>&1
echo "ncu launches rc=$?"; tail -12 $out/launches_r2g.csv | cut -d'"' -f10,28-30
