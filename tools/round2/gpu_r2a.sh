#!/usr/bin/env bash
# round 2, first GPU visit: tests, full bench line, single-package timing, launch list
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/r2a_gpus.txt; df -h /dev/shm >> $out/r2a_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2a.log
tail -15 $out/pytest_gpu_r2a.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_r2a.json 2> $out/bench_r2a.err; echo "bench rc=$?"
tail -c 1500 $out/bench_r2a.err; head -c 6000 $out/bench_r2a.json
NPK=1 WITH_E=0 timeout 120 python tools/gpu_time.py > $out/single_r2a.log 2>&1; NPK=115 WITH_E=0 timeout 120 python tools/gpu_time.py >> $out/single_r2a.log 2>&1; cat $out/single_r2a.log
SHORT="python bench.py --steps 2 --warmup 1 --packages 16 --no-cpu-baseline --no-config-e"
timeout 300 $SHORT > $out/plain_r2a.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_r2a.csv $SHORT > $out/ncu_launches_r2a.log 2>&1
echo "ncu launches rc=$?"
