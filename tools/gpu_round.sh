#!/usr/bin/env bash
# One GPU-box visit: parity tests, bench line, ncu launch list and one full capture of the EGA kernel.
# usage: tools/gpu_round.sh <tag> [skip-tests]
set -u
tag=${1:-r}
out=gpurun_out
mkdir -p $out
if [ "${2:-}" != "skip-tests" ]; then
  python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
  tail -5 $out/pytest_gpu_$tag.log
fi
python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
tail -c 3000 $out/bench_$tag.json
SHORT="python bench.py --steps 2 --warmup 1 --packages 16 --no-cpu-baseline --no-config-e"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $out/launches_$tag.csv $SHORT > $out/ncu_launches_$tag.log 2>&1
echo "ncu launches rc=$?"
$SHORT > $out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ega_fast -s 1 -c 1 -o $out/prof_ega_$tag -f $SHORT > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"
ls -la $out | tail -12
