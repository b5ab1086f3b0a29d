#!/usr/bin/env bash
# ncu --set full capture of the specialised EGA kernel on Config E (8 packages): tools/gpu_prof_e.sh <tag>
tag=${1:-e}
SHORT="python bench.py --config e --steps 2 --warmup 1 --packages 8 --no-cpu-baseline --no-config-e"
$SHORT > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ega_fast -s 1 -c 1 -o gpurun_out/prof_ega_$tag -f $SHORT > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu rc=$?"; tail -c 600 gpurun_out/plain_$tag.log
