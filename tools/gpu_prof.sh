#!/usr/bin/env bash
# one ncu --set full capture of the specialised EGA kernel (Config D, 16 packages): tools/gpu_prof.sh <tag>
tag=${1:-p}
SHORT="python bench.py --steps 2 --warmup 1 --packages 16 --no-cpu-baseline --no-config-e"
$SHORT > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ega_fast -s 1 -c 1 -o gpurun_out/prof_ega_$tag -f $SHORT > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu rc=$?"
