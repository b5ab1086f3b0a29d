#!/usr/bin/env bash
# ncu --set full capture of the two tracer kernels (Config D, 32 packages): tools/gpu_prof_rt.sh <tag>
tag=${1:-p}
SHORT="python bench.py --steps 2 --warmup 1 --packages 32 --no-cpu-baseline --no-config-e"
$SHORT > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ray_step|los_finalize" -s 2 -c 2 -o gpurun_out/prof_rt_$tag -f $SHORT > gpurun_out/ncu_rt_$tag.log 2>&1
echo "ncu rc=$?"
