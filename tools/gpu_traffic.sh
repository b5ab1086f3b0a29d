#!/usr/bin/env bash
# DRAM traffic of the EGA kernel at the bench size (115 packages): one ncu --set full capture
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config-e"
$CMD > gpurun_out/plain_traffic.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ega_fast -s 1 -c 1 -o gpurun_out/prof_ega_benchsize -f $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "ncu rc=$?"
