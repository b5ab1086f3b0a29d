#!/usr/bin/env bash
# round 2: stepping kernels write compact raw points (64 B per step), records are 112 B without the tail -- full validation,
# default bench, and DRAM traffic + pipe numbers of the main kernel with the new record size (metrics subset, no --set full)
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r2v.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -6 $out/pytest_gpu_r2v.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2v.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_r2v.log
timeout 900 python bench.py > $out/bench_r2v.json 2> $out/bench_r2v.err; rc=$?; echo "bench rc=$rc"; tail -c 300 $out/bench_r2v.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2v.json'))
print("value", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "tracer", d["roofline"]["raytrace_ms_per_step"], "ega", d["roofline"]["kernel_ms"], "parity", d["parity"]["ok"], "traffic", d["roofline"]["traffic"])
e=d["extra"]["config_e"]; print("  E", e["value"], e["e2e"]["value"], e["roofline"]["frac"], e["parity"]["ok"])
print("  single", d["extra"]["single_package"]["ms_device"], d["extra"]["single_package"]["ms_wall_per_call"])
PY
[ $rc -ne 0 ] && exit 1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
timeout 500 ncu --metrics $M --clock-control none -k regex:ega_tiled_kernel -c 1 --csv --log-file $out/ncu_traffic_r2v.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config-e > $out/ncu_traffic_r2v.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $out/ncu_traffic_r2v.csv | cut -d'"' -f10,26,28,30 | tail -16
