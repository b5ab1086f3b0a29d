#!/usr/bin/env bash
# round 2, 1-GPU visit: tile 6 + 32-channel groups for Config E; tests, bench, full-size ncu capture of the final kernel
set -u
out=gpurun_out; mkdir -p $out
{
echo "== D 115 / E 58 (final defaults)"; NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== 30 gases"; NPK=1 WITH_E=0 WITH_R=1 timeout 600 python tools/gpu_time.py
} > $out/variants_r2m.log 2>&1
grep -E "^==|^\[|Error" $out/variants_r2m.log
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2m.log 2>&1; echo "pytest all rc=$?"; tail -4 $out/pytest_gpu_r2m.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_r2m.json 2> $out/bench_r2m.err; echo "bench rc=$?"
tail -c 400 $out/bench_r2m.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2m.json'))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "kernel", d["roofline"]["kernel"], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "parity", d["parity"]["ok"])
print(d["extra"]["single_package"]); e=d["extra"]["config_e"]; print("E", e["value"], e["e2e"]["value"], e["roofline"]["frac"], e["roofline"]["kernel"], e["parity"]["ok"])
PY
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config-e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ega_tiled_kernel -s 2 -c 1 -o $out/prof_ega_tiled_full_r2m -f $CMD > $out/ncu_full_r2m.log 2>&1
echo "ncu full rc=$?"
