#!/usr/bin/env bash
# round 2, 6th GPU visit (1 GPU): tiled default, tile sizes, latency mode, full tests + bench
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2f.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2f.log
tail -8 $out/pytest_gpu_r2f.log
{
echo "== single package"; timeout 120 python tools/gpu_single.py
echo "== default (T=8) 115"; NPK=115 NPK_E=58 timeout 300 python tools/gpu_time.py
for t in 4 6 12 16; do echo "== T=$t 115"; JRB_LIBDIR=$PWD/build/lib_t$t NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py; done
echo "== default 32 pkgs"; NPK=32 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== fused 32 pkgs"; JRB_EGA_TILED=0 NPK=32 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== default 8 pkgs"; NPK=8 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== fused 8 pkgs"; JRB_EGA_TILED=0 NPK=8 WITH_E=0 timeout 300 python tools/gpu_time.py
} > $out/variants_r2f.log 2>&1
grep -E "^==|^\[" $out/variants_r2f.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_r2f.json 2> $out/bench_r2f.err; echo "bench rc=$?"
tail -c 600 $out/bench_r2f.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2f.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "kernel", d["roofline"]["kernel"], "kernel_ms", d["roofline"]["kernel_ms"], "parity", d["parity"]["ok"])
e=d["extra"]["config_e"]; print("E", e["value"], e["e2e"]["value"], e["roofline"]["frac"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/launches_single_r2f.csv python tools/gpu_single.py > $out/ncu_single_r2f.log 2>&1
grep -E "ega_|ray_step|los_fin|stage_k" $out/launches_single_r2f.csv | tail -8 | cut -d'"' -f10,28-30
