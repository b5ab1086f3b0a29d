#!/usr/bin/env bash
# round 2, 1-GPU visit: tracer overlap (opt-in) A/B with the 23-warp CTA; final tests + bench with defaults
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_split_mode.py -m gpu -q -x -k "tiled" > $out/pytest_gpu_r2j.log 2>&1; echo "pytest tiled rc=$?" | tee -a $out/pytest_gpu_r2j.log
tail -4 $out/pytest_gpu_r2j.log
{
echo "== no overlap 115"; NPK=115 WITH_E=0 timeout 200 python tools/gpu_time.py
echo "== overlap 115"; JRB_OVERLAP_TRACER=1 NPK=115 WITH_E=0 timeout 200 python tools/gpu_time.py
echo "== 736 threads, no overlap 115"; JRB_EGA_THREADS=736 NPK=115 WITH_E=0 timeout 200 python tools/gpu_time.py
echo "== no overlap 460"; NPK=460 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== overlap 460"; JRB_OVERLAP_TRACER=1 NPK=460 WITH_E=0 timeout 300 python tools/gpu_time.py
} > $out/variants_r2j.log 2>&1
grep -E "^==|^\[|Error" $out/variants_r2j.log
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2j_all.log 2>&1; echo "pytest all rc=$?"; tail -4 $out/pytest_gpu_r2j_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_r2j.json 2> $out/bench_r2j.err; echo "bench rc=$?"
tail -c 400 $out/bench_r2j.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2j.json'))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["kernel"], "kernel_ms", d["roofline"]["kernel_ms"], "rt", d["roofline"]["raytrace_ms_per_step"], "parity", d["parity"]["ok"], "traffic", d["roofline"]["traffic"])
print(d["extra"]["single_package"]); e=d["extra"]["config_e"]; print("E", e["value"], e["e2e"]["value"], e["roofline"]["frac"])
PY
