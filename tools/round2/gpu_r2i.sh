#!/usr/bin/env bash
# round 2, 1-GPU visit: tracer beside the EGA kernel (watermark) -- tests, A/B, bench
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_split_mode.py tests/test_gpu_parity.py -m gpu -q -x -k "tiled or full_baseline or large_batch or pipelined" > $out/pytest_gpu_r2i.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2i.log
tail -6 $out/pytest_gpu_r2i.log
{
echo "== overlap 115"; NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== no overlap 115"; JRB_OVERLAP_TRACER=0 NPK=115 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== overlap 32"; NPK=32 WITH_E=0 timeout 300 python tools/gpu_time.py
echo "== no overlap 32"; JRB_OVERLAP_TRACER=0 NPK=32 WITH_E=0 timeout 300 python tools/gpu_time.py
} > $out/variants_r2i.log 2>&1
grep -E "^==|^\[" $out/variants_r2i.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-config-e > $out/bench_r2i.json 2> $out/bench_r2i.err; echo "bench rc=$?"
tail -c 600 $out/bench_r2i.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2i.json'))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["kernel"], "kernel_ms", d["roofline"]["kernel_ms"], "rt", d["roofline"]["raytrace_ms_per_step"], "parity", d["parity"]["ok"], d["roofline"]["traffic"])
print(d["extra"]["single_package"])
PY
