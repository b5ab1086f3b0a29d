#!/usr/bin/env bash
# round 2: tracer with the Newton forms of sqrt / reciprocal in the stepping loop -- parity, single-package latency, big-batch tracer time
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_split_mode.py tests/test_gpu_atm_2d_3d.py -q -x > $out/pytest_r2t.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_r2t.log | cut -c1-300
timeout 300 python tools/gpu_single.py > $out/single_r2t.log 2>&1; cat $out/single_r2t.log
JRB_NO_COOP_TRACER=1 timeout 300 python tools/gpu_single.py 2>&1 | sed 's/^/[thread per ray] /' | tee -a $out/single_r2t.log
timeout 600 python bench.py --no-config-e --no-cpu-baseline --steps 3 --warmup 3 > $out/bench_r2t.json 2> $out/bench_r2t.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2t.json'))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "extra", {k:v for k,v in d["extra"].items() if k in ("ms_raytrace","ms_ega","single_package")})
PY
