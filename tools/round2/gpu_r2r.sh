#!/usr/bin/env bash
# round 2, last 1-GPU visit: final tree -- full tests, smoke, default bench line
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r2r.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu_r2r.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2r.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_r2r.log
timeout 900 python bench.py > $out/bench_r2r.json 2> $out/bench_r2r.err; echo "bench rc=$?"
tail -c 300 $out/bench_r2r.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2r.json'))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], "traffic", d["roofline"]["traffic"], "parity", d["parity"]["ok"])
print(d["extra"]["single_package"]); e=d["extra"]["config_e"]; print("E", e["value"], e["e2e"]["value"], e["roofline"]["frac"], e["parity"]["ok"])
PY
