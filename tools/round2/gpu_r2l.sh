#!/usr/bin/env bash
# round 2, 1-GPU visit: sorted tail of the work list, deferred mask scan, Config E with 32-channel groups on the tiled kernel; full tests + bench
set -u
out=gpurun_out; mkdir -p $out
{
echo "== D 115 default (tail sorted)"; NPK=115 WITH_E=0 timeout 200 python tools/gpu_time.py
echo "== D 115 tail not sorted"; JRB_NO_TAIL_SORT=1 NPK=115 WITH_E=0 timeout 200 python tools/gpu_time.py
echo "== D 32 default"; NPK=32 WITH_E=0 timeout 200 python tools/gpu_time.py
echo "== D 32 tail not sorted"; JRB_NO_TAIL_SORT=1 NPK=32 WITH_E=0 timeout 200 python tools/gpu_time.py
} > $out/variants_r2l.log 2>&1
grep -E "^==|^\[|Error" $out/variants_r2l.log
bash tools/round2/gpu_r2k.sh
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2l.log 2>&1; echo "pytest all rc=$?"; tail -4 $out/pytest_gpu_r2l.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-config-e > $out/bench_r2l.json 2> $out/bench_r2l.err; echo "bench rc=$?"
tail -c 400 $out/bench_r2l.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2l.json'))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_ms"], "kernel_ms", d["roofline"]["kernel_ms"], "parity", d["parity"]["ok"])
print(d["extra"]["single_package"])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --packages 115 --no-config-e --no-cpu-baseline > $out/bench115_r2l.json 2> $out/bench115_r2l.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench115_r2l.json'))
print("115 pkgs: value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_ms"], "kernel_ms", d["roofline"]["kernel_ms"])
PY
