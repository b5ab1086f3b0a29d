#!/usr/bin/env bash
# round 2: validation of the tree after the 2-D/3-D row and the tracer change (full tests, smoke, default bench) + the tracer at 20 / 24 warps per SM
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r2u.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu_r2u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2u.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_r2u.log
timeout 900 python bench.py > $out/bench_r2u.json 2> $out/bench_r2u.err; echo "bench rc=$?"; tail -c 300 $out/bench_r2u.err
for occ in 5 6; do
  JRB_TRACER_OCC=$occ timeout 600 python bench.py --no-config-e --no-cpu-baseline --steps 3 --warmup 2 > $out/bench_r2u_occ$occ.json 2> $out/bench_r2u_occ$occ.err; echo "occ $occ rc=$?"
done
python - <<'PY'
import json
for f in ("bench_r2u","bench_r2u_occ5","bench_r2u_occ6"):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, "value", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "tracer", d["roofline"]["raytrace_ms_per_step"], "ega", d["roofline"]["kernel_ms"], "parity", d.get("parity",{}).get("ok"))
        if "config_e" in d["extra"]: e=d["extra"]["config_e"]; print("  E", e["value"], e["e2e"]["value"], e["roofline"]["frac"], e["parity"]["ok"])
        print("  single", d["extra"]["single_package"]["ms_device"], d["extra"]["single_package"]["ms_wall_per_call"])
    except Exception as ex: print(f, "failed", ex)
PY
