#!/usr/bin/env bash
# round 2, 8-GPU visit: multi-device tests, plain-C caller over 8 devices, torchrun bench at N=8 (and N=4 if time permits)
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/r2h_gpus.txt; nproc >> $out/r2h_gpus.txt
timeout 600 python -m pytest tests/test_gpu_io_and_lanes.py tests/test_gpu_c_callers.py tests/test_gpu_split_mode.py -m gpu -q -k "several_devices or c_caller or tiled or single_package_runs" > $out/pytest_gpu_r2h.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2h.log
tail -6 $out/pytest_gpu_r2h.log; cat $out/c_caller_8gpu.json
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 \
   > $out/bench_n8_r2h.json 2> $out/bench_n8_r2h.err; echo "bench n8 rc=$?"
tail -c 800 $out/bench_n8_r2h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8_r2h.json'))
print("N8 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_ms"])
print("parity", d["parity"]); print("gather", d["extra"]["gather"]); print("tables", d["extra"]["tables"])
e=d["extra"]["config_e"]; print("E", e["value"], e["e2e"]["value"], e["parity"])
PY
