#!/usr/bin/env bash
# round 2, 2-GPU visit: multi-device tests, C caller on 2 devices, torchrun bench at N=2, single-package latency
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/r2c_gpus.txt
timeout 900 python -m pytest tests/test_gpu_io_and_lanes.py tests/test_gpu_c_callers.py tests/test_gpu_split_mode.py tests/test_gpu_parity.py -m gpu -q \
   -k "io_and_lanes or c_callers or split_mode or pipelined or full_baseline or channel_dependent" > $out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_r2c.log
tail -15 $out/pytest_gpu_r2c.log
cat $out/c_caller_2gpu.json
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 \
   > $out/bench_n2_r2c.json 2> $out/bench_n2_r2c.err; echo "bench n2 rc=$?"
tail -c 1500 $out/bench_n2_r2c.err; head -c 9000 $out/bench_n2_r2c.json; echo
timeout 120 python tools/gpu_single.py > $out/single_r2c.log 2>&1; cat $out/single_r2c.log
