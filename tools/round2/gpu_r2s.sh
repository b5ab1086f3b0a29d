#!/usr/bin/env bash
# round 2: 2-D / 3-D atmosphere interpolation (ip = 2, 3) on the CUDA path -- new tests first, then the whole GPU suite
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_atm_2d_3d.py -q > $out/pytest_atm23_r2s.log 2>&1; echo "atm23 rc=$?"; tail -30 $out/pytest_atm23_r2s.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_atm_2d_3d.py > $out/pytest_gpu_r2s.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu_r2s.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2s.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_r2s.log
