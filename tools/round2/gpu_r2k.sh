#!/usr/bin/env bash
# round 2, 1-GPU visit: Config E with 32-channel groups on the tiled kernel
set -u
out=gpurun_out; mkdir -p $out
{
echo "== E default (16-channel groups, fused lock-step)"; NPK=1 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== E cpw=32 tiled T=8"; JRB_EGA_CPW=32 NPK=1 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== E cpw=32 tiled T=8 free-running"; JRB_EGA_LOCKSTEP=0 JRB_EGA_CPW=32 NPK=1 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== E cpw=32 tiled T=4"; JRB_LIBDIR=$PWD/build/lib_t4 JRB_EGA_CPW=32 NPK=1 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== E cpw=32 tiled T=6"; JRB_LIBDIR=$PWD/build/lib_t6 JRB_EGA_CPW=32 NPK=1 NPK_E=58 timeout 300 python tools/gpu_time.py
echo "== E cpw=32 fused"; JRB_EGA_TILED=0 JRB_EGA_CPW=32 NPK=1 NPK_E=58 timeout 300 python tools/gpu_time.py
} > $out/variants_r2k.log 2>&1
grep -E "^==|^\[E|Error" $out/variants_r2k.log
