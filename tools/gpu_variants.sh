#!/usr/bin/env bash
# time kernel build variants (development helper): tools/gpu_variants.sh <libdir>...
for d in "$@"; do
  echo "=== $d"
  JRB_LIBDIR=$d WITH_E=1 NPK=32 python tools/gpu_time.py 2>&1 | grep -E "^\[" 
done
