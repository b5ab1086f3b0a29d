#!/usr/bin/env bash
# development helper: scan the work-chunk size of the EGA kernel (JRB_EGA_CHUNK) -- tools/gpu_chunk_scan.sh <npk> <chunks...>
npk=${1:-32}; shift
for c in "$@"; do echo "=== chunk $c"; JRB_EGA_CHUNK=$c WITH_E=${WITH_E:-1} NPK=$npk python tools/gpu_time.py 2>&1 | grep -E "^\["; done
