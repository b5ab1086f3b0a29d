#!/usr/bin/env python3
"""SASS instruction histogram of the hot kernels of libjurassic_b200.so (no GPU needed): tools/sass_histogram.py > profiles/rN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jurassic-gpu_b200", "lib", "libjurassic_b200.so")
WANT = [("ega_tiled_kernel<12,0,0>  (Config D: CO2+H2O continua, segment-tiled, the default)", "ega_tiled_kernelILi12ELb0ELb0EE"),
        ("ega_tiled_kernel<15,0,0>  (Config E: all continua, segment-tiled, 32-channel groups)", "ega_tiled_kernelILi15ELb0ELb0EE"),
        ("ega_tiled_kernel<0,0,1>   (segment-tiled gas-block pass of the split mode)", "ega_tiled_kernelILi0ELb0ELb1EE"),
        ("ega_fast_kernel<12,0,0,0,0>  (Config D, segment by segment: the round-1 form)", "ega_fast_kernelILi12ELb0ELb0ELb0ELb0EE"),
        ("ega_fast_kernel<15,1,0,0,0>  (Config E: all continua, 16 channels per warp, fused)", "ega_fast_kernelILi15ELb1ELb0ELb0ELb0EE"),
        ("ega_fast_kernel<0,1,0,1,0>   (gas-block pass, several rays per warp)", "ega_fast_kernelILi0ELb1ELb0ELb1ELb0EE"),
        ("ega_fast_kernel<12,0,1,0,1>  (channel-dependent axes)", "ega_fast_kernelILi12ELb0ELb1ELb0ELb1EE"),
        ("ega_combine_kernel", "ega_combine_kernel"), ("ega_segment_kernel", "ega_segment_kernel"),
        ("ray_step_kernel<1>  (thread per ray)", "ray_step_kernelILi1EE"), ("ray_step_kernel<8>  (8 lanes per ray)", "ray_step_kernelILi8EE"), ("ray_geo_kernel  (2-D / 3-D atmospheres)", "ray_geo_kernel"), ("los_finalize_kernel", "los_finalize_kernel"), ("stage_kernel", "stage_kernel")]
KEYS = ["UBLKCP", "SYNCS", "LDG", "LDS", "STS", "STG", "DFMA", "DMUL", "DADD", "DSETP", "F2F", "MUFU", "FSETP", "IMAD", "SHFL", "BAR", "BRA", "ATOM"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
funcs, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
print(f"# SASS instruction histogram (cuobjdump -sass {os.path.relpath(LIB, ROOT)}; cubins: {', '.join(arch)})\n")
print("TMA bulk copies appear as `UBLKCP`, mbarrier traffic as `SYNCS.*`; there is no tcgen05/UTCMMA by design (no contraction on this path).\n")
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for title, pat in WANT:
    hits = [f for f in funcs if pat in f]
    if not hits:
        print(f"| {title} | (not found) |" + " |" * len(KEYS)); continue
    ins = funcs[hits[0]]
    c = collections.Counter()
    for i in ins:
        for k in KEYS:
            if i.startswith(k):
                c[k] += 1
    print(f"| {title} | {len(ins)} | " + " | ".join(str(c[k]) for k in KEYS) + " |")
tot = collections.Counter()
for f, ins in funcs.items():
    for i in ins:
        if i.startswith("UBLKCP"): tot["UBLKCP"] += 1
        if i.startswith("SYNCS"): tot["SYNCS"] += 1
        if "TCGEN" in i or "UTCMMA" in i: tot["tcgen05"] += 1
print(f"\nLibrary-wide: {len(funcs)} kernels, UBLKCP x{tot['UBLKCP']}, SYNCS x{tot['SYNCS']}, tcgen05/UTCMMA x{tot['tcgen05']}.")
