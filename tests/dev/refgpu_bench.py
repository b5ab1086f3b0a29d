#!/usr/bin/env python3
"""The kernel to beat, on the same GPU: the reference's own CUDA path (src/GPUdrivers.cu, unmodified, recompiled for
sm_100a by `make -C oracle refgpu`) on Config-D packages, beside this repo's path on the same packages.

Lives under tests/ because it drives the reference through oracle/.  Facts that shape the measurement
(SURVEY.md Appendix D #14, #21): formod_one_package repeats its kernel sequence 100 times per call (hard-coded benchmark
loop) around one H2D and one D2H copy, and formod_GPU must run with OMP_NUM_THREADS=1.  So one call = 100 device passes
over one 1088-ray package; the rate below counts all 100 and is therefore the reference's DEVICE path with its
transfers amortised 100-fold.  The reference's concurrency mechanism (up to 4 lanes, one per calling host thread) is
exercised with 4 caller threads.  Tables reach the reference through its binary cache (READ_BINARY=1), written natively.

usage: python tests/dev/refgpu_bench.py [calls_per_thread]   -> one JSON line, also gpurun_out/refgpu_bench.json
"""
import os
os.environ["OMP_NUM_THREADS"] = "1"
import copy, ctypes as C, importlib, json, sys, tempfile, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
jr = importlib.import_module("jurassic-gpu_b200")
import refdrv

ND, NG = 32, 5


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, seed=20240517 + i) for i in range(4)]
    work = tempfile.mkdtemp(prefix="refgpu_")
    os.chdir(work)
    jr.core.write_binary_tables(jr.core.binary_tables_filename(NG, ND), tbl, ctl, NG=NG, ND=ND)
    ref = refdrv.Reference(ND, NG, gpu=True)
    ref.lib.jrref_set_threads(1)
    ctl.tblbase = "/nonexistent/boxcar"
    c = ref.make_ctl(ctl, useGPU=1)
    c.read_binary = 1
    atms = [ref.make_atm(p) for p in pkgs]
    obss = [ref.make_obs(p) for p in pkgs]

    t0 = time.time()
    ref.formod(c, atms[0], obss[0])           # first call: tables, lanes, unified-memory migration
    t_first = time.time() - t0
    got = copy.deepcopy(pkgs[0]); ref.read_obs(obss[0], got)

    # parity of the reference's GPU path with this repo's path on the same package (sanity of the comparison)
    ctx = jr.Context(0)
    ctx.set_control(ctl); ctx.set_tables(tbl)
    mine = [copy.deepcopy(p) for p in pkgs]
    ctx.formod_batch(mine)
    rel = float(np.max(np.abs(got.rad - mine[0].rad) / (np.abs(mine[0].rad) + 1e-12 * np.max(np.abs(mine[0].rad)))))

    rc_per_call = 100 * pkgs[0].n_rays * ctl.nd   # the hard-coded 100-fold loop
    res = {}
    for nthreads in (1, 4):
        def worker(i):
            for _ in range(calls):
                ref.formod(c, atms[i], obss[i])
        th = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
        t0 = time.time()
        [x.start() for x in th]; [x.join() for x in th]
        dt = time.time() - t0
        res[f"lanes{nthreads}"] = {"calls": calls * nthreads, "seconds": round(dt, 3),
                                   "ms_per_package_pass": round(dt / (calls * nthreads * 100) * 1e3, 4),
                                   "ray_channels_per_s": rc_per_call * calls * nthreads / dt}

    # this repo on the same 4 packages (a batch 29x smaller than the bench's, so launch tails weigh more) and device path only
    ctx.stage(pkgs)
    for _ in range(3):
        ctx.run_staged()
    st = ctx.stats()
    ours_small = st["n_ray_channels"] / st["ms_total_device"] * 1e3
    out = {"what": "reference GPUdrivers.cu (sm_100a, stock flags, unmodified) vs this repo, Config-D packages, same B200",
           "reference_gpu": res, "reference_first_call_s": round(t_first, 2),
           "max_rel_diff_rad_reference_gpu_vs_ours": rel,
           "ours_4_packages_device_path": {"ms": st["ms_total_device"], "ray_channels_per_s": ours_small},
           "note": "reference rate counts the 100 passes of its built-in benchmark loop per call (transfers amortised 100x)"}
    line = json.dumps(out)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "refgpu_bench.json"), "w").write(line + "\n")


if __name__ == "__main__":
    main()
