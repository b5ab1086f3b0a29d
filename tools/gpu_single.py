#!/usr/bin/env python3
"""Single-package latency (development helper): one Config-D package through jrb_formod_batch, staged and page-locked."""
import importlib, os, sys, time, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
jr = importlib.import_module("jurassic-gpu_b200")
ctl = jr.synth.control_config_d(); tbl = jr.synth.make_tables(ctl)
pkg = jr.synth.limb_package(ctl, seed=20240517)
ctx = jr.Context(0); ctx.set_control(ctl); ctx.set_tables(tbl)
for mode in ("staged", "direct"):
    p = copy.deepcopy(pkg)
    if mode == "direct":
        jr.core.register_package(p)
    for _ in range(3):
        ctx.formod_batch([p])
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        ctx.formod_batch([p])
    dt = (time.perf_counter() - t0) / n * 1e3
    st = ctx.stats()
    print(f"[single {mode}] wall {dt:.3f} ms per call; device {st['ms_total_device']:.3f} ms (tracer {st['ms_raytrace']:.3f}, ega {st['ms_ega']:.3f}), "
          f"stage {st['host_ms_stage']:.3f} ms, scatter {st['host_ms_scatter']:.3f} ms, gas blocks {st['ega_gas_blocks']}, direct {st['io_direct']}", flush=True)
jr.core.host_unregister_all()
ctx.close()
