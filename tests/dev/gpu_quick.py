#!/usr/bin/env python3
"""Quick GPU sanity run: CUDA vs CPU oracle on small cases + first timings (development helper; lives under tests/
because it uses the oracle as checker)."""
import copy, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
jr = importlib.import_module("jurassic-gpu_b200")
import refdrv
synth = jr.synth

def rel(a, b, floor):
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))

def check(name, ctl, tbl, pkgs, variant, orc):
    ctx = jr.Context(0)
    ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(variant)
    mine = [copy.deepcopy(p) for p in pkgs]
    t0 = time.time(); ctx.formod_batch(mine); t1 = time.time()
    st = ctx.stats()
    worst = (0, 0)
    for p, m in zip(pkgs, mine):
        o = copy.deepcopy(p); orc.formod(ctl, tbl, o)
        fl = 1e-12 * np.max(np.abs(o.rad))
        worst = (max(worst[0], rel(m.rad, o.rad, fl)), max(worst[1], rel(m.tau, o.tau, 1e-12)))
        tpe = max(np.max(np.abs(m.tpz - o.tpz)), np.max(np.abs(m.tplat - o.tplat)), np.max(np.abs(m.tplon - o.tplon)))
    print(f"[{name}] variant={st['ega_kernel_variant']} ngb={st['ega_ngb']} mask={st['ega_ctm_mask']} rel_rad={worst[0]:.3e} rel_tau={worst[1]:.3e} tp_err={tpe:.2e} "
          f"wall={t1-t0:.3f}s rt={st['ms_raytrace']:.2f}ms ega={st['ms_ega']:.2f}ms los={st['n_los_points']}", flush=True)
    ctx.close()

def timing(name, ctl, tbl, pkgs, variant, reps=3):
    ctx = jr.Context(0)
    ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(variant)
    ctx.stage(pkgs)
    for _ in range(reps):
        ctx.run_staged()
        st = ctx.stats()
        rc = st['n_ray_channels']
        print(f"[{name}] variant={st['ega_kernel_variant']} rays={st['n_rays']} rt={st['ms_raytrace']:.2f}ms ega={st['ms_ega']:.2f}ms total={st['ms_total_device']:.2f}ms "
              f"-> {rc/st['ms_total_device']*1e3/1e6:.3f} M ray-ch/s  Sbar={st['n_los_points']/max(1,st['n_rays']):.1f}", flush=True)
    ctx.close()

def main():
    orc = refdrv.Oracle()
    print("oracle threads", orc.threads(), flush=True)
    ctl = synth.control_limb_example(); tbl = synth.make_tables(ctl)
    pk = synth.example_package("limb", ctl)
    check("A limb generic", ctl, tbl, [pk], 0, orc)
    check("A limb fast", ctl, tbl, [pk], 1, orc)
    ctl = synth.control_nadir_example(); tbl = synth.make_tables(ctl)
    pk = synth.example_package("nadir", ctl)
    check("B nadir generic", ctl, tbl, [pk], 0, orc)
    check("B nadir fast", ctl, tbl, [pk], 1, orc)
    ctl = synth.control_config_d(); t0 = time.time(); tbl = synth.make_tables(ctl); print("D tables", time.time() - t0, flush=True)
    small = synth.limb_package(ctl, n_profiles=2, rays_per_profile=64, seed=1)
    check("D small generic", ctl, tbl, [small], 0, orc)
    check("D small fast", ctl, tbl, [small], 1, orc)
    pkgs = [synth.limb_package(ctl, seed=20240517 + i) for i in range(int(os.environ.get("NPK", "16")))]
    timing("D fast", ctl, tbl, pkgs, 1)
    timing("D generic", ctl, tbl, pkgs[:4], 0, reps=2)
    if os.environ.get("WITH_E", "1") == "1":
        ctl = synth.control_config_e(); t0 = time.time(); tbl = synth.make_tables(ctl); print("E tables", time.time() - t0, flush=True)
        small = synth.nadir_package(ctl, n_profiles=1, rays_per_profile=32, seed=2)
        check("E small generic", ctl, tbl, [small], 0, orc)
        check("E small fast", ctl, tbl, [small], 1, orc)
        pkgs = [synth.nadir_package(ctl, seed=20240518 + i) for i in range(8)]
        timing("E fast", ctl, tbl, pkgs, 1)

if __name__ == "__main__":
    main()
