#!/usr/bin/env python3
"""Timing only (development helper): Config D (NPK packages) and Config E (8 packages), specialised kernel."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
jr = importlib.import_module("jurassic-gpu_b200")
synth = jr.synth

def timing(name, ctl, tbl, pkgs, reps=3):
    ctx = jr.Context(0); ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(0 if os.environ.get('GENERIC', '0') == '1' else 1)
    if os.environ.get("FOV", "0") == "1":  # field-of-view epilogue with the 13-point test shape
        import numpy as np
        shape = np.loadtxt(os.path.join(ROOT, "tests", "golden", "fov_shape.tab"))
        ctx.set_fov(shape[:, 0], shape[:, 1])
    ctx.stage(pkgs)
    best = None
    for _ in range(reps):
        ctx.run_staged(); st = ctx.stats()
        if best is None or st['ms_total_device'] < best['ms_total_device']: best = st
    rc = best['n_ray_channels']
    print(f"[{name}] rays={best['n_rays']} rt={best['ms_raytrace']:.2f}ms ega={best['ms_ega']:.2f}ms total={best['ms_total_device']:.2f}ms "
          f"-> total {rc/best['ms_total_device']/1e3:.2f} M/s, ega-only {rc/best['ms_ega']/1e3:.2f} M/s", flush=True)
    ctx.close()

ctl = synth.control_config_d(); tbl = synth.make_tables(ctl, axis_jitter=os.environ.get("JITTER", "0") == "1")
timing("D" + (" jitter" if os.environ.get("JITTER", "0") == "1" else ""), ctl, tbl, [synth.limb_package(ctl, seed=20240517 + i) for i in range(int(os.environ.get("NPK", "32")))])
if os.environ.get("WITH_E", "1") == "1":
    ctl = synth.control_config_e(); tbl = synth.make_tables(ctl)
    timing("E", ctl, tbl, [synth.nadir_package(ctl, seed=20240518 + i) for i in range(int(os.environ.get("NPK_E", "8")))])
if os.environ.get("WITH_C", "0") == "1":  # refspec shape: 30 gases (channels cut to 32 to keep the tables small)
    gases = ["CO2", "H2O", "O3", "N2O", "CH4", "CO", "HNO3", "SO2", "F11", "CCl4"] + [f"X{i}" for i in range(20)]
    ctl = jr.Control(gases, 2150.0 + synth.np.arange(32)); tbl = synth.make_tables(ctl)
    pk = [synth.limb_package(ctl, seed=20240517 + i) for i in range(int(os.environ.get("NPK_C", "8")))]
    for p in pk: p.q[10:, :] = 1e-9
    timing("C-like ng=30", ctl, tbl, pk)
if os.environ.get("WITH_J", "0") == "1":  # Jacobian-like batch: one package, NJ atmospheres with a single perturbed temperature each
    import copy
    ctl = synth.control_config_d(); tbl = synth.make_tables(ctl)
    base = synth.limb_package(ctl, seed=20240517)
    pk = []
    for j in range(int(os.environ.get("NJ", "115"))):
        p = copy.deepcopy(base); p.t[(7 * j) % p.n_atm] += 1.0; pk.append(p)
    timing("J-like (perturbed copies of one package)", ctl, tbl, pk)
if os.environ.get("WITH_A", "0") == "1":
    ctl = synth.control_limb_example(); tbl = synth.make_tables(ctl)
    timing("A-like nd=2", ctl, tbl, [synth.limb_package(ctl, seed=20240517 + i) for i in range(32)])

if os.environ.get("WITH_R", "0") == "1":  # refspec shape at full width: 30 gases x 100 channels, 66-ray packages
    gases = ["CO2", "H2O", "O3", "N2O", "CH4", "CO", "HNO3", "SO2", "F11", "CCl4"] + [f"X{i}" for i in range(20)]
    ctl = jr.Control(gases, 2150.0 + synth.np.arange(100)); tbl = synth.make_tables(ctl)
    pk = [synth.limb_package(ctl, n_profiles=1, rays_per_profile=66, z0=3.0, dz=1.0, seed=20240517 + i) for i in range(int(os.environ.get("NPK_R", "64")))]
    for p in pk:
        for ig in range(10, 30): p.q[ig, :] = p.q[ig % 10, :] * (0.2 + 0.05 * ig)
    timing("refspec 30 gases x 100 channels", ctl, tbl, pk)
    os.environ["JRB_NO_SPLIT"] = "1"
    timing("refspec 30 gases x 100 channels, fused", ctl, tbl, pk)
