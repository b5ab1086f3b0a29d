"""Parity of the CUDA path against the CPU oracle, through the C ABI (include/jurassic_b200.h).  Needs a GPU.

Tolerance (north-star): <= 1e-6 relative on rad and tau in double, with the absolute floors of SURVEY.md section 8c
for entries below the opaque cut-off; tangent points <= 1e-9 (km / deg).  Observed errors are ~1e-11.
"""
import copy
import os

import numpy as np
import pytest

from helpers import assert_parity, run_cuda, run_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _both(gpu_ctx_factory, oracle, ctl, tbl, pkgs, what, variants=(1, 0)):
    ref = run_oracle(oracle, ctl, tbl, pkgs)
    ctx = gpu_ctx_factory()
    outs = {}
    for v in variants:
        mine = run_cuda(ctx, ctl, tbl, pkgs, v)
        st = ctx.stats()
        assert st["ega_kernel_variant"] == v and st["n_kernel_launches"] >= 2
        for m, r in zip(mine, ref):
            assert_parity(m, r, f"{what} variant {v}")
        outs[v] = mine
    return outs, ref


def test_config_a_limb_example(jr, oracle, gpu_ctx_factory):
    ctl = jr.synth.control_limb_example()
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [jr.synth.example_package("limb", ctl)], "A")
    # the reference's golden geometry columns (example/limb/rad.org)
    gold = jr.synth.read_tab(os.path.join(ROOT, "tests", "golden", "limb", "rad.org"))
    assert np.allclose(gold[:, 7], outs[1][0].tpz, rtol=5.1e-6, atol=2e-6)
    assert np.allclose(gold[:, 9], outs[1][0].tplat, rtol=5.1e-6, atol=2e-6)


def test_config_b_nadir_example_bt_and_surface(jr, oracle, gpu_ctx_factory):
    ctl = jr.synth.control_nadir_example()
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [jr.synth.example_package("nadir", ctl)], "B")
    gold = jr.synth.read_tab(os.path.join(ROOT, "tests", "golden", "nadir", "rad.org"))
    assert np.allclose(gold[:, 7], outs[1][0].tpz, rtol=5.1e-6, atol=2e-6)
    assert np.allclose(gold[:, 9], outs[1][0].tplat, rtol=5.1e-6, atol=2e-6)


def test_config_c_refspec_shape_30_gases_100_channels(jr, oracle, gpu_ctx_factory):
    """NG=30, ND=100 (example/refspec shape)"""
    gases = ["CO2", "H2O", "O3", "N2O", "CH4", "CO", "HNO3", "SO2", "F11", "CCl4"] + [f"X{i}" for i in range(20)]
    ctl = jr.Control(gases, 2150.0 + np.arange(100))  # window with the N2 continuum
    assert ctl.ctm_mask == 14
    tbl = jr.synth.make_tables(ctl, skip_pairs=[(g, d) for g in range(10, 30) for d in range(0, 100, 3)])
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=6, dz=11.0, seed=5)
    for ig in range(10, 30):
        pkg.q[ig, :] = 1e-9
    _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg], "C")


def test_more_than_32_gases_use_generic_kernel(jr, oracle, gpu_ctx_factory):
    gases = [f"G{i}" for i in range(34)]
    ctl = jr.Control(gases, [900.0, 901.0])
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=3, dz=20.0, seed=9)
    pkg.q[:, :] = 2e-9
    _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg], "ng34", variants=(0,))
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(1)
    with pytest.raises(jr.JrbError, match="not applicable"):
        ctx.formod_batch([copy.deepcopy(pkg)])


def test_config_d_package(jr, oracle, gpu_ctx_factory):
    """one full Config-D package: 17 profiles x 64 rays x 32 channels x 5 gases"""
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, seed=20240517)
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg], "D", variants=(1,))
    assert ref[0].tau.min() < 1e-6  # covers opaque rays


def test_config_e_slice(jr, oracle, gpu_ctx_factory):
    """Config-E slice: 128 channels x 8 gases, all four continua, surface term"""
    ctl = jr.synth.control_config_e()
    assert ctl.ctm_mask == 15
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.nadir_package(ctl, n_profiles=4, rays_per_profile=17, dlat=0.7, seed=20240518)
    _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg], "E")


@pytest.mark.parametrize("bits", range(16))
def test_all_continuum_switches(jr, oracle, gpu_ctx_factory, bits):
    ctl = jr.Control(["CO2", "H2O", "O3"], [700.0, 850.5, 1400.0, 1805.0, 2200.0, 2605.0],
                     ctm_co2=(bits >> 3) & 1, ctm_h2o=(bits >> 2) & 1, ctm_n2=(bits >> 1) & 1, ctm_o2=bits & 1)
    pkg = jr.synth.nadir_package(ctl, n_profiles=1, rays_per_profile=5, seed=bits)
    pkg.k[0, :] = 1e-4 * np.exp(-pkg.z / 7.0)
    _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [pkg], f"ctm{bits:04b}")


@pytest.mark.parametrize("ng", [1, 2, 3, 4, 6, 7, 8])
def test_every_gas_count_of_the_specialised_kernel(jr, oracle, gpu_ctx_factory, ng):
    ctl = jr.Control(jr.synth.NADIR_GASES[:ng], [690.0, 1400.0, 2200.0])
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=6, dz=8.0, seed=ng)
    _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [pkg], f"ng{ng}")


def test_quirks(jr, oracle, gpu_ctx_factory):
    """NaN mask, missing tables, no CO2/H2O emitter, T below the table axis, observer inside, rejected rays, ground hit"""
    ctl = jr.Control(["O3", "F11", "N2O"], [792.0, 832.0, 1000.0, 2300.0])
    tbl = jr.synth.make_tables(ctl, skip_pairs=[(1, 0), (1, 1), (1, 2), (1, 3), (2, 2)])
    pkg = jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=8, dz=9.0, seed=3)
    pkg.t[:] -= 25.0
    pkg.obsz[3] = 40.0; pkg.vpz[3] = 10.0; pkg.vplat[3] = 3.0
    pkg.obsz[4] = -1.0
    pkg.vpz[5] = 95.0
    pkg.vpz[6] = -50.0; pkg.vplat[6] = 1.0
    pkg.rad[2, 1] = np.nan; pkg.rad[9, 0] = np.inf
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg], "quirks")
    m = outs[1][0]
    assert np.isnan(m.rad[2, 1]) and np.isnan(m.rad[9, 0])
    assert np.all(m.rad[4] == 0) and np.all(m.tau[4] == 1) and np.all(m.rad[5] == 0) and np.all(m.tau[5] == 1)
    assert m.tpz[4] == pkg.vpz[4] and m.tplat[5] == pkg.vplat[5]  # rejected rays keep the view point


def test_opaque_cutoff(jr, oracle, gpu_ctx_factory):
    ctl = jr.synth.control_limb_example()
    pkg = jr.synth.example_package("limb", ctl)
    pkg.q[0, :] *= 5000.0
    pkg.q[2, :] *= 2000.0
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [pkg], "opaque")
    assert ref[0].tau.min() < 1e-9


def test_channel_dependent_axes_are_located_per_lane(jr, oracle, gpu_ctx_factory):
    """(p,T) axes that differ between channels (init_tbl stores them per channel, src/jurassic.c:383-384; the reference
    locates them per channel, src/jr_common.h:237-247): the specialised kernel resolves the table cell per lane"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl, axis_jitter=True)
    assert jr.core.tables_pack_info(tbl, ctl.ng, ctl.nd)["all_shared"] == 0
    pkg = jr.synth.example_package("limb", ctl)
    ref = run_oracle(oracle, ctl, tbl, [pkg])
    ctx = gpu_ctx_factory()
    mine = run_cuda(ctx, ctl, tbl, [pkg], -1)
    st = ctx.stats()
    assert st["ega_kernel_variant"] == 1 and st["ega_per_channel_axes"] == 1
    assert_parity(mine[0], ref[0], "jitter, specialised kernel")
    assert_parity(run_cuda(ctx, ctl, tbl, [pkg], 0)[0], ref[0], "jitter, generic kernel")


def test_channel_dependent_axes_config_d_package(jr, oracle, gpu_ctx_factory):
    """a full Config-D package (32 channels per warp, one ray per warp) on tables with channel-dependent axes, fused and
    in gas-block passes; some pairs without a table, T below the axes (extrapolation + clamping)"""
    import os
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl, axis_jitter=True, skip_pairs=[(3, 5), (4, 31)])
    pkg = jr.synth.limb_package(ctl, seed=4711)
    ref = run_oracle(oracle, ctl, tbl, [pkg])[0]
    ctx = gpu_ctx_factory()
    split = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    st = ctx.stats()
    assert st["ega_per_channel_axes"] == 1 and st["ega_gas_blocks"] == 5
    assert_parity(split, ref, "jitter D, split")
    os.environ["JRB_NO_SPLIT"] = "1"
    try:
        fused = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
        assert ctx.stats()["ega_gas_blocks"] == 1
    finally:
        del os.environ["JRB_NO_SPLIT"]
    assert np.array_equal(split.rad, fused.rad) and np.array_equal(split.tau, fused.tau)


def test_gas_dependent_axes_use_per_gas_cells(jr, oracle, gpu_ctx_factory):
    """every gas has its own (p,T) grid: the LOS records carry one table cell per gas instead of a shared one"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl, gas_axis_shift=True)
    info = jr.core.tables_pack_info(tbl, ctl.ng, ctl.nd)
    assert info["all_shared"] == 1 and info["gas_axes_same"] == 0
    _both(gpu_ctx_factory, oracle, ctl, tbl, [jr.synth.example_package("limb", ctl)], "gas axes")


def test_no_refraction_and_extinction_windows(jr, oracle, gpu_ctx_factory):
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0], refrac=0, rayds=8.0, raydz=1.0)  # (rayds 5 would need > 400 points: fatal)
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=10, dz=6.0, seed=11)
    pkg.k[0, :] = 3e-4 * np.exp(-pkg.z / 6.0)
    _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [pkg], "norefrac")


def test_hydrostatic_adjustment(jr, oracle, gpu_ctx_factory):
    ctl = jr.synth.control_limb_example()
    ctl.hydz = 20.0
    pkg = jr.synth.example_package("limb", ctl)
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, jr.synth.make_tables(ctl), [pkg], "hydz", variants=(1,))
    assert np.allclose(outs[1][0].p, ref[0].p, rtol=1e-14)  # caller's atm->p is updated like on the reference's CPU path


def test_los_parity(jr, oracle, gpu_ctx_factory):
    """T3: per-point LOS of the device ray tracer vs the oracle's traceray"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.example_package("limb", ctl)
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(0)
    ctx.stage([copy.deepcopy(pkg)]); ctx.run_staged()
    ng, nw = ctl.ng, ctl.nw
    for ir in (0, 1, 20, 40, 65):
        los_o, ts_o = oracle.traceray(ctl, copy.deepcopy(pkg), ir)
        rec, ts = ctx.debug_los(ir)
        assert rec.shape[0] == los_o.shape[0] and ts == pytest.approx(ts_o, rel=1e-12)
        u0, tail = 4 + nw, rec.shape[1] - 6   # record layout: csrc/jrb_device.cuh (LosLayout)
        np.testing.assert_allclose(rec[:, tail], los_o[:, 0], rtol=1e-10, atol=1e-10)             # altitude
        np.testing.assert_allclose(rec[:, 0:2], los_o[:, 3:5], rtol=1e-10)                         # p, T
        np.testing.assert_allclose(rec[:, 2], los_o[:, 5], rtol=1e-9, atol=1e-9)                   # ds (trapezoid; the clipped last step is ill-conditioned: 1e-9 km = 1 um)
        np.testing.assert_allclose(rec[:, 4:4 + nw], los_o[:, 6:6 + nw], rtol=1e-9, atol=1e-300)   # extinction
        # column densities; the segment clipped at the atmosphere boundary inherits the conditioning of its length
        np.testing.assert_allclose(rec[:-2, u0:u0 + ng], los_o[:-2, 6 + nw + ng:], rtol=1e-9)
        np.testing.assert_allclose(rec[-2:, u0:u0 + ng], los_o[-2:, 6 + nw + ng:], rtol=1e-6)
        # Cartesian position of the point <-> the oracle's (z, lon, lat)
        lon, lat, zz = np.radians(los_o[:, 1]), np.radians(los_o[:, 2]), los_o[:, 0] + jr.synth.RE
        xyz = np.stack([zz * np.cos(lat) * np.cos(lon), zz * np.cos(lat) * np.sin(lon), zz * np.sin(lat)], axis=1)
        np.testing.assert_allclose(rec[:, tail + 3:tail + 6], xyz, rtol=0, atol=1e-8)


def test_batch_equals_loop_and_is_deterministic(jr, oracle, gpu_ctx_factory):
    """T7: a batch of K packages == K single calls, bit-identical; repeated runs are bit-identical"""
    ctl = jr.synth.control_config_d(nd=8)
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=16, dz=4.0, seed=100 + i) for i in range(5)]
    ctx = gpu_ctx_factory()
    batch = run_cuda(ctx, ctl, tbl, pkgs, 1)
    again = run_cuda(ctx, ctl, tbl, pkgs, 1)
    for b, a in zip(batch, again):
        assert np.array_equal(b.rad, a.rad) and np.array_equal(b.tau, a.tau)
    for i, p in enumerate(pkgs):
        one = run_cuda(ctx, ctl, tbl, [p], 1)[0]
        assert np.array_equal(one.rad, batch[i].rad) and np.array_equal(one.tau, batch[i].tau)
        assert np.array_equal(one.tpz, batch[i].tpz)


def test_los_chunking_gives_identical_results(jr, gpu_ctx_factory, monkeypatch):
    ctl = jr.synth.control_config_d(nd=8)
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=4, rays_per_profile=64, seed=300 + i) for i in range(6)]
    ctx = gpu_ctx_factory()
    full = run_cuda(ctx, ctl, tbl, pkgs, 1)
    monkeypatch.setenv("JRB_LOS_GB", "0.02")  # forces several LOS chunks (floor: 1024 rays per chunk)
    ctx2 = gpu_ctx_factory()
    chunked = run_cuda(ctx2, ctl, tbl, pkgs, 1)
    assert ctx2.stats()["n_kernel_launches"] > ctx.stats()["n_kernel_launches"]
    for a, b in zip(full, chunked):
        assert np.array_equal(a.rad, b.rad) and np.array_equal(a.tau, b.tau)


def test_pipelined_chunks_give_identical_results(jr, gpu_ctx_factory, monkeypatch):
    """JRB_PIPELINE=1: tracer of chunk c+1 beside the EGA kernel of chunk c on separate streams, 3 rotating LOS buffers"""
    ctl = jr.synth.control_config_d(nd=4)
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, seed=500 + i) for i in range(34)]  # 36 992 rays >= the pipelining threshold
    ctx = gpu_ctx_factory()
    monkeypatch.setenv("JRB_NO_SPLIT", "1")  # the pipeline runs the fused kernel; gas blocks of several gases regroup the product
    plain = run_cuda(ctx, ctl, tbl, pkgs, 1)
    assert ctx.stats()["pipelined"] == 0
    monkeypatch.setenv("JRB_PIPELINE", "1")
    ctx2 = gpu_ctx_factory()
    piped = run_cuda(ctx2, ctl, tbl, pkgs, 1)
    st = ctx2.stats()
    assert st["pipelined"] == 1 and st["n_chunks"] >= 4
    for a, b in zip(plain, piped):
        assert np.array_equal(a.rad, b.rad) and np.array_equal(a.tau, b.tau) and np.array_equal(a.tpz, b.tpz)


def test_large_batch_properties(jr, oracle, gpu_ctx_factory):
    """Size-independent properties at a bench-like size (32 Config-D packages = 34 816 rays x 32 channels):
    duplicated packages give bit-identical results wherever they sit in the batch, tau in [0,1], rad >= 0,
    and a sample of rays agrees with the oracle."""
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    base = [jr.synth.limb_package(ctl, seed=20240517 + i) for i in range(16)]
    pkgs = base + [copy.deepcopy(p) for p in reversed(base)]
    ctx = gpu_ctx_factory()
    out = run_cuda(ctx, ctl, tbl, pkgs, -1)
    st = ctx.stats()
    assert st["ega_kernel_variant"] == 1 and st["n_rays"] == 32 * 1088
    for i in range(16):
        assert np.array_equal(out[i].rad, out[31 - i].rad) and np.array_equal(out[i].tau, out[31 - i].tau)
    allrad = np.concatenate([o.rad for o in out]); alltau = np.concatenate([o.tau for o in out])
    assert np.all(np.isfinite(allrad)) and np.all(allrad >= 0) and np.all((alltau >= 0) & (alltau <= 1))
    # sample: 24 rays of package 7 through the oracle
    sub = copy.deepcopy(base[7])
    idx = np.arange(0, 1088, 46)
    small = jr.Package(ctl.ng, ctl.nw, ctl.nd, sub.n_atm, idx.size)
    for name in ("atm_time", "z", "lon", "lat", "p", "t", "q", "k"):
        getattr(small, name)[...] = getattr(sub, name)
    for name in ("time", "obsz", "obslon", "obslat", "vpz", "vplon", "vplat"):
        getattr(small, name)[:] = getattr(sub, name)[idx]
    oracle.formod(ctl, tbl, small)
    got = jr.Package(ctl.ng, ctl.nw, ctl.nd, sub.n_atm, idx.size)
    for name in ("rad", "tau", "tpz", "tplon", "tplat"):
        getattr(got, name)[...] = getattr(out[7], name)[idx]
    assert_parity(got, small, "large batch sample")


def test_edge_shapes_no_gases_odd_channel_count_empty_packages(jr, oracle, gpu_ctx_factory):
    """ng = 0 (continua + extinction only), nd = 33 (two channel groups, the second almost empty), a package with
    nr = 0 in the middle of a batch, and an empty batch"""
    ctl = jr.Control([], 2140.0 + 3.0 * np.arange(33))  # N2 continuum only
    assert ctl.ng == 0 and ctl.ctm_mask == 2
    tbl = jr.synth.make_tables(ctl)
    a = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=5, dz=10.0, seed=1)
    a.k[0, :] = 2e-4 * np.exp(-a.z / 8.0)
    empty = jr.Package(0, 1, 33, a.n_atm, 0)
    for name in ("atm_time", "z", "lon", "lat", "p", "t", "k"):
        getattr(empty, name)[...] = getattr(a, name)
    b = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=3, dz=20.0, seed=2)
    ref = run_oracle(oracle, ctl, tbl, [a, b])
    ctx = gpu_ctx_factory()
    for variant in (0, 1):
        out = run_cuda(ctx, ctl, tbl, [a, empty, b], variant)
        assert_parity(out[0], ref[0], f"ng0 a v{variant}")
        assert_parity(out[2], ref[1], f"ng0 b v{variant}")
    ctx.formod_batch([])  # empty batch is a no-op
    assert ctx.stats()["n_rays"] == 0


@pytest.mark.parametrize("nd", [1, 2, 3, 5, 7, 11, 16, 17])
def test_few_channels_pack_several_rays_per_warp(jr, oracle, gpu_ctx_factory, nd):
    """nd <= 16: floor(32/nd) rays share a warp (lane = ray_in_warp*nd + channel); rays of different length, opaque rays,
    rejected rays and a ray count that is not a multiple of the rays per warp"""
    ctl = jr.Control(jr.synth.LIMB_GASES, 780.0 + 4.0 * np.arange(nd))
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=23, dz=2.7, seed=nd)
    pkg.q[0, :] *= 40.0          # some channels of the low rays go opaque
    pkg.vpz[7] = 95.0            # rejected ray (np = 0) inside a warp
    pkg.vpz[30] = -20.0          # ray into the ground
    _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg], f"nd{nd}")


def test_random_geometries_and_descending_atmosphere(jr, oracle, gpu_ctx_factory):
    """randomised rays: space-borne limb and nadir-slant views, observers inside the atmosphere looking up, down and
    sideways; the same atmosphere stored bottom-up and top-down (the reference's `locate` handles both orders)"""
    rng = np.random.default_rng(42)
    ctl = jr.Control(["CO2", "H2O", "O3"], [700.0, 792.0, 1400.0, 2200.0])
    tbl = jr.synth.make_tables(ctl)
    n = 96
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=n, seed=8)
    kind = rng.integers(0, 4, n)
    for r in range(n):
        if kind[r] == 0:      # limb from orbit
            pkg.obsz[r] = rng.uniform(300, 900); pkg.vpz[r] = rng.uniform(2, 80)
            pkg.vplat[r] = np.degrees(np.arccos((jr.synth.RE + pkg.vpz[r]) / (jr.synth.RE + pkg.obsz[r])))
        elif kind[r] == 1:    # nadir / slant from orbit to the ground
            pkg.obsz[r] = rng.uniform(300, 900); pkg.vpz[r] = 0.0; pkg.vplat[r] = rng.uniform(-15, 15); pkg.vplon[r] = rng.uniform(-5, 5)
        elif kind[r] == 2:    # aircraft / balloon looking up or sideways
            pkg.obsz[r] = rng.uniform(8, 40); pkg.vpz[r] = rng.uniform(pkg.obsz[r] - 3, 85); pkg.vplat[r] = rng.uniform(0.2, 4)
        else:                 # aircraft looking down
            pkg.obsz[r] = rng.uniform(8, 40); pkg.vpz[r] = rng.uniform(0.0, pkg.obsz[r] - 1); pkg.vplat[r] = rng.uniform(0.0, 2)
    pkg.k[0, :] = 1e-4 * np.exp(-pkg.z / 7.0)
    flipped = copy.deepcopy(pkg)
    for name in ("atm_time", "z", "lon", "lat", "p", "t"):
        getattr(flipped, name)[:] = getattr(pkg, name)[::-1]
    flipped.q[:, :] = pkg.q[:, ::-1]
    flipped.k[:, :] = pkg.k[:, ::-1]
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, tbl, [pkg, flipped], "random geometry")
    # bottom-up and top-down storage describe the same atmosphere
    assert np.allclose(ref[0].rad, ref[1].rad, rtol=1e-9) and np.allclose(outs[1][0].rad, outs[1][1].rad, rtol=1e-9)
    assert (ref[0].tau < 1).any() and (ref[0].rad > 0).any()


def test_non_monotone_columns_are_bisected_like_the_reference(jr, oracle, gpu_ctx_factory):
    """a column that is not sorted in eps / u (possible after init_tbl's overwrite quirk, src/jurassic.c:369-384) is
    flagged at pack time and searched with the reference's plain bisection; all other columns keep the hinted search"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    bad = copy.deepcopy(tbl)
    for ip in range(8, 30):          # damage the columns the rays actually cross
        for it in range(0, 12, 2):
            n = bad.nu[0, ip, it, 0]
            k = n // 2
            bad.eps[0, ip, it, k, 0], bad.eps[0, ip, it, k + 1, 0] = bad.eps[0, ip, it, k + 1, 0], bad.eps[0, ip, it, k, 0]
            bad.u[2, ip, it, k + 3, 1] = bad.u[2, ip, it, k + 1, 1]
    info = jr.core.tables_pack_info(bad, ctl.ng, ctl.nd)
    assert info["monotone"] == 0 and info["all_shared"] == 1
    pkg = jr.synth.example_package("limb", ctl)
    outs, ref = _both(gpu_ctx_factory, oracle, ctl, bad, [pkg], "non-monotone")
    good = run_oracle(oracle, ctl, tbl, [pkg])[0]
    assert not np.allclose(good.rad, ref[0].rad, rtol=1e-9)  # the damaged entries are really used


def test_sharding_over_contexts_is_bit_identical(jr, gpu_ctx_factory):
    """T8: contiguous package slices processed by separate contexts (= ranks, one per GPU in production) give exactly
    the results of a single context processing everything; tables reach the second context as the packed blob"""
    ctl = jr.synth.control_config_d(nd=8)
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=32, seed=700 + i) for i in range(7)]
    one = gpu_ctx_factory()
    whole = run_cuda(one, ctl, tbl, pkgs, -1)
    blob = jr.core.tables_pack_host(tbl, ctl.ng, ctl.nd)
    got = []
    for rank in range(3):
        first, count = jr.shard.shard_range(len(pkgs), rank, 3)
        ctx = gpu_ctx_factory()
        ctx.set_control(ctl)
        ctx.tables_upload_blob(blob)            # what a rank receives from the broadcast
        part = [copy.deepcopy(p) for p in pkgs[first:first + count]]
        ctx.formod_batch(part)
        got += part
    for a, b in zip(whole, got):
        assert np.array_equal(a.rad, b.rad) and np.array_equal(a.tau, b.tau) and np.array_equal(a.tplat, b.tplat)


# ---- row f3: field-of-view convolution as a device epilogue ---------------------------------------------------------
def _fov_shape(jr):
    s = jr.synth.read_tab(os.path.join(ROOT, "tests", "golden", "fov_shape.tab"))
    return s[:, 0].copy(), s[:, 1].copy()


def test_fov_epilogue_matches_oracle(jr, oracle, gpu_ctx_factory):
    """jrb_set_fov: results == formod(); formod_fov(); of the reference (pinned bit for bit by the CPU test
    test_oracle_fov_matches_reference_formod_fov); ascending and descending scans, windows limited to the own package and
    time, NaN mask applied before the convolution; several packages per batch"""
    from test_oracle_vs_reference import fov_cases
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    dz, w = _fov_shape(jr)
    pkgs = [p for _, p in fov_cases(jr, ctl)] + [jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=2, z0=20.0, seed=3)]
    ref = run_oracle(oracle, ctl, tbl, pkgs)
    plain = copy.deepcopy(ref)
    for r in ref:
        assert oracle.formod_fov(r, dz, w)
    ctx = gpu_ctx_factory()
    for variant in (1, 0):
        ctx.set_control(ctl)
        ctx.set_tables(tbl)
        ctx.set_kernel_variant(variant)
        ctx.set_fov(dz, w)
        mine = [copy.deepcopy(p) for p in pkgs]
        ctx.formod_batch(mine)
        for m, r in zip(mine, ref):
            assert_parity(m, r, f"fov variant {variant}")
        assert np.isnan(mine[1].rad).sum() > 1  # the masked measurement spreads to its neighbours, as in the reference
        ctx.set_fov(None)                       # and off again: pencil-beam results
        mine = [copy.deepcopy(p) for p in pkgs]
        ctx.formod_batch(mine)
        for m, r in zip(mine, plain):
            assert_parity(m, r, f"fov off variant {variant}")


def test_fov_epilogue_fails_like_the_reference_on_single_ray_scans(jr, gpu_ctx_factory):
    ctl = jr.synth.control_limb_example()
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl)
    ctx.set_tables(jr.synth.make_tables(ctl))
    ctx.set_fov([0.0, 0.5], [1.0, 0.5])
    with pytest.raises(jr.core.JrbError, match="Cannot apply FOV convolution"):
        ctx.formod_batch([jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=1, seed=5)])
    ctx.set_fov(None)


def test_inputs_outside_the_reference_domain_stay_defined(jr, oracle, gpu_ctx_factory):
    """Two inputs that are undefined behaviour in the reference (out-of-bounds reads) must
    neither disturb the device nor other rays: temperatures outside the 100..400 K source table (here: the Planck index is
    clipped, the edge interval extrapolates) and rays whose time matches no profile (here: rejected, rad = 0, tau = 1)."""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    good = jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=8, z0=10.0, dz=4.0, seed=11)
    hot = copy.deepcopy(good)
    hot.t[:] += 250.0                      # 450..550 K
    cold = copy.deepcopy(good)
    cold.t[:] = 60.0
    lost = copy.deepcopy(good)
    lost.time[3] = 17.0                    # later than every profile
    lost.time[12] = -5.0                   # earlier than every profile
    ctx = gpu_ctx_factory()
    ref = run_oracle(oracle, ctl, tbl, [good])[0]
    for variant in (1, 0):
        out = run_cuda(ctx, ctl, tbl, [hot, good, cold, lost], variant)
        assert_parity(out[1], ref, f"neighbour of out-of-domain packages, variant {variant}")
        for o in (out[0], out[2]):
            assert np.all(np.isfinite(o.rad)) and np.all(np.isfinite(o.tau)) and np.all((o.tau >= 0) & (o.tau <= 1))
        assert np.all(out[3].rad[3] == 0.0) and np.all(out[3].tau[3] == 1.0)
        keep = np.ones(good.n_rays, bool)
        keep[[3, 12]] = False
        if not np.all(out[3].rad[12] == 0.0):  # a time before the first profile selects it (locate_atm, src/jr_common.h:127-154)
            keep[12] = True
        assert np.allclose(out[3].rad[keep], ref.rad[keep], rtol=1e-6, atol=0)


def test_full_baseline_size_in_one_call(jr, gpu_ctx_factory):
    """BASELINE.json's full Config-D size in ONE call: 920 packages = 1 000 960 rays x 32 channels (a LOS scratch
    limit of 24 GB forces several chunks).  Size-independent properties: the 8 copies of each of 115 distinct packages, scattered
    over the batch, come back bit-identical (checksum of checksums), tau in [0,1], rad finite and >= 0, every ray traced."""
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    base = [jr.synth.limb_package(ctl, seed=20240517 + i) for i in range(115)]
    order = np.random.default_rng(7).permutation(920)
    pkgs = [copy.deepcopy(base[j % 115]) for j in order]
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl)
    ctx.set_tables(tbl)
    jr.load_core().jrb_set_los_limit_gb(ctx.h, 24.0)  # the default (72 GB) would hold all rays in one chunk
    ctx.formod_batch(pkgs)
    st = ctx.stats()
    assert st["n_rays"] == 1000960 and st["n_ray_channels"] == 1000960 * 32 and st["ega_kernel_variant"] == 1
    assert st["n_chunks"] >= 2
    assert st["n_los_points"] > 250 * 1000960           # mean LOS length of the recipe is 257
    sums = {}
    for j, p in zip(order, pkgs):
        key = (p.rad.tobytes(), p.tau.tobytes(), p.tpz.tobytes())
        sums.setdefault(j % 115, set()).add(hash(key))
    assert len(sums) == 115 and all(len(s) == 1 for s in sums.values())
    assert len({next(iter(s)) for s in sums.values()}) == 115   # and distinct packages give distinct results
    for p in pkgs[::37]:
        assert np.all(np.isfinite(p.rad)) and np.all(p.rad >= 0) and np.all((p.tau >= 0) & (p.tau <= 1))


def test_lock_step_mode_is_chosen_for_equal_length_rays_and_changes_nothing(jr, oracle, gpu_ctx_factory, monkeypatch):
    """The specialised kernel starts the rays of a CTA together (lock step -> neighbouring warps share table brackets in L1)
    when the rays of a work chunk are of nearly equal length: nadir swaths yes, limb scans (130..393 segments) no.  The
    decision is made on the device; results are bit-identical either way."""
    ctx = gpu_ctx_factory()
    for name, ctl, pkgs, expect in (
            ("nadir", jr.synth.control_config_e(nd=32), [jr.synth.nadir_package(jr.synth.control_config_e(nd=32), n_profiles=2, rays_per_profile=68, seed=3 + i) for i in range(3)], 1),
            ("limb", jr.synth.control_config_d(nd=32), [jr.synth.limb_package(jr.synth.control_config_d(nd=32), n_profiles=2, rays_per_profile=64, seed=5 + i) for i in range(3)], 0)):
        tbl = jr.synth.make_tables(ctl)
        monkeypatch.delenv("JRB_EGA_LOCKSTEP", raising=False)
        auto = run_cuda(ctx, ctl, tbl, pkgs, 1)
        assert ctx.stats()["ega_phase_lock"] == expect, name
        ref = run_oracle(oracle, ctl, tbl, pkgs[:1])[0]
        assert_parity(auto[0], ref, f"lock step auto {name}")
        for forced in ("0", "1"):
            monkeypatch.setenv("JRB_EGA_LOCKSTEP", forced)
            out = run_cuda(ctx, ctl, tbl, pkgs, 1)
            assert ctx.stats()["ega_phase_lock"] == int(forced)
            for a, b in zip(auto, out):
                assert np.array_equal(a.rad, b.rad) and np.array_equal(a.tau, b.tau)
        monkeypatch.delenv("JRB_EGA_LOCKSTEP", raising=False)


def test_narrow_channel_groups_give_identical_results(jr, oracle, gpu_ctx_factory, monkeypatch):
    """The specialised kernel can give a warp fewer than 32 channels of several consecutive rays (chosen by the runtime when
    the tables of 32 channels x ng gases exceed the L2, here forced): any split gives bit-identical results, including a
    channel count that is not a multiple of the group width and a ray count that is not a multiple of the rays per warp."""
    ctl = jr.Control(jr.synth.LIMB_GASES, 785.0 + np.arange(37))
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=13, z0=5.0, dz=4.0, seed=21 + i) for i in range(2)]
    ctx = gpu_ctx_factory()
    monkeypatch.delenv("JRB_EGA_CPW", raising=False)
    base = run_cuda(ctx, ctl, tbl, pkgs, 1)
    assert ctx.stats()["ega_channels_per_warp"] == 32   # 37 channels x 5 gases of tables fit the L2 budget
    ref = run_oracle(oracle, ctl, tbl, pkgs)
    for b, r in zip(base, ref):
        assert_parity(b, r, "cpw default")
    for cpw in ("16", "8", "4"):
        monkeypatch.setenv("JRB_EGA_CPW", cpw)
        out = run_cuda(ctx, ctl, tbl, pkgs, 1)
        assert ctx.stats()["ega_channels_per_warp"] == int(cpw)
        for a, b in zip(base, out):
            assert np.array_equal(a.rad, b.rad) and np.array_equal(a.tau, b.tau), cpw
    monkeypatch.delenv("JRB_EGA_CPW", raising=False)


def test_too_many_los_points_is_an_error(jr, gpu_ctx_factory):
    """a ray that needs NLOS = 400 points or more: the reference's CPU path is fatal ("Too many LOS points!",
    src/jr_common.h:693-695); the CUDA path fails the call instead of returning a truncated ray.  Afterwards the context works."""
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0], rayds=4.0, raydz=0.2)
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=4, z0=2.0, dz=20.0, seed=1)
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl)
    with pytest.raises(jr.JrbError, match="Too many LOS points"):
        ctx.formod_batch([copy.deepcopy(pkg)])
    ok = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=4, z0=80.0, dz=2.0, seed=1)  # short paths: ~200 points
    ctx.formod_batch([ok])
    assert np.all(np.isfinite(ok.rad)) and ok.tau.max() <= 1.0
