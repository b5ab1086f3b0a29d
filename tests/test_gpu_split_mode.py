"""Split mode of the EGA step (jrb_ega_split.cu): gas-block passes + combine kernel.  Needs a GPU.

* latency mode: a single package is cut into one-gas blocks so that it fills the GPU; bit-identical to the fused kernel;
* many gases (30-gas refspec shape, example/refspec/template.ctl): blocks of 10 gases; equal to the fused kernel up to the
  rounding of the regrouped product, and within the 1e-6 tolerance of the oracle on a full 66-ray x 100-channel x 30-gas package.
"""
import copy
import os

import numpy as np
import pytest

from helpers import assert_parity, run_cuda, run_oracle

pytestmark = pytest.mark.gpu


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _same_bits(a, b, what):
    for name in ("rad", "tau"):
        assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), f"{what}: {name} differs"


def test_single_package_runs_split_and_is_bit_identical_to_fused(jr, oracle, gpu_ctx_factory):
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, seed=20240517)
    pkg.rad[100, 7] = np.nan
    ctx = gpu_ctx_factory()
    with env(JRB_NO_SPLIT=None, JRB_EGA_GAS_BLOCK=None):
        split = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
        st = ctx.stats()
        assert st["ega_gas_blocks"] == 5, st  # 1088 rays x 1 channel group: one gas per block
    with env(JRB_NO_SPLIT=1):
        fused = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
        assert ctx.stats()["ega_gas_blocks"] == 1
    _same_bits(split, fused, "one-gas blocks vs fused")
    assert np.isnan(split.rad[100, 7])
    assert fused.tau.min() < 1e-6  # opaque rays: early termination of a block is covered
    assert_parity(split, run_oracle(oracle, ctl, tbl, [pkg])[0], "split vs oracle")


@pytest.mark.parametrize("gpb", [2, 3])
def test_gas_blocks_of_several_gases(jr, gpu_ctx_factory, gpb):
    """blocks of 2 / 3 gases (the last block is shorter): regrouped product, equal to fused within rounding"""
    ctl = jr.synth.control_config_e()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.nadir_package(ctl, n_profiles=4, rays_per_profile=17, dlat=0.7, seed=20240518)
    ctx = gpu_ctx_factory()
    with env(JRB_NO_SPLIT=1):
        fused = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    with env(JRB_EGA_GAS_BLOCK=gpb):
        split = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
        assert ctx.stats()["ega_gas_blocks"] == (8 + gpb - 1) // gpb
    assert np.max(np.abs(split.rad - fused.rad) / np.abs(fused.rad)) < 1e-13
    assert np.max(np.abs(split.tau - fused.tau) / (np.abs(fused.tau) + 1e-300)) < 1e-13


def test_few_channel_instrument_in_split_mode(jr, oracle, gpu_ctx_factory):
    """nd = 2 (several rays per warp) in split mode"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.example_package("limb", ctl)
    ctx = gpu_ctx_factory()
    split = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    assert ctx.stats()["ega_gas_blocks"] == 5
    with env(JRB_NO_SPLIT=1):
        fused = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    _same_bits(split, fused, "nd=2 split vs fused")
    assert_parity(split, run_oracle(oracle, ctl, tbl, [pkg])[0], "nd=2 split vs oracle")


def test_refspec_shape_full_package(jr, oracle, gpu_ctx_factory):
    """example/refspec shape at full size: 66 limb rays x 100 channels x 30 gases, every gas with a realistic profile
    and a table for every (gas, channel) pair; window 2150.. cm^-1 (CO2, H2O and N2 continua)"""
    gases = ["CO2", "H2O", "O3", "N2O", "CH4", "CO", "HNO3", "SO2", "F11", "CCl4"] + [f"X{i}" for i in range(20)]
    ctl = jr.Control(gases, 2150.0 + np.arange(100))
    assert ctl.ctm_mask == 14
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=66, z0=3.0, dz=1.0, seed=77)  # Z0 3, Z1 68, DZ 1 (template.ctl)
    for ig in range(10, 30):  # the unnamed emitters: profiles of the named ones, scaled
        pkg.q[ig, :] = pkg.q[ig % 10, :] * (0.2 + 0.05 * ig)
    ctx = gpu_ctx_factory()
    mine = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    st = ctx.stats()
    assert st["ega_gas_blocks"] >= 3 and st["ega_kernel_variant"] == 1
    ref = run_oracle(oracle, ctl, tbl, [pkg])[0]
    assert_parity(mine, ref, "refspec 66 x 100 x 30")
    assert ref.tau.min() < 1e-3 and ref.tau.max() > 0.9  # from opaque to nearly transparent


def test_results_do_not_depend_on_how_a_batch_is_cut(jr, gpu_ctx_factory):
    """the order in which the gas factors are multiplied is a function of ng alone: a package alone (one gas per pass item)
    and the same package inside a large batch (fused kernel, or one pass per group of 10 gases) give the same bits --
    which is what makes sharding over devices bit-identical"""
    ctx = gpu_ctx_factory()
    # 5 gases: alone -> 5 one-gas blocks; in a batch of 8 -> fused
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, seed=900 + i) for i in range(8)]
    batch = run_cuda(ctx, ctl, tbl, pkgs, 1)
    assert ctx.stats()["ega_gas_blocks"] == 1
    alone = run_cuda(ctx, ctl, tbl, [pkgs[5]], 1)[0]
    assert ctx.stats()["ega_gas_blocks"] == 5
    _same_bits(alone, batch[5], "5 gases: alone vs in a batch")
    # 30 gases (3 product groups of 10): alone -> 30 one-gas blocks; in a batch -> one pass per group
    gases = ["CO2", "H2O", "O3", "N2O", "CH4", "CO", "HNO3", "SO2", "F11", "CCl4"] + [f"X{i}" for i in range(20)]
    ctl = jr.Control(gases, 2150.0 + np.arange(32))
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, seed=950 + i) for i in range(8)]
    for p in pkgs:
        p.q[10:, :] = 1e-9 * (1 + np.arange(20))[:, None]
    batch = run_cuda(ctx, ctl, tbl, pkgs, 1)
    assert ctx.stats()["ega_gas_blocks"] == 3
    alone = run_cuda(ctx, ctl, tbl, [pkgs[2]], 1)[0]
    assert ctx.stats()["ega_gas_blocks"] == 30
    _same_bits(alone, batch[2], "30 gases: alone vs in a batch")


def test_cooperative_tracer_is_bit_identical_to_thread_per_ray(jr, gpu_ctx_factory):
    """small batches walk a ray with a group of 8 lanes (the five (p,T) evaluations of a step side by side); same expressions,
    same bits: radiances, tangent points and every line-of-sight record; limb with refraction, observer inside the
    atmosphere, rejected rays, nadir (ground hit), refraction off"""
    ctx = gpu_ctx_factory()
    cases = []
    ctl = jr.synth.control_config_d()
    pkg = jr.synth.limb_package(ctl, n_profiles=3, rays_per_profile=21, dz=3.1, seed=31)
    pkg.obsz[3] = 40.0; pkg.vpz[3] = 10.0; pkg.vplat[3] = 3.0     # observer inside the atmosphere
    pkg.obsz[4] = -1.0                                             # rejected
    pkg.vpz[5] = 95.0                                              # view point above the atmosphere
    pkg.vpz[6] = -50.0; pkg.vplat[6] = 1.0                         # ground hit
    cases.append((ctl, pkg))
    ctl_e = jr.synth.control_config_e()
    cases.append((ctl_e, jr.synth.nadir_package(ctl_e, n_profiles=2, rays_per_profile=9, dlat=0.9, seed=32)))
    ctl_n = jr.Control(["CO2", "H2O"], [792.0, 832.0], refrac=0, rayds=8.0, raydz=1.0)
    cases.append((ctl_n, jr.synth.limb_package(ctl_n, n_profiles=1, rays_per_profile=10, dz=6.0, seed=33)))
    for ctl, pkg in cases:
        tbl = jr.synth.make_tables(ctl)
        coop = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
        los_c = [ctx.debug_los(r) for r in range(pkg.n_rays)]
        with env(JRB_NO_COOP_TRACER=1):
            plain = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
            los_p = [ctx.debug_los(r) for r in range(pkg.n_rays)]
        _same_bits(coop, plain, "cooperative tracer")
        for name in ("tpz", "tplon", "tplat"):
            assert np.array_equal(getattr(coop, name), getattr(plain, name)), name
        for (a, ta), (b, tb) in zip(los_c, los_p):
            assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True) and ta == tb  # (a record may hold an unwritten padding word)


def test_segment_tiled_kernel_is_bit_identical(jr, gpu_ctx_factory):
    """the segment-tiled form of the specialised kernel (jrb_ega_tiled.cuh: brackets and descriptors stay in registers across
    the segments of a tile) against the segment-by-segment kernel on a batch large enough to select it: Config D
    (opaque rays, rays of different length), a table set with missing pairs and unsorted columns, and a NaN-masked entry"""
    ctx = gpu_ctx_factory()
    ctl = jr.synth.control_config_d()
    for kind in ("plain", "holes"):
        tbl = jr.synth.make_tables(ctl, skip_pairs=[(3, 5), (4, 31), (1, 0)] if kind == "holes" else ())
        if kind == "holes":  # one column out of order in u: flagged at pack time, evaluated by plain bisection
            tbl.u[2, 7, 3, 10, 4], tbl.u[2, 7, 3, 11, 4] = tbl.u[2, 7, 3, 11, 4], tbl.u[2, 7, 3, 10, 4]
        pkgs = [jr.synth.limb_package(ctl, seed=1200 + i) for i in range(56)]  # 60 928 rays >= 16 rounds of 24 warps on 148 SMs
        pkgs[3].rad[40, 9] = np.nan
        with env(JRB_EGA_TILED=0):
            ref = run_cuda(ctx, ctl, tbl, pkgs, 1)
            assert ctx.stats()["ega_tiled"] == 0
        with env(JRB_EGA_TILED=1):
            til = run_cuda(ctx, ctl, tbl, pkgs, 1)
            assert ctx.stats()["ega_tiled"] == 1
        for a, b in zip(til, ref):
            _same_bits(a, b, f"tiled vs segment-by-segment ({kind})")
        assert min(p.tau.min() for p in ref) < 1e-6


def test_tiled_gas_block_pass_is_bit_identical(jr, gpu_ctx_factory):
    """latency mode with the segment-tiled pass kernel (default) against the segment-by-segment pass kernel and the fused one"""
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl, skip_pairs=[(2, 3)])
    pkg = jr.synth.limb_package(ctl, seed=77)
    ctx = gpu_ctx_factory()
    tiled = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    st = ctx.stats()
    assert st["ega_gas_blocks"] == 5 and st["ega_tiled"] == 1
    with env(JRB_EGA_TILED=0):
        plain = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
        assert ctx.stats()["ega_tiled"] == 0
    with env(JRB_NO_SPLIT=1, JRB_EGA_TILED=0):
        fused = run_cuda(ctx, ctl, tbl, [pkg], 1)[0]
    _same_bits(tiled, plain, "tiled pass vs plain pass")
    _same_bits(tiled, fused, "tiled pass vs fused")
