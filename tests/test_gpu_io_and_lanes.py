"""I/O modes, lanes and device groups of the runtime (include/jurassic_b200.h, second half).  Needs a GPU.

* staged vs direct I/O: page-locked caller memory (jrb_host_register) makes the device gather the inputs from, and store
  the results into, the caller's own arrays -- results must be bit-identical to the staged path;
* lanes: concurrent host threads are served by different contexts of one device (reference: src/GPUdrivers.cu:275-335);
* groups: the same batch over several devices gives the same bits -- runs when >= 2 GPUs are visible;
* advisor findings of round 1: nw = 0, few channels x many gases (shared-memory fit), stage failures leave no stale state.
"""
import copy
import ctypes as C
import os
import threading

import numpy as np
import pytest

from helpers import assert_parity, run_cuda, run_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits_equal(a, b, what):
    for name in ("rad", "tau", "tpz", "tplon", "tplat"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x, y, equal_nan=True), f"{what}: {name} differs"


def test_direct_io_is_bit_identical_to_staged(jr, oracle, gpu_ctx_factory):
    """registered packages: inputs gathered from / results stored into the caller's arrays; NaN mask; rejected rays"""
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=3, rays_per_profile=16, dz=4.0, seed=40 + i) for i in range(5)]
    pkgs[1].rad[3, 2] = np.nan
    pkgs[4].rad[7, 31] = np.inf
    pkgs[2].obsz[5] = -1.0  # rejected ray: rad 0, tau 1
    ctx = gpu_ctx_factory()
    staged = run_cuda(ctx, ctl, tbl, pkgs, 1)
    st = ctx.stats()
    assert st["io_direct"] == 0
    direct = [copy.deepcopy(p) for p in pkgs]
    try:
        for p in direct:
            jr.core.register_package(p)
        assert jr.load_core().jrb_host_is_registered(direct[0].rad.ctypes.data_as(C.c_void_p), direct[0].rad.nbytes) == 1
        ctx.formod_batch(direct)
        st = ctx.stats()
        assert st["io_direct"] == 1 and st["host_ms_scatter"] == 0.0
        for d, s in zip(direct, staged):
            _bits_equal(d, s, "direct vs staged")
        assert np.isnan(direct[1].rad[3, 2]) and np.isnan(direct[4].rad[7, 31])
        assert np.all(direct[2].rad[5] == 0.0) and np.all(direct[2].tau[5] == 1.0)
        # split API in direct mode: stage / run / fetch (fetch is a no-op, the results are in place)
        again = direct
        for p in again:
            p.rad[np.isfinite(p.rad)] = -1.0
            p.tau[:] = -1.0
        ctx.stage(again); ctx.run_staged(); ctx.fetch_staged(again)
        for d, s in zip(again, staged):
            _bits_equal(d, s, "direct split API")
        # a mixed batch (one package not registered) falls back to the staged path with the same bits
        mixed = again[:4] + [copy.deepcopy(pkgs[4])]
        ctx.formod_batch(mixed)
        assert ctx.stats()["io_direct"] == 0
        for d, s in zip(mixed, staged):
            _bits_equal(d, s, "mixed batch")
    finally:
        jr.core.host_unregister_all()
    ref = run_oracle(oracle, ctl, tbl, pkgs)
    for d, r in zip(staged, ref):
        assert_parity(d, r, "staged vs oracle")


def test_direct_io_with_struct_padding_and_generic_kernel(jr, oracle, gpu_ctx_factory):
    """row stride larger than nd (obs_t with ND > nd): the columns beyond nd are reset; generic kernel writes host rows too"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.example_package("limb", ctl)
    ctx = gpu_ctx_factory()
    ref = run_oracle(oracle, ctl, tbl, [pkg])[0]

    class Wide(jr.Package):  # rad/tau rows of 7 columns, only the first nd = 2 are channels
        def obs_view(self):
            v = super().obs_view()
            v.rad = self.wide_rad.ctypes.data_as(jr.abi.c_double_p)
            v.tau = self.wide_tau.ctypes.data_as(jr.abi.c_double_p)
            v.row_stride, v.nd_reset = 7, 7
            return v

    for variant in (1, 0):
        w = copy.deepcopy(pkg)
        w.__class__ = Wide
        w.wide_rad = np.full((w.n_rays, 7), 5.0)
        w.wide_tau = np.full((w.n_rays, 7), 5.0)
        w.wide_rad[4, 1] = np.nan
        try:
            jr.core.register_package(w)
            jr.core.host_register(w.wide_rad); jr.core.host_register(w.wide_tau)
            ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(variant)
            ctx.formod_batch([w])
            assert ctx.stats()["io_direct"] == 1
        finally:
            jr.core.host_unregister_all()
        assert np.all(w.wide_rad[:, 2:] == 0.0) and np.all(w.wide_tau[:, 2:] == 1.0)
        assert np.isnan(w.wide_rad[4, 1])
        w.wide_rad[4, 1] = ref.rad[4, 1]
        w.rad, w.tau = w.wide_rad[:, :2].copy(), w.wide_tau[:, :2].copy()
        assert_parity(w, ref, f"wide rows variant {variant}")


def test_lanes_serve_concurrent_callers(jr, oracle):
    """4 host threads x 6 calls of one package each through a 1-device / 4-lane group == sequential results"""
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=20, dz=3.0, seed=70 + i) for i in range(24)]
    g = jr.Group(ndev=1, nlanes=4)
    try:
        g.set_tables(ctl, tbl)
        seq = [copy.deepcopy(p) for p in pkgs]
        for p in seq:
            g.formod_batch([p], ctl)
        par = [copy.deepcopy(p) for p in pkgs]
        errs = []

        def worker(t):
            try:
                for i in range(t, len(par), 4):
                    g.formod_batch([par[i]], ctl)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        for a, b in zip(par, seq):
            _bits_equal(a, b, "lanes")
        used = sum(1 for l in range(4) if g.context_stats(0, l)["n_rays"] > 0)
        assert used >= 2, "concurrent callers were not spread over lanes"
    finally:
        g.close()
    ref = run_oracle(oracle, ctl, tbl, pkgs[:2])
    for a, r in zip(seq[:2], ref):
        assert_parity(a, r, "lanes vs oracle")


def test_control_travels_with_the_call(jr, oracle):
    """two callers with different control values (write_bbt) share a group: each gets its own semantics"""
    ctl_a = jr.synth.control_nadir_example()
    ctl_b = copy.deepcopy(ctl_a)
    ctl_b.write_bbt = 0 if ctl_a.write_bbt else 1
    tbl = jr.synth.make_tables(ctl_a)
    pkg = jr.synth.example_package("nadir", ctl_a)
    g = jr.Group(ndev=1, nlanes=2)
    try:
        g.set_tables(ctl_a, tbl)
        a, b, a2 = copy.deepcopy(pkg), copy.deepcopy(pkg), copy.deepcopy(pkg)
        g.formod_batch([a], ctl_a); g.formod_batch([b], ctl_b); g.formod_batch([a2], ctl_a)
    finally:
        g.close()
    assert_parity(a, run_oracle(oracle, ctl_a, tbl, [pkg])[0], "ctl a")
    assert_parity(b, run_oracle(oracle, ctl_b, tbl, [pkg])[0], "ctl b")
    _bits_equal(a, a2, "ctl a again")


def test_group_over_several_devices_is_bit_identical(jr):
    """1 device vs all visible devices, same bits; tables reach the other devices by ncclBroadcast"""
    ndev = jr.load_core().jrb_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, seed=300 + i) for i in range(4 * ndev + 1)]
    g1 = jr.Group(ndev=1, nlanes=1)
    g1.set_tables(ctl, tbl)
    one = [copy.deepcopy(p) for p in pkgs]
    g1.formod_batch(one, ctl)
    g1.close()
    gn = jr.Group(ndev=ndev, nlanes=2)
    try:
        gn.set_tables(ctl, tbl)
        st = gn.stats()
        assert st["nccl_nranks"] == ndev and st["table_bytes"] > 0
        many = [copy.deepcopy(p) for p in pkgs]
        gn.formod_batch(many, ctl)
        assert gn.stats()["n_slices"] == ndev
        for d in range(ndev):
            assert gn.context_stats(d, 0)["n_packages"] >= 4
        for a, b in zip(many, one):
            _bits_equal(a, b, f"{ndev} devices vs 1")
    finally:
        gn.close()


def test_nw_zero_means_no_extinction(jr, oracle, gpu_ctx_factory):
    """advisor (round 1): with nw = 0 the extinction coefficient is 0 (reference: los->k[0] stays 0, src/jr_common.h:591)"""
    ctl = jr.Control(["CO2", "H2O", "O3"], [792.0, 832.0, 900.0], nw=0)
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=8, dz=6.0, seed=11)
    ref = run_oracle(oracle, ctl, tbl, [pkg])[0]
    assert ref.tau.max() > 0.5  # extinction read from a column-density slot would drive tau to 0
    ctx = gpu_ctx_factory()
    for v in (1, 0):
        assert_parity(run_cuda(ctx, ctl, tbl, [pkg], v)[0], ref, f"nw=0 variant {v}")


def test_few_channels_many_gases_fit_or_fall_back(jr, oracle, gpu_ctx_factory):
    """advisor (round 1): nd = 1 packs 32 rays into a warp; with 25 gases the per-warp state no longer fits the largest CTA.
    The launcher shrinks the CTA; when even one warp does not fit, the generic kernel is chosen -- never a launch failure."""
    ctl = jr.Control([f"G{i}" for i in range(25)], [900.0])
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=20, dz=3.0, seed=21)
    pkg.q[:, :] = 3e-9
    ref = run_oracle(oracle, ctl, tbl, [pkg])[0]
    ctx = gpu_ctx_factory()
    mine = run_cuda(ctx, ctl, tbl, [pkg], -1)[0]
    assert_parity(mine, ref, "nd=1 ng=25")
    # per-gas (p,T) grids make the staged records 4 doubles per gas larger: 30 gases x 32 rays per warp
    ctl2 = jr.Control([f"G{i}" for i in range(30)], [900.0])
    tbl2 = jr.synth.make_tables(ctl2)
    tbl2.p[1::2] *= 1.03  # every other gas on its own pressure axis
    pkg2 = jr.synth.limb_package(ctl2, n_profiles=1, rays_per_profile=12, dz=5.0, seed=22)
    pkg2.q[:, :] = 3e-9
    ref2 = run_oracle(oracle, ctl2, tbl2, [pkg2])[0]
    mine2 = run_cuda(ctx, ctl2, tbl2, [pkg2], -1)[0]
    assert_parity(mine2, ref2, "nd=1 ng=30 per-gas grids")


def test_failed_stage_leaves_nothing_staged(jr, gpu_ctx_factory):
    """advisor (round 1): a failing jrb_stage after a successful one must not leave the old batch runnable"""
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    pkg = jr.synth.example_package("limb", ctl)
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl)
    ctx.stage([copy.deepcopy(pkg)]); ctx.run_staged()
    bad = copy.deepcopy(pkg)
    bad.z = bad.z[:0]  # empty atmosphere
    bad.atm_time, bad.lon, bad.lat, bad.p, bad.t = (x[:0] for x in (bad.atm_time, bad.lon, bad.lat, bad.p, bad.t))
    with pytest.raises(jr.JrbError):
        ctx.stage([bad])
    with pytest.raises(jr.JrbError, match="nothing staged"):
        ctx.run_staged()
    with pytest.raises(jr.JrbError):
        ctx.fetch_staged([copy.deepcopy(pkg)])


def test_dropin_pinned_packages_and_concurrent_formod_gpu(jr, refdrv, oracle):
    """jr_b200_pin_packages: direct I/O on the reference's own structs; formod_GPU from 3 host threads (lanes)"""
    ND, NG = 32, 5
    ctl = jr.synth.control_config_d()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=24, dz=2.5, seed=90 + i) for i in range(6)]
    from test_gpu_dropin import StructIO, _load_dropin, _tbl_struct
    io = StructIO(jr, refdrv, ND, NG)
    lib = _load_dropin(jr, ND, NG)
    lib.jr_b200_pin_packages.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
    lib.jr_b200_core_context.restype = C.c_void_p
    c = io.r.make_ctl(ctl, useGPU=1)
    t, keep = _tbl_struct(io, tbl)
    assert lib.jr_b200_init(C.addressof(c), C.addressof(t), 0) == 0
    atms = [io.r.make_atm(p) for p in pkgs]
    obss = [io.r.make_obs(p) for p in pkgs]
    ap = (C.c_void_p * len(pkgs))(*[C.addressof(x) for x in atms])
    op = (C.c_void_p * len(pkgs))(*[C.addressof(x) for x in obss])
    ref = run_oracle(oracle, ctl, tbl, pkgs)
    try:
        assert lib.jr_b200_pin_packages(ap, op, len(pkgs)) == 0
        lib.jr_b200_formod_batch(C.addressof(c), ap, op, len(pkgs))
        core = C.c_void_p(lib.jr_b200_core_context())
        st = jr.abi.Stats()
        jr.load_core().jrb_get_stats(core, C.byref(st))
        assert st.io_direct == 1
        for p, o, r in zip(pkgs, obss, ref):
            m = copy.deepcopy(p)
            io.r.read_obs(o, m)
            assert_parity(m, r, "dropin direct")
        # concurrent single-package callers
        for o in obss:
            np.ctypeslib.as_array(o.rad)[:, :] = 0.0
        errs = []

        def worker(tid):
            try:
                for i in range(tid, len(pkgs), 3):
                    lib.formod_GPU(C.addressof(c), C.addressof(atms[i]), C.addressof(obss[i]))
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        assert not errs
        for p, o, r in zip(pkgs, obss, ref):
            m = copy.deepcopy(p)
            io.r.read_obs(o, m)
            assert_parity(m, r, "dropin concurrent formod_GPU")
    finally:
        lib.jr_b200_finalize()
        jr.core.host_unregister_all()
