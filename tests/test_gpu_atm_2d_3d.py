"""2-D / 3-D atmosphere interpolation (ctl->ip = 2, 3; src/jurassic.c:685-804) on the CUDA path -- SURVEY.md row f3.

The reference's formod() stops at an assert for ip != 1 (src/jr_common.h:573,581), so there is no reference behaviour of the
whole call.  The oracle composes the reference's tracer with the reference's intpol_atm_geo dispatch (oracle/jr_oracle.c);
the interpolation itself is pinned bit for bit against the reference's own intpol_atm_geo on the CPU
(tests/test_oracle_vs_reference.py::test_intpol_atm_geo_restatement_equals_reference).  Here: CUDA against that oracle,
same tolerance as everywhere (1e-6 relative on rad / tau, 1e-9 on the tangent point).
"""
import copy

import numpy as np
import pytest

from helpers import assert_parity, run_cuda, run_oracle

pytestmark = pytest.mark.gpu


def _case(jr, ip, nd=None, **kw):
    synth = jr.synth
    ctl = synth.control_limb_example() if nd is None else synth.control_config_d(nd=nd)
    ctl.ip = ip
    ctl.cz, ctl.cx = 1.7, 420.0
    return ctl, synth.make_tables(ctl), synth.track_package(ctl, **kw)


@pytest.mark.parametrize("ip", [2, 3])
def test_track_atmosphere_matches_oracle(jr, oracle, gpu_ctx_factory, ip):
    ctl, tbl, pkg = _case(jr, ip)
    ref = run_oracle(oracle, ctl, tbl, [pkg])
    ctx = gpu_ctx_factory()
    for variant in (1, 0):
        mine = run_cuda(ctx, ctl, tbl, [pkg], variant)
        assert ctx.stats()["ega_kernel_variant"] == variant
        assert_parity(mine[0], ref[0], f"ip={ip} variant {variant}")
    assert np.all(np.isfinite(ref[0].rad)) and ref[0].rad.max() > 0
    # the columns matter: the 1-D result of the first column alone is different
    c1 = copy.deepcopy(ctl)
    c1.ip = 1
    nz = pkg.n_atm // 9
    one = copy.deepcopy(pkg)
    for name in ("atm_time", "z", "lon", "lat", "p", "t"):
        setattr(one, name, getattr(one, name)[:nz].copy())
    one.q, one.k = one.q[:, :nz].copy(), one.k[:, :nz].copy()
    flat = run_cuda(ctx, c1, tbl, [one], 1)
    assert np.max(np.abs(flat[0].rad - ref[0].rad) / ref[0].rad) > 1e-3


@pytest.mark.parametrize("ip", [2, 3])
def test_track_atmosphere_los_records(jr, oracle, gpu_ctx_factory, ip):
    """every point of the line of sight: altitude, p, T, extinction, column densities against the oracle's tracer"""
    ctl, tbl, pkg = _case(jr, ip)
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl); ctx.set_kernel_variant(1)
    ctx.formod_batch([copy.deepcopy(pkg)])
    ng, nw = ctl.ng, ctl.nw
    for ir in (0, 7, 23):
        los_o, ts_o = oracle.traceray(ctl, copy.deepcopy(pkg), ir)
        rec, ts = ctx.debug_los(ir)
        assert rec.shape[0] == los_o.shape[0] and ts == pytest.approx(ts_o, rel=1e-12)
        u0, tail = 4 + nw, rec.shape[1] - 6
        np.testing.assert_allclose(rec[:, tail], los_o[:, 0], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(rec[:, 0:2], los_o[:, 3:5], rtol=1e-9)
        np.testing.assert_allclose(rec[:, 4:4 + nw], los_o[:, 6:6 + nw], rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(rec[:-2, u0:u0 + ng], los_o[:-2, 6 + nw + ng:], rtol=1e-8)


@pytest.mark.parametrize("ip", [2, 3])
def test_track_atmosphere_batch_of_packages_32_channels(jr, oracle, gpu_ctx_factory, ip):
    """Config-D control (32 channels x 5 gases), three packages with different tracks in one call; time-keyed slices: the
    second half of each atmosphere carries another time stamp and must be ignored by the rays"""
    ctl = jr.synth.control_config_d(nd=32)
    ctl.ip = ip
    ctl.cz, ctl.cx = 2.2, 480.0
    tbl = jr.synth.make_tables(ctl)
    pkgs = []
    for i in range(3):
        a = jr.synth.track_package(ctl, n_profiles=7, rays=10, lat0=-10.0 + i, dlat=4.0, z0=8.0 + i, dz=3.0, seed=77 + i)
        b = jr.synth.track_package(ctl, n_profiles=4, rays=1, lat0=-3.0, dlat=2.0, seed=177 + i)
        both = jr.Package(ctl.ng, ctl.nw, ctl.nd, a.n_atm + b.n_atm, a.n_rays)
        for name in ("atm_time", "z", "lon", "lat", "p", "t"):
            getattr(both, name)[:] = np.concatenate([getattr(a, name), getattr(b, name)])
        both.atm_time[a.n_atm:] = 1.0
        both.q[:] = np.concatenate([a.q, b.q], axis=1)
        both.k[:] = np.concatenate([a.k, b.k], axis=1)
        for name in ("time", "obsz", "obslon", "obslat", "vpz", "vplon", "vplat"):
            getattr(both, name)[:] = getattr(a, name)
        pkgs.append(both)
    ref = run_oracle(oracle, ctl, tbl, pkgs)
    only = run_oracle(oracle, ctl, tbl, [jr.synth.track_package(ctl, n_profiles=7, rays=10, lat0=-10.0, dlat=4.0, z0=8.0, dz=3.0, seed=77)])
    assert np.array_equal(only[0].rad, ref[0].rad)  # the other time slice does not matter
    ctx = gpu_ctx_factory()
    mine = run_cuda(ctx, ctl, tbl, pkgs, 1)
    for m, r in zip(mine, ref):
        assert_parity(m, r, f"ip={ip} batch")
    again = run_cuda(ctx, ctl, tbl, pkgs, 1)
    for m, r in zip(mine, again):
        assert np.array_equal(m.rad, r.rad) and np.array_equal(m.tau, r.tau)


def test_2d_profile_list_errors_and_unknown_ip(jr, oracle, gpu_ctx_factory):
    """the fatal conditions of intpol_atm_2d (src/jurassic.c:727-728) fail the call with the reference's message; afterwards
    the context works; ip outside 1..3 is rejected like the dispatch does (:690)"""
    ctl, tbl, pkg = _case(jr, 2)
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl)
    nz = pkg.n_atm // 9
    bad = copy.deepcopy(pkg)
    bad.lat[4 * nz:] += 15.0
    with pytest.raises(jr.JrbError, match="Distance of profiles is too large"):
        ctx.formod_batch([bad])
    with pytest.raises(RuntimeError, match="Distance of profiles is too large"):
        oracle.formod(ctl, tbl, copy.deepcopy(bad))
    bad = copy.deepcopy(pkg)
    bad.lat[5] += 0.01
    with pytest.raises(jr.JrbError, match="Cannot identify profiles"):
        ctx.formod_batch([bad])
    good = run_cuda(ctx, ctl, tbl, [pkg], 1)
    assert_parity(good[0], run_oracle(oracle, ctl, tbl, [pkg])[0], "after the errors")
    c9 = copy.deepcopy(ctl)
    c9.ip = 4
    with pytest.raises(jr.JrbError, match="Unknown interpolation method"):
        ctx.set_control(c9)
    c3 = copy.deepcopy(ctl)
    c3.ip, c3.cz, c3.cx = 3, 0.0, 0.0
    with pytest.raises(jr.JrbError, match="CZ > 0"):
        ctx.set_control(c3)


def test_3d_empty_influence_sphere(jr, oracle, gpu_ctx_factory):
    """3-D form with influence radii so small that points of the ray see no data: p = T = NaN (src/jurassic.c:799-803), the
    ray never leaves the atmosphere and the reference's tracer runs into "Too many LOS points!" -- same here"""
    ctl, tbl, pkg = _case(jr, 3)
    ctl.cz, ctl.cx = 0.2, 30.0
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(tbl)
    with pytest.raises(RuntimeError, match="Too many LOS points"):
        oracle.formod(ctl, tbl, copy.deepcopy(pkg))
    with pytest.raises(jr.JrbError, match="Too many LOS points"):
        ctx.formod_batch([copy.deepcopy(pkg)])


def test_dropin_formod_gpu_with_2d_atmosphere(jr, oracle, refdrv):
    """formod_GPU(ctl_t*, atm_t*, obs_t*) of the drop-in library with ctl->ip = 2 (the reference's own struct layouts)"""
    import ctypes as C

    from test_gpu_dropin import StructIO, _load_dropin, _tbl_struct
    ND, NG = 2, 5
    ctl, tbl, pkg = _case(jr, 2)
    io = StructIO(jr, refdrv, ND, NG)
    lib = _load_dropin(jr, ND, NG)
    c, a, o = io.r.make_ctl(ctl, useGPU=1), io.r.make_atm(pkg), io.r.make_obs(pkg)
    assert c.ip == 2
    t, keep = _tbl_struct(io, tbl)
    assert lib.jr_b200_init(C.addressof(c), C.addressof(t), 0) == 0
    lib.formod_GPU(C.addressof(c), C.addressof(a), C.addressof(o))
    out = copy.deepcopy(pkg)
    io.r.read_obs(o, out)
    lib.jr_b200_finalize()
    assert_parity(out, run_oracle(oracle, ctl, tbl, [pkg])[0], "drop-in ip=2")
