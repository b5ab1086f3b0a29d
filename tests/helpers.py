"""Shared helpers of the test-suite: tolerances of SURVEY.md section 8c and small case builders."""
import copy

import numpy as np

# north-star tolerance: <= 1e-6 relative in double on rad and tau.  The only discontinuity of the recurrence is
# the opaque cut-off tau_path < 1e-9 (src/jr_common.h:239), below which the *relative* tau error is unbounded
# while |tau| <~ 1e-9; hence the absolute floors (SURVEY.md 8c "Tolerance").
RTOL = 1e-6
TAU_FLOOR = 1e-12
RAD_FLOOR_REL = 1e-12
TP_ATOL = 1e-9  # tangent point outputs [km / deg]


def assert_parity(mine, ref, what="", rtol=RTOL, bbt=False):
    """mine/ref: Package-like objects with rad, tau, tpz, tplon, tplat."""
    rad_m, rad_r = np.asarray(mine.rad), np.asarray(ref.rad)
    nan_m, nan_r = np.isnan(rad_m), np.isnan(rad_r)
    assert np.array_equal(nan_m, nan_r), f"{what}: NaN mask differs"
    ok = ~nan_r
    floor = RAD_FLOOR_REL * (np.max(np.abs(rad_r[ok])) if ok.any() else 0.0)
    err_rad = np.abs(rad_m[ok] - rad_r[ok])
    assert np.all(err_rad <= rtol * np.abs(rad_r[ok]) + floor), \
        f"{what}: rad max rel err {np.max(err_rad / (np.abs(rad_r[ok]) + 1e-300)):.3e}"
    err_tau = np.abs(mine.tau - ref.tau)
    assert np.all(err_tau <= rtol * np.abs(ref.tau) + TAU_FLOOR), \
        f"{what}: tau max rel err {np.max(err_tau / (np.abs(ref.tau) + 1e-300)):.3e}"
    for name in ("tpz", "tplon", "tplat"):
        a, b = getattr(mine, name), getattr(ref, name)
        assert np.all(np.abs(a - b) <= TP_ATOL), f"{what}: {name} max abs err {np.max(np.abs(a - b)):.3e}"


def max_rel(a, b, floor=0.0):
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor + 1e-300)))


def run_oracle(oracle, ctl, tbl, pkgs):
    out = []
    for p in pkgs:
        q = copy.deepcopy(p)
        oracle.formod(ctl, tbl, q)
        out.append(q)
    return out


def run_cuda(ctx, ctl, tbl, pkgs, variant=-1):
    ctx.set_control(ctl)
    ctx.set_tables(tbl)
    ctx.set_kernel_variant(variant)
    out = [copy.deepcopy(p) for p in pkgs]
    ctx.formod_batch(out)
    return out
