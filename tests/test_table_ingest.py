"""Row f2: native ingest of the reference's ASCII emissivity tables / filter files versus the reference's own init_tbl
(src/jurassic.c:311-672).  CPU-only (the loader is host code); one GPU test runs the forward model on natively loaded
tables."""
import ctypes as C
import os

import numpy as np
import pytest


def _write_quirky_table(path):
    """a table file exercising the acceptance rules: comment and blank lines, a row with non-increasing eps (rejected,
    but it overwrites the current entry), a column longer than TBLNU = 304 (surplus rows ignored), T axes of different
    length per pressure level"""
    rows = ["# p T u eps", ""]
    for ip, p in enumerate([1.0, 10.0, 100.0]):
        for it, t in enumerate([200.0, 230.0, 260.0][: 2 + (ip % 2)]):
            n = 320 if (ip, it) == (1, 0) else 12
            for iu in range(n):
                u = 1e15 * 1.02 ** iu * (1 + ip)
                eps = 1.0 - np.exp(-1e-17 * u * (1 + 0.1 * it))
                rows.append("%.9g %.9g %.9g %.9g" % (p, t, u, eps))
                if (ip, it, iu) == (0, 1, 5):
                    rows.append("%.9g %.9g %.9g %.9g" % (p, t, u * 1.01, eps * 0.5))   # eps decreases -> not a new entry
                    rows.append("%.9g %.9g %.9g %.9g" % (p, t, u * 0.9, eps * 1.01))   # u decreases   -> not a new entry
            rows.append("")
    open(path, "w").write("\n".join(rows) + "\n")


def test_native_ingest_equals_reference_init_tbl(jr, refdrv, tmp_path):
    if not refdrv.reference_available(2, 5):
        pytest.skip("oracle/_ref not built")
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl, skip_pairs=[(4, 1)])                 # one missing file
    base = jr.synth.write_ascii_tables(ctl, tbl, str(tmp_path), "boxcar")
    _write_quirky_table(base + "_%.4f_%s.tab" % (ctl.nu[0], ctl.emitters[3]))  # replaces F11 @ 792
    ctl.tblbase = base
    mine = jr.core.read_ascii_tables(ctl, base)
    assert mine.n_missing == 1
    ref = refdrv.Reference(2, 5)
    c = ref.make_ctl(ctl)
    ptr = ref.tables_from_files(c)
    t = ref.tbl_t.from_address(ptr)
    g, P, T, U, d = mine.dims
    assert U == 304  # the over-long column is cut at TBLNU
    r_np = np.ctypeslib.as_array(t.np)[:g, :d]
    r_nt = np.ctypeslib.as_array(t.nt)[:g, :P, :d]
    r_nu = np.ctypeslib.as_array(t.nu)[:g, :P, :T, :d]
    assert np.array_equal(mine.np, r_np)
    # counts beyond the populated levels are whatever init_tbl's ++ pass leaves; compare the populated part
    for ig in range(g):
        for id_ in range(d):
            n_p = mine.np[ig, id_]
            assert np.array_equal(mine.nt[ig, :n_p, id_], r_nt[ig, :n_p, id_])
            assert np.array_equal(mine.p[ig, :n_p, id_], np.ctypeslib.as_array(t.p)[ig, :n_p, id_])
            for ip in range(n_p):
                n_t = mine.nt[ig, ip, id_]
                assert np.array_equal(mine.nu[ig, ip, :n_t, id_], r_nu[ig, ip, :n_t, id_])
                assert np.array_equal(mine.t[ig, ip, :n_t, id_], np.ctypeslib.as_array(t.t)[ig, ip, :n_t, id_])
                for it in range(n_t):
                    n_u = mine.nu[ig, ip, it, id_]
                    assert np.array_equal(mine.u[ig, ip, it, :n_u, id_], np.ctypeslib.as_array(t.u)[ig, ip, it, :n_u, id_])
                    assert np.array_equal(mine.eps[ig, ip, it, :n_u, id_], np.ctypeslib.as_array(t.eps)[ig, ip, it, :n_u, id_])
    assert np.allclose(mine.sr, np.ctypeslib.as_array(t.sr)[:, :d], rtol=1e-13)
    assert np.array_equal(mine.st, np.ctypeslib.as_array(t.st))
    ref.free_tbl(ptr)


def test_native_ingest_errors(jr, tmp_path):
    ctl = jr.synth.control_nadir_example()
    with pytest.raises(jr.JrbError, match="missing filter file"):   # the filter file is mandatory (read_shape -> mkFile exits)
        jr.core.read_ascii_tables(ctl, os.path.join(str(tmp_path), "nothing"))


@pytest.mark.gpu
def test_forward_model_on_natively_loaded_tables(jr, oracle, gpu_ctx_factory, tmp_path):
    import copy
    from helpers import assert_parity
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    base = jr.synth.write_ascii_tables(ctl, tbl, str(tmp_path), "boxcar")
    loaded = jr.core.read_ascii_tables(ctl, base)
    pkg = jr.synth.example_package("limb", ctl)
    ref = copy.deepcopy(pkg); oracle.formod(ctl, tbl, ref)
    ctx = gpu_ctx_factory()
    ctx.set_control(ctl); ctx.set_tables(loaded)
    mine = copy.deepcopy(pkg); ctx.formod_batch([mine])
    assert_parity(mine, ref, "native tables")
    # and through the drop-in layer: jr_b200_init_from_files(ctl_t*)
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = C.CDLL(os.path.join(ROOT, "jurassic-gpu_b200", "lib", "libjurassic_b200_dropin_nd2_ng5.so"))
    import refdrv
    r = refdrv.Reference.__new__(refdrv.Reference)
    r.ctl_t, r.atm_t, r.obs_t, r.tbl_t = jr.abi.structs(2, 5)
    ctl.tblbase = base
    c, a, o = r.make_ctl(ctl, useGPU=1), r.make_atm(pkg), r.make_obs(pkg)
    lib.jr_b200_init_from_files.argtypes = [C.c_void_p, C.c_int]
    assert lib.jr_b200_init_from_files(C.addressof(c), 0) == 0
    lib.formod_GPU.argtypes = [C.c_void_p] * 3
    lib.formod_GPU(C.addressof(c), C.addressof(a), C.addressof(o))
    got = copy.deepcopy(pkg); r.read_obs(o, got)
    assert_parity(got, ref, "drop-in with native tables")
    lib.jr_b200_finalize()
