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


# ---- the reference's binary table cache (src/jr_binary_tables_io.h) ------------------------------------------------------
def _blob_sha(jr, tables, ctl):
    import hashlib
    return hashlib.sha1(jr.core.tables_pack_host(tables, ctl.ng, ctl.nd).tobytes()).hexdigest()


def test_binary_cache_extents_and_name(jr):
    lib = jr.load_core()
    assert jr.core.binary_tables_filename(30, 100) == "bin.jurassic-fp32-tables-g30-p40-T30-u304-d100"
    for ND, NG in ((100, 30), (32, 5), (2, 5), (3, 1)):   # odd counts of int32 before the first double need padding
        assert lib.jrb_binary_tables_size(NG, 40, 30, 304, ND) == 16384 + C.sizeof(jr.abi.structs(ND, NG)[3])


def test_binary_cache_written_by_reference_is_read_natively(jr, refdrv, tmp_path, monkeypatch):
    """init_tbl with WRITE_BINARY=1 leaves bin.jurassic-... in the working directory; the native reader maps it and the
    packed device blob equals the one made from the ASCII files"""
    if not refdrv.reference_available(2, 5):
        pytest.skip("oracle/_ref not built")
    monkeypatch.chdir(tmp_path)
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl, skip_pairs=[(4, 1)])
    ctl.tblbase = jr.synth.write_ascii_tables(ctl, tbl, str(tmp_path), "boxcar")
    ref = refdrv.Reference(2, 5)
    c = ref.make_ctl(ctl)
    c.write_binary = 1
    ptr = ref.tables_from_files(c)
    ref.free_tbl(ptr)
    name = jr.core.binary_tables_filename(5, 2)
    assert os.path.getsize(name) == 16384 + C.sizeof(ref.tbl_t)
    ascii_tables = jr.core.read_ascii_tables(ctl, ctl.tblbase)
    mapped = jr.core.read_binary_tables(ctl, name)
    assert mapped.extents == dict(NG=5, TBLNP=40, TBLNT=30, TBLNU=304, ND=2)
    assert _blob_sha(jr, mapped, ctl) == _blob_sha(jr, ascii_tables, ctl)
    compact = mapped.compact(ctl.ng, ctl.nd)
    assert np.array_equal(compact.np, ascii_tables.np) and np.array_equal(compact.sr, ascii_tables.sr)
    # a sub-set of the gases/channels at the same indices is accepted (src/jr_binary_tables_io.h:151,170), anything else is not
    fewer = jr.Control(ctl.emitters[:3], ctl.nu[:1])
    jr.core.read_binary_tables(fewer, name).close()
    with pytest.raises(jr.JrbError, match="gas 0 is CO2"):
        jr.core.read_binary_tables(jr.Control(["H2O", "CO2"], ctl.nu), name)
    with pytest.raises(jr.JrbError, match="channel 1 is"):
        jr.core.read_binary_tables(jr.Control(ctl.emitters, [ctl.nu[0], ctl.nu[1] + 1.0]), name)
    with pytest.raises(jr.JrbError, match="fewer gases"):
        jr.core.read_binary_tables(jr.Control(list(ctl.emitters) + ["N2O"], ctl.nu), name)
    with open(name, "r+b") as f:
        f.truncate(os.path.getsize(name) - 8)
    with pytest.raises(jr.JrbError, match="truncated"):
        jr.core.read_binary_tables(ctl, name)
    with pytest.raises(jr.JrbError, match="cannot open"):
        jr.core.read_binary_tables(ctl, "no-such-file")
    mapped.close()


_REF_READS_CACHE = r"""
import ctypes as C, hashlib, importlib, sys
import numpy as np
import refdrv
jr = importlib.import_module("jurassic-gpu_b200")
ctl = jr.synth.control_limb_example()
ctl.tblbase = "/nonexistent/boxcar"              # no ASCII tables anywhere: only the binary cache can satisfy init_tbl
ref = refdrv.Reference(2, 5)
c = ref.make_ctl(ctl)
c.read_binary = 1                                # fatal inside the reference if the file is not accepted
ptr = ref.tables_from_files(c)
t = ref.tbl_t.from_address(ptr)
v = jr.abi.TblView()
v.dim_g, v.dim_p, v.dim_t, v.dim_u, v.dim_d, v.dim_s = 5, 40, 30, 304, 2, 1201
cast = lambda f, ty: C.cast(C.addressof(f), C.POINTER(ty))
v.np, v.nt, v.nu = cast(t.np, C.c_int32), cast(t.nt, C.c_int32), cast(t.nu, C.c_int32)
v.p, v.t, v.sr, v.st = cast(t.p, C.c_double), cast(t.t, C.c_double), cast(t.sr, C.c_double), cast(t.st, C.c_double)
v.u, v.eps = cast(t.u, C.c_float), cast(t.eps, C.c_float)
class Holder:
    def view(self): return v
print("SHA", hashlib.sha1(jr.core.tables_pack_host(Holder(), ctl.ng, ctl.nd).tobytes()).hexdigest())
"""


def test_binary_cache_written_natively_is_accepted_by_reference(jr, refdrv, tmp_path):
    """the other direction: a cache file from jrb_tables_write_binary satisfies the reference's init_tbl (READ_BINARY=1, no
    ASCII files present) and yields the same tables.  Runs the reference in a child process (it exits on rejection)."""
    import subprocess
    import sys
    if not refdrv.reference_available(2, 5):
        pytest.skip("oracle/_ref not built")
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl, skip_pairs=[(1, 0)])
    jr.core.write_binary_tables(os.path.join(str(tmp_path), jr.core.binary_tables_filename(5, 2)), tbl, ctl, NG=5, ND=2)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "oracle")]))
    out = subprocess.run([sys.executable, "-c", _REF_READS_CACHE], cwd=str(tmp_path), env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "matching binary tables file found" in out.stdout
    sha = [l.split()[1] for l in out.stdout.splitlines() if l.startswith("SHA ")][0]
    assert sha == _blob_sha(jr, tbl, ctl)


def test_binary_cache_round_trip_with_larger_extents(jr, tmp_path):
    """ng < NG, nd < ND, compact source arrays; the default extents give an 8.8 GB file that stays sparse on disk"""
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0, 950.0])
    tbl = jr.synth.make_tables(ctl)
    want = _blob_sha(jr, tbl, ctl)
    for ND, NG in ((8, 3), (100, 30)):
        path = os.path.join(str(tmp_path), jr.core.binary_tables_filename(NG, ND))
        jr.core.write_binary_tables(path, tbl, ctl, NG=NG, ND=ND)
        st = os.stat(path)
        assert st.st_size == 16384 + C.sizeof(jr.abi.structs(ND, NG)[3])
        assert st.st_blocks * 512 < 0.05 * st.st_size + 16e6   # holes where no table entry lives
        mapped = jr.core.read_binary_tables(ctl, path)
        assert _blob_sha(jr, mapped, ctl) == want
        mapped.close()
        os.remove(path)
    with pytest.raises(jr.JrbError, match="exceed"):
        jr.core.write_binary_tables(os.path.join(str(tmp_path), "x"), tbl, ctl, NG=1, ND=8)


def test_binary_cache_header_is_untrusted_input(jr, tmp_path):
    """advisor (round 1): header_size and the extents come from the file -- a header length other than the reference's
    16384, or extents whose products would overflow, are rejected instead of being turned into array offsets"""
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0])
    tbl = jr.synth.make_tables(ctl)
    good = os.path.join(str(tmp_path), jr.core.binary_tables_filename(3, 8))
    jr.core.write_binary_tables(good, tbl, ctl, NG=3, ND=8)
    jr.core.read_binary_tables(ctl, good).close()
    raw = open(good, "rb").read()
    head, body = raw[:16384], raw[16384:]

    def variant(name, old, new):
        assert old in head
        h = head.replace(old, new, 1)
        h = h[:16384].ljust(16384, b"\0")
        path = os.path.join(str(tmp_path), name)
        with open(path, "wb") as f:
            f.write(h + body)
        return path

    import re
    m = re.search(rb"header_size\s+16384", head)
    assert m
    with pytest.raises(jr.JrbError, match="header_size"):
        jr.core.read_binary_tables(ctl, variant("bad_header_size", m.group(0), m.group(0).replace(b"16384", b"16388")))
    m = re.search(rb"TBLNU\s+304", head)
    assert m
    with pytest.raises(jr.JrbError, match="4096|table_size|extents"):
        jr.core.read_binary_tables(ctl, variant("bad_extent", m.group(0), m.group(0).replace(b"304", b"99999999")))


@pytest.mark.gpu
def test_dropin_init_from_files_uses_and_writes_the_binary_cache(jr, oracle, tmp_path, monkeypatch):
    """jr_b200_init_from_files follows init_tbl's READ_BINARY / WRITE_BINARY protocol (src/jurassic.c:312-320, 669-671)"""
    import copy
    from helpers import assert_parity
    import refdrv
    monkeypatch.chdir(tmp_path)
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    ctl.tblbase = jr.synth.write_ascii_tables(ctl, tbl, str(tmp_path), "boxcar")
    pkg = jr.synth.example_package("limb", ctl)
    want = copy.deepcopy(pkg); oracle.formod(ctl, tbl, want)
    jr.load_core()
    lib = C.CDLL(os.path.join(ROOT, "jurassic-gpu_b200", "lib", "libjurassic_b200_dropin_nd2_ng5.so"))
    lib.jr_b200_init_from_files.argtypes = [C.c_void_p, C.c_int]
    lib.formod_GPU.argtypes = [C.c_void_p] * 3
    r = refdrv.Reference.__new__(refdrv.Reference)
    r.ctl_t, r.atm_t, r.obs_t, r.tbl_t = jr.abi.structs(2, 5)
    name = jr.core.binary_tables_filename(5, 2)
    for step in ("ascii, writes the cache", "cache only"):
        c, a, o = r.make_ctl(ctl, useGPU=1), r.make_atm(pkg), r.make_obs(pkg)
        if step.startswith("ascii"):
            c.read_binary, c.write_binary = -1, 1      # read_ctl's defaults: try the cache, fall back, then write it
            assert not os.path.exists(name)
        else:
            c.read_binary, c.write_binary = 1, 0
            c.tblbase = b"/nonexistent/boxcar"
        assert lib.jr_b200_init_from_files(C.addressof(c), 0) == 0
        assert os.path.exists(name)
        lib.formod_GPU(C.addressof(c), C.addressof(a), C.addressof(o))
        got = copy.deepcopy(pkg); r.read_obs(o, got)
        assert_parity(got, want, step)
        lib.jr_b200_finalize()
