"""The reference-facing drop-in layer: formod_GPU(ctl_t*, atm_t*, obs_t*) and the batched entry with the
reference's own struct layouts (include/jurassic_b200_dropin.h).  Needs a GPU."""
import copy
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_parity, run_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StructIO:
    """fills the reference's structs from the flat containers (same code as oracle/refdrv.py uses for the reference)"""

    def __init__(self, jr, refdrv, ND, NG):
        self.ND, self.NG = ND, NG
        self.ctl_t, self.atm_t, self.obs_t, self.tbl_t = jr.abi.structs(ND, NG)
        r = refdrv.Reference.__new__(refdrv.Reference)
        r.ND, r.NG = ND, NG
        r.ctl_t, r.atm_t, r.obs_t, r.tbl_t = self.ctl_t, self.atm_t, self.obs_t, self.tbl_t
        self.r = r


def _load_dropin(jr, ND, NG):
    jr.load_core()
    lib = C.CDLL(os.path.join(ROOT, "jurassic-gpu_b200", "lib", f"libjurassic_b200_dropin_nd{ND}_ng{NG}.so"))
    lib.jr_b200_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.formod_GPU.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.jr_b200_formod_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
    return lib


def _tbl_struct(io, tbl):
    buf = (C.c_char * C.sizeof(io.tbl_t))()  # zero-initialised by ctypes (lazily mapped pages)
    t = io.tbl_t.from_buffer(buf)
    g, P, T, U, d = tbl.dims
    np.ctypeslib.as_array(t.np)[:g, :d] = tbl.np
    np.ctypeslib.as_array(t.nt)[:g, :P, :d] = tbl.nt
    np.ctypeslib.as_array(t.nu)[:g, :P, :T, :d] = tbl.nu
    np.ctypeslib.as_array(t.p)[:g, :P, :d] = tbl.p
    np.ctypeslib.as_array(t.t)[:g, :P, :T, :d] = tbl.t
    np.ctypeslib.as_array(t.u)[:g, :P, :T, :U, :d] = tbl.u
    np.ctypeslib.as_array(t.eps)[:g, :P, :T, :U, :d] = tbl.eps
    np.ctypeslib.as_array(t.sr)[:, :d] = tbl.sr
    np.ctypeslib.as_array(t.st)[:] = tbl.st
    return t, buf


@pytest.mark.parametrize("case,ND,NG", [("limb", 2, 5), ("nadir", 3, 1)])
def test_formod_gpu_dropin(jr, refdrv, oracle, case, ND, NG):
    """formod_GPU with the reference's struct layouts == oracle; ND-wide reset; NaN mask; struct dims larger than ng/nd"""
    synth = jr.synth
    ctl = synth.control_limb_example() if case == "limb" else synth.control_nadir_example()
    tbl = synth.make_tables(ctl)
    pkg = synth.example_package(case, ctl)
    pkg.rad[1, 0] = np.nan
    io = StructIO(jr, refdrv, ND, NG)
    lib = _load_dropin(jr, ND, NG)
    c = io.r.make_ctl(ctl, useGPU=1)
    a = io.r.make_atm(pkg)
    o = io.r.make_obs(pkg)
    np.ctypeslib.as_array(o.rad)[:, :] = np.where(np.isnan(np.ctypeslib.as_array(o.rad)), np.nan, 7.0)  # stale values
    np.ctypeslib.as_array(o.tau)[:, :] = 7.0
    t, keep = _tbl_struct(io, tbl)
    assert lib.jr_b200_init(C.addressof(c), C.addressof(t), 0) == 0
    lib.formod_GPU(C.addressof(c), C.addressof(a), C.addressof(o))
    mine = copy.deepcopy(pkg)
    full_rad, full_tau = io.r.read_obs(o, mine)
    ref = run_oracle(oracle, ctl, tbl, [pkg])[0]
    assert_parity(mine, ref, f"dropin {case}")
    assert np.isnan(mine.rad[1, 0])
    lib.jr_b200_finalize()


def test_dropin_with_larger_struct_dims_and_batch(jr, refdrv, oracle):
    """ng < NG and nd < ND (dims 8x3 with 2 gases / 3 channels); batched entry == repeated formod_GPU"""
    ND, NG = 8, 3
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0, 950.0])
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=2, rays_per_profile=9, dz=7.0, seed=40 + i) for i in range(3)]
    io = StructIO(jr, refdrv, ND, NG)
    lib = _load_dropin(jr, ND, NG)
    c = io.r.make_ctl(ctl, useGPU=1)
    big = jr.Tables(NG, ND, 36, 12, 200)  # tables padded to the struct's gas/channel extents
    for name in ("np", "nt", "nu", "p", "t", "u", "eps"):
        getattr(big, name)[:ctl.ng, ..., :ctl.nd] = getattr(tbl, name)
    big.sr[:, :ctl.nd] = tbl.sr
    t, keep = _tbl_struct(io, big)
    assert lib.jr_b200_init(C.addressof(c), C.addressof(t), 0) == 0
    atms = [io.r.make_atm(p) for p in pkgs]
    obss = [io.r.make_obs(p) for p in pkgs]
    ap = (C.c_void_p * 3)(*[C.addressof(x) for x in atms])
    op = (C.c_void_p * 3)(*[C.addressof(x) for x in obss])
    lib.jr_b200_formod_batch(C.addressof(c), ap, op, 3)
    refs = run_oracle(oracle, ctl, tbl, pkgs)
    for i, p in enumerate(pkgs):
        mine = copy.deepcopy(p)
        full_rad, full_tau = io.r.read_obs(obss[i], mine)
        assert_parity(mine, refs[i], f"batch {i}")
        assert np.all(full_rad[:, ctl.nd:] == 0.0) and np.all(full_tau[:, ctl.nd:] == 1.0)  # ND-wide reset
        single = io.r.make_obs(p)
        lib.formod_GPU(C.addressof(c), C.addressof(atms[i]), C.addressof(single))
        assert np.array_equal(np.ctypeslib.as_array(single.rad), np.ctypeslib.as_array(obss[i].rad))
        assert np.array_equal(np.ctypeslib.as_array(single.tau), np.ctypeslib.as_array(obss[i].tau))
    lib.jr_b200_finalize()


def test_dropin_uses_reference_get_tbl_when_linked(jr, refdrv, oracle, tmp_path):
    """True drop-in wiring: the reference library is loaded with RTLD_GLOBAL, so formod_GPU obtains its tables from the
    reference's own get_tbl()/init_tbl() (ASCII files) exactly as src/GPUdrivers.cu:82 does."""
    if not refdrv.reference_available(2, 5):
        pytest.skip("oracle/_ref not built")
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    ctl.tblbase = jr.synth.write_ascii_tables(ctl, tbl, str(tmp_path), "boxcar")
    pkg = jr.synth.example_package("limb", ctl)
    ref = refdrv.Reference(2, 5, rtld_global=True)
    lib = _load_dropin(jr, 2, 5)
    c, a, o = ref.make_ctl(ctl, useGPU=1), ref.make_atm(pkg), ref.make_obs(pkg)
    lib.formod_GPU(C.addressof(c), C.addressof(a), C.addressof(o))
    mine = copy.deepcopy(pkg)
    ref.read_obs(o, mine)
    # CPU path of the reference on the same structs
    c0, o0 = ref.make_ctl(ctl, useGPU=0), ref.make_obs(pkg)
    ref.formod(c0, a, o0)
    r = copy.deepcopy(pkg)
    ref.read_obs(o0, r)
    assert_parity(mine, r, "dropin vs reference formod()")
    lib.jr_b200_finalize()


def test_concurrent_formod_gpu_callers_are_serialised(jr, refdrv, oracle):
    """several host threads calling formod_GPU at once (the reference supports this with lanes + omp critical,
    src/GPUdrivers.cu:275-335; here a mutex serialises them) all get correct results"""
    import threading
    ND, NG = 2, 5
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    pkgs = [jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=12, dz=5.0, seed=70 + i) for i in range(6)]
    io = StructIO(jr, refdrv, ND, NG)
    lib = _load_dropin(jr, ND, NG)
    c = io.r.make_ctl(ctl, useGPU=1)
    t, keep = _tbl_struct(io, tbl)
    assert lib.jr_b200_init(C.addressof(c), C.addressof(t), 0) == 0
    atms = [io.r.make_atm(p) for p in pkgs]
    obss = [io.r.make_obs(p) for p in pkgs]

    def work(i):
        for _ in range(3):
            lib.formod_GPU(C.addressof(c), C.addressof(atms[i]), C.addressof(obss[i]))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(pkgs))]
    [x.start() for x in th]
    [x.join() for x in th]
    refs = run_oracle(oracle, ctl, tbl, pkgs)
    for i, p in enumerate(pkgs):
        mine = copy.deepcopy(p)
        io.r.read_obs(obss[i], mine)
        assert_parity(mine, refs[i], f"thread {i}")
    lib.jr_b200_finalize()


def test_batched_jacobian_matches_reference_kernel(jr, refdrv, tmp_path):
    """Row f1: jr_b200_kernel (all perturbed forward models as device batches) == the reference's kernel()
    (src/jurassic.c:812-857, one formod() per state element on the CPU)"""
    if not refdrv.reference_available(2, 5):
        pytest.skip("oracle/_ref not built")
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    ctl.tblbase = jr.synth.write_ascii_tables(ctl, tbl, str(tmp_path), "boxcar")
    pkg = jr.synth.example_package("limb", ctl)
    ref = refdrv.Reference(2, 5, rtld_global=True)

    def structs(useGPU):
        c, a, o = ref.make_ctl(ctl, useGPU=useGPU), ref.make_atm(pkg), ref.make_obs(pkg)
        c.retp_zmin, c.retp_zmax, c.rett_zmin, c.rett_zmax = 6.0, 9.0, 10.0, 20.0  # 4 pressures, 11 temperatures
        for ig in range(5):
            c.retq_zmin[ig], c.retq_zmax[ig] = -999.0, -999.0
        c.retq_zmin[2], c.retq_zmax[2] = 15.0, 24.0                                 # 10 ozone values
        c.retq_zmin[1], c.retq_zmax[1] = 3.0, 5.0                                   # 3 water vapour values
        c.retk_zmin[0], c.retk_zmax[0] = 12.0, 13.0                                 # 2 extinction values
        np.ctypeslib.as_array(o.rad)[5, 1] = np.nan                                 # one masked measurement
        return c, a, o

    c, a, o = structs(0)
    k_ref = ref.kernel(c, a, o)
    assert k_ref.shape == (66 * 2 - 1, 4 + 11 + 10 + 3 + 2)

    lib = _load_dropin(jr, 2, 5)
    lib.jr_b200_kernel_dims.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
    lib.jr_b200_kernel_dims.restype = C.c_size_t
    lib.jr_b200_kernel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, jr.abi.c_double_p, C.c_size_t, C.c_size_t]
    c, a, o = structs(1)
    m = C.c_size_t()
    n = lib.jr_b200_kernel_dims(C.addressof(c), C.addressof(a), C.addressof(o), C.byref(m))
    assert (m.value, n) == k_ref.shape
    k = np.zeros((m.value, n))
    lib.jr_b200_kernel(C.addressof(c), C.addressof(a), C.addressof(o), k.ctypes.data_as(jr.abi.c_double_p), m.value, n)
    # finite differences amplify the 1e-11 forward-model differences by |y|/|dy| ~ 1e2..1e4
    scale = np.max(np.abs(k_ref), axis=0, keepdims=True)
    assert np.all(np.abs(k - k_ref) <= 1e-5 * np.abs(k_ref) + 1e-7 * scale), float(np.max(np.abs(k - k_ref) / (scale + 1e-300)))
    assert np.count_nonzero(k_ref) > 300  # the Jacobian is sparse: a ray only sees levels above its tangent point
    lib.jr_b200_finalize()


def test_formod_fov_batch_matches_reference(jr, refdrv, oracle):
    """Row f3: jr_b200_formod_fov_batch == the reference's formod(); formod_fov(); per package (shape file ctl->fov);
    formod_GPU itself keeps ignoring ctl->fov like the reference's formod()"""
    from test_oracle_vs_reference import FOV_SHAPE, fov_cases
    ND, NG = 2, 5
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    shape = jr.synth.read_tab(FOV_SHAPE)
    pkgs = [p for _, p in fov_cases(jr, ctl)]
    io = StructIO(jr, refdrv, ND, NG)
    lib = _load_dropin(jr, ND, NG)
    lib.jr_b200_formod_fov_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
    c = io.r.make_ctl(ctl, useGPU=1)
    c.fov = FOV_SHAPE.encode()
    t, keep = _tbl_struct(io, tbl)
    assert lib.jr_b200_init(C.addressof(c), C.addressof(t), 0) == 0
    atms = [io.r.make_atm(p) for p in pkgs]
    obss = [io.r.make_obs(p) for p in pkgs]
    ap = (C.c_void_p * len(pkgs))(*[C.addressof(x) for x in atms])
    op = (C.c_void_p * len(pkgs))(*[C.addressof(x) for x in obss])
    lib.jr_b200_formod_fov_batch(C.addressof(c), ap, op, len(pkgs))
    plain = run_oracle(oracle, ctl, tbl, pkgs)
    want = copy.deepcopy(plain)
    for i, p in enumerate(pkgs):
        assert oracle.formod_fov(want[i], shape[:, 0], shape[:, 1])
        mine = copy.deepcopy(p)
        io.r.read_obs(obss[i], mine)
        assert_parity(mine, want[i], f"fov batch {i}")
        if refdrv.reference_available(ND, NG):   # and the reference itself, same structs
            ref = refdrv.Reference(ND, NG)
            c0, a0, o0 = ref.make_ctl(ctl), ref.make_atm(p), ref.make_obs(p)
            c0.fov = FOV_SHAPE.encode()
            tp = ref.make_tbl(tbl)
            ref.formod_tbl(c0, a0, o0, tp)
            ref.formod_fov(c0, o0)
            r = copy.deepcopy(p)
            ref.read_obs(o0, r)
            ref.free_tbl(tp)
            assert_parity(mine, r, f"fov batch {i} vs reference")
        single = io.r.make_obs(p)                # plain entry point: no convolution although ctl->fov is set
        lib.formod_GPU(C.addressof(c), C.addressof(atms[i]), C.addressof(single))
        m0 = copy.deepcopy(p)
        io.r.read_obs(single, m0)
        assert_parity(m0, plain[i], f"formod_GPU ignores fov {i}")
    lib.jr_b200_finalize()
