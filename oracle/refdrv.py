"""TEST INFRASTRUCTURE ONLY -- Python drivers for the two CPU checkers:

  * Oracle     : oracle/libjr_oracle.so, the dimension-agnostic C restatement (oracle/jr_oracle.c)
  * Reference  : oracle/_ref/libjurassic_ref_nd<ND>_ng<NG>.so, the reference's own CPUdrivers.c + jurassic.c compiled
                 unmodified (oracle/Makefile); exists only where it was built from /root/reference (it travels to the
                 GPU box as a prebuilt file)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
import ctypes as C
import importlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_pkg = importlib.import_module("jurassic-gpu_b200")
abi = _pkg.abi


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p)


class Oracle:
    def __init__(self):
        path = os.path.join(_HERE, "libjr_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle oracle`")
        lib = C.CDLL(path)
        lib.jro_formod.argtypes = [C.POINTER(abi.CtlView), C.POINTER(abi.TblView), C.POINTER(abi.AtmView), C.POINTER(abi.ObsView)]
        lib.jro_formod.restype = C.c_int
        lib.jro_traceray.argtypes = [C.POINTER(abi.CtlView), C.POINTER(abi.AtmView), C.POINTER(abi.ObsView), C.c_int,
                                     abi.c_double_p, abi.c_double_p]
        lib.jro_traceray.restype = C.c_int
        lib.jro_max_threads.restype = C.c_int
        lib.jro_formod_fov.argtypes = [C.POINTER(abi.ObsView), C.c_int, C.c_int, abi.c_double_p, abi.c_double_p]
        lib.jro_formod_fov.restype = C.c_int
        lib.jro_intpol_atm_geo.argtypes = [C.POINTER(abi.CtlView), C.POINTER(abi.AtmView), C.c_double, C.c_double, C.c_double,
                                           abi.c_double_p]
        lib.jro_intpol_atm_geo.restype = C.c_int
        self.lib = lib

    def formod(self, ctl, tbl, pkg):
        cv, tv, av, ov = ctl.view(), tbl.view(), pkg.atm_view(), pkg.obs_view()
        rc = self.lib.jro_formod(C.byref(cv), C.byref(tv), C.byref(av), C.byref(ov))
        if rc == -2:
            raise RuntimeError("Too many LOS points!")  # where the reference's CPU path exits (src/jr_common.h:693-695)
        if rc == -3:
            raise RuntimeError("Cannot identify profiles. Check ordering of data points!")  # src/jurassic.c:727
        if rc == -4:
            raise RuntimeError("Distance of profiles is too large!")  # src/jurassic.c:728
        if rc != 0:
            raise RuntimeError("jro_formod failed")

    def formod_fov(self, pkg, dz, w):
        """field-of-view convolution of the results held in pkg, in place; False if the reference would abort"""
        ov = pkg.obs_view()
        dz, w = np.ascontiguousarray(dz, dtype=np.float64), np.ascontiguousarray(w, dtype=np.float64)
        return self.lib.jro_formod_fov(C.byref(ov), pkg.nd, len(dz), _dp(dz), _dp(w)) == 0

    def traceray(self, ctl, pkg, ir):
        """-> (los[np][6+nw+2ng] = z,lon,lat,p,t,ds,k..,q..,u.., tsurf); also updates pkg.tp*."""
        cv, av, ov = ctl.view(), pkg.atm_view(), pkg.obs_view()
        stride = 6 + ctl.nw + 2 * ctl.ng
        buf = np.zeros(abi.NLOS * stride)
        ts = C.c_double()
        n = self.lib.jro_traceray(C.byref(cv), C.byref(av), C.byref(ov), ir, _dp(buf), C.byref(ts))
        return buf[: n * stride].reshape(n, stride).copy(), ts.value

    def intpol_atm_geo(self, ctl, pkg, z, lon, lat):
        """intpol_atm_geo (ctl.ip = 1, 2, 3) over the whole atmosphere of pkg -> (rc, [p, t, q.., k..])"""
        cv, av = ctl.view(), pkg.atm_view()
        out = np.zeros(2 + ctl.ng + ctl.nw)
        rc = self.lib.jro_intpol_atm_geo(C.byref(cv), C.byref(av), z, lon, lat, _dp(out))
        return rc, out

    def threads(self):
        return self.lib.jro_max_threads()


def ref_lib_path(ND, NG, gpu=False):
    """gpu=True: the build that also holds the reference's own CUDA path (GPUdrivers.cu for sm_100a, `make refgpu`)"""
    return os.path.join(_HERE, "_ref", f"libjurassic_ref{'gpu' if gpu else ''}_nd{ND}_ng{NG}.so")


def reference_available(ND, NG):
    return os.path.exists(ref_lib_path(ND, NG))


class Reference:
    """The unmodified reference CPU path for one compile-time dimension set."""

    def __init__(self, ND, NG, rtld_global=False, gpu=False):
        path = ref_lib_path(ND, NG, gpu)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.ND, self.NG = ND, NG
        self.ctl_t, self.atm_t, self.obs_t, self.tbl_t = abi.structs(ND, NG)
        lib = C.CDLL(path, mode=C.RTLD_GLOBAL if rtld_global else C.RTLD_LOCAL)
        vp = C.c_void_p
        lib.jrref_tbl_calloc.restype = vp
        lib.jrref_tbl_free.argtypes = [vp]
        lib.jrref_init_tbl.argtypes = [vp, vp]
        lib.jrref_get_tbl.argtypes = [vp]
        lib.jrref_get_tbl.restype = vp
        lib.jrref_formod_tbl.argtypes = [vp, vp, vp, vp]
        lib.jrref_formod.argtypes = [vp, vp, vp]
        lib.jrref_traceray.argtypes = [vp, vp, vp, C.c_int, abi.c_double_p, abi.c_double_p]
        lib.jrref_traceray.restype = C.c_int
        lib.jrref_max_threads.restype = C.c_int
        lib.jrref_set_threads.argtypes = [C.c_int]
        lib.jrref_layout.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_longlong), C.c_int]
        lib.jrref_layout.restype = C.c_int
        lib.jrref_dims.argtypes = [C.POINTER(C.c_int)]
        lib.jrref_kernel_dims.argtypes = [vp, vp, vp, C.POINTER(C.c_size_t)]
        lib.jrref_kernel_dims.restype = C.c_size_t
        lib.jrref_kernel.argtypes = [vp, vp, vp, abi.c_double_p, C.c_size_t, C.c_size_t]
        lib.formod_fov.argtypes = [vp, vp]  # the reference's own public symbol (src/jurassic.c:214)
        lib.formod_fov.restype = None
        if hasattr(lib, "jrref_intpol_atm_geo"):
            lib.jrref_intpol_atm_geo.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, abi.c_double_p]
            lib.jrref_intpol_atm_geo.restype = None
        self.lib = lib

    # ---- ABI facts ----
    def dims(self):
        d = (C.c_int * 12)()
        self.lib.jrref_dims(d)
        return dict(zip(["ND", "NG", "NP", "NR", "NW", "NLOS", "TBLNP", "TBLNT", "TBLNU", "TBLNS", "LEN", "NSHAPE"], d))

    def layout(self):
        names = (C.c_char_p * 128)()
        vals = (C.c_longlong * 128)()
        n = self.lib.jrref_layout(names, vals, 128)
        return {names[i].decode(): vals[i] for i in range(n)}

    # ---- struct filling ----
    def make_ctl(self, ctl, useGPU=0):
        c = self.ctl_t()
        c.ng, c.nd, c.nw = ctl.ng, ctl.nd, ctl.nw
        for i, e in enumerate(ctl.emitters):
            c.emitter[i].value = e.encode()
        for i in range(ctl.nd):
            c.nu[i] = ctl.nu[i]
            c.window[i] = int(ctl.window[i])
        c.tblbase = ctl.tblbase.encode()
        c.hydz = ctl.hydz
        c.ctm_co2, c.ctm_h2o, c.ctm_n2, c.ctm_o2 = ctl.ctm_co2, ctl.ctm_h2o, ctl.ctm_n2, ctl.ctm_o2
        c.ip, c.refrac, c.rayds, c.raydz = ctl.ip, ctl.refrac, ctl.rayds, ctl.raydz
        c.cz, c.cx = getattr(ctl, "cz", 0.0), getattr(ctl, "cx", 0.0)
        c.write_bbt, c.formod, c.useGPU = ctl.write_bbt, ctl.formod, useGPU
        c.read_binary, c.write_binary = 0, 0
        c.fov = b"-"
        return c

    def make_atm(self, pkg):
        a = self.atm_t()
        n = pkg.n_atm
        a.np = n
        for name, src in (("time", pkg.atm_time), ("z", pkg.z), ("lon", pkg.lon), ("lat", pkg.lat), ("p", pkg.p), ("t", pkg.t)):
            np.ctypeslib.as_array(getattr(a, name))[:n] = src
        q = np.ctypeslib.as_array(a.q)
        k = np.ctypeslib.as_array(a.k)
        q[: pkg.ng, :n] = pkg.q[: pkg.ng]
        k[: pkg.nw, :n] = pkg.k[: pkg.nw]
        return a

    def make_obs(self, pkg):
        o = self.obs_t()
        n = pkg.n_rays
        o.nr = n
        for name in ("time", "obsz", "obslon", "obslat", "vpz", "vplon", "vplat"):
            np.ctypeslib.as_array(getattr(o, name))[:n] = getattr(pkg, name)
        np.ctypeslib.as_array(o.rad)[:n, : pkg.nd] = pkg.rad
        np.ctypeslib.as_array(o.tau)[:n, : pkg.nd] = pkg.tau
        return o

    def read_obs(self, o, pkg):
        """copy the outputs of obs_t back into pkg (all ND columns are returned as well)"""
        n = pkg.n_rays
        pkg.rad[:] = np.ctypeslib.as_array(o.rad)[:n, : pkg.nd]
        pkg.tau[:] = np.ctypeslib.as_array(o.tau)[:n, : pkg.nd]
        for name in ("tpz", "tplon", "tplat"):
            getattr(pkg, name)[:] = np.ctypeslib.as_array(getattr(o, name))[:n]
        return np.ctypeslib.as_array(o.rad)[:n].copy(), np.ctypeslib.as_array(o.tau)[:n].copy()

    def make_tbl(self, tbl):
        """calloc a tbl_t (lazily zero pages) and fill the populated part from a Tables container."""
        ptr = self.lib.jrref_tbl_calloc()
        if not ptr:
            raise MemoryError("tbl_t")
        t = self.tbl_t.from_address(ptr)
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
        return ptr

    def free_tbl(self, ptr):
        self.lib.jrref_tbl_free(ptr)

    def tables_from_files(self, c):
        """the reference's own ASCII loader (init_tbl) into a fresh tbl_t; returns the pointer"""
        ptr = self.lib.jrref_tbl_calloc()
        self.lib.jrref_init_tbl(C.addressof(c), ptr)
        return ptr

    # ---- compute ----
    def formod_tbl(self, c, a, o, tbl_ptr):
        self.lib.jrref_formod_tbl(C.addressof(c), C.addressof(a), C.addressof(o), tbl_ptr)

    def formod(self, c, a, o):
        self.lib.jrref_formod(C.addressof(c), C.addressof(a), C.addressof(o))

    def traceray(self, c, a, o, ir, ng, nw=1):
        stride = 6 + nw + 2 * ng
        buf = np.zeros(abi.NLOS * stride)
        ts = C.c_double()
        n = self.lib.jrref_traceray(C.addressof(c), C.addressof(a), C.addressof(o), ir, _dp(buf), C.byref(ts))
        return buf[: n * stride].reshape(n, stride).copy(), ts.value

    def intpol_atm_geo(self, c, a, z, lon, lat, ng, nw):
        """the reference's intpol_atm_geo (src/jurassic.c:685-691); a must be a fresh atm_t (init == 0) per atmosphere"""
        out = np.zeros(2 + ng + nw)
        self.lib.jrref_intpol_atm_geo(C.addressof(c), C.addressof(a), z, lon, lat, _dp(out))
        return out

    def formod_fov(self, c, o):
        """the reference's FOV convolution; NOTE it caches the shape file of the first call for the life of the process"""
        self.lib.formod_fov(C.addressof(c), C.addressof(o))

    def kernel(self, c, a, o):
        """the reference's finite-difference Jacobian kernel() (tables via get_tbl, i.e. from ctl.tblbase files)"""
        m = C.c_size_t()
        n = self.lib.jrref_kernel_dims(C.addressof(c), C.addressof(a), C.addressof(o), C.byref(m))
        k = np.zeros((m.value, n))
        self.lib.jrref_kernel(C.addressof(c), C.addressof(a), C.addressof(o), _dp(k), m.value, n)
        return k

    def threads(self):
        return self.lib.jrref_max_threads()
