"""ctypes binding of the core C ABI (include/jurassic_b200.h) plus dimension-agnostic host containers.

This is harness glue for tests and bench.py: all computing happens in libjurassic_b200.so (hand-written CUDA for
sm_100a).  There is deliberately no CPU fallback -- a missing library or missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.environ.get("JRB_LIBDIR") or os.path.join(_HERE, "lib")  # override only for kernel-variant experiments
CORE_LIB = os.path.join(LIBDIR, "libjurassic_b200.so")

_lib = None


class JrbError(RuntimeError):
    pass


def load_core():
    """dlopen the core library and declare every symbol of include/jurassic_b200.h (no compute is triggered)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(CORE_LIB):
        raise JrbError(f"{CORE_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the CUDA extension is mandatory, there is no fallback path)")
    lib = C.CDLL(CORE_LIB, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    lib.jrb_version.restype = C.c_char_p
    lib.jrb_device_count.restype = C.c_int
    lib.jrb_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.jrb_destroy.argtypes = [vp]
    lib.jrb_destroy.restype = None
    lib.jrb_last_error.argtypes = [vp]
    lib.jrb_last_error.restype = C.c_char_p
    lib.jrb_set_control.argtypes = [vp, C.POINTER(abi.CtlView)]
    lib.jrb_set_tables.argtypes = [vp, C.POINTER(abi.TblView)]
    lib.jrb_tables_pack_info.argtypes = [C.POINTER(abi.TblView), C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                         abi.c_int_p, abi.c_int_p, C.POINTER(C.c_ulonglong), abi.c_int_p]
    lib.jrb_tables_pack_info.restype = C.c_int
    lib.jrb_tables_pack_host.argtypes = [C.POINTER(abi.TblView), C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.jrb_tables_pack_host.restype = C.c_int
    lib.jrb_tables_upload_blob.argtypes = [vp, C.c_void_p, C.c_size_t]
    lib.jrb_tables_upload_blob.restype = C.c_int
    lib.jrb_tables_read_ascii.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, abi.c_double_p, C.c_int, C.c_int,
                                          C.c_int, C.POINTER(vp)]
    lib.jrb_tables_read_ascii.restype = C.c_int
    lib.jrb_host_tables_view.argtypes = [vp, C.POINTER(abi.TblView), abi.c_int_p]
    lib.jrb_host_tables_view.restype = C.c_int
    lib.jrb_host_tables_free.argtypes = [vp]
    lib.jrb_host_tables_free.restype = None
    lib.jrb_ingest_last_error.restype = C.c_char_p
    lib.jrb_binary_tables_filename.argtypes = [C.c_char_p, C.c_size_t] + [C.c_int] * 5
    lib.jrb_binary_tables_size.argtypes = [C.c_int] * 5
    lib.jrb_binary_tables_size.restype = C.c_size_t
    lib.jrb_tables_read_binary.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, abi.c_double_p, C.POINTER(vp)]
    lib.jrb_tables_read_binary.restype = C.c_int
    lib.jrb_tables_write_binary.argtypes = [C.c_char_p, C.POINTER(abi.TblView), C.c_int, C.POINTER(C.c_char_p), C.c_int,
                                            abi.c_double_p] + [C.c_int] * 5
    lib.jrb_tables_write_binary.restype = C.c_int
    lib.jrb_tables_blob.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.jrb_tables_alloc_blob.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.jrb_tables_adopt_blob.argtypes = [vp]
    lib.jrb_set_kernel_variant.argtypes = [vp, C.c_int]
    lib.jrb_set_fov.argtypes = [vp, C.c_int, abi.c_double_p, abi.c_double_p]
    lib.jrb_formod_batch.argtypes = [vp, C.c_int, C.POINTER(abi.AtmView), C.POINTER(abi.ObsView)]
    lib.jrb_stage.argtypes = [vp, C.c_int, C.POINTER(abi.AtmView), C.POINTER(abi.ObsView)]
    lib.jrb_run_staged.argtypes = [vp]
    lib.jrb_fetch_staged.argtypes = [vp, C.c_int, C.POINTER(abi.ObsView)]
    lib.jrb_staged_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    lib.jrb_debug_los.argtypes = [vp, C.c_longlong, abi.c_double_p, C.c_int, abi.c_int_p, abi.c_int_p, abi.c_double_p]
    lib.jrb_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
    # lanes, page-locked caller memory, device groups, rank-style NCCL (include/jurassic_b200.h, second half)
    lib.jrb_context_device.argtypes = [vp]
    lib.jrb_tables_share.argtypes = [vp, vp]
    lib.jrb_set_los_limit_gb.argtypes = [vp, C.c_double]
    lib.jrb_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.jrb_host_unregister.argtypes = [C.c_void_p, C.c_size_t]
    lib.jrb_host_unregister_all.argtypes = []
    lib.jrb_host_is_registered.argtypes = [C.c_void_p, C.c_size_t]
    lib.jrb_staged_results_blob.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_longlong), abi.c_int_p]
    lib.jrb_group_create.argtypes = [C.POINTER(vp), C.c_int, abi.c_int_p, C.c_int]
    lib.jrb_group_destroy.argtypes = [vp]
    lib.jrb_group_destroy.restype = None
    lib.jrb_group_last_error.argtypes = [vp]
    lib.jrb_group_last_error.restype = C.c_char_p
    lib.jrb_group_size.argtypes = [vp, abi.c_int_p, abi.c_int_p]
    lib.jrb_group_context.argtypes = [vp, C.c_int, C.c_int]
    lib.jrb_group_context.restype = vp
    lib.jrb_group_set_control.argtypes = [vp, C.POINTER(abi.CtlView)]
    lib.jrb_group_set_fov.argtypes = [vp, C.c_int, abi.c_double_p, abi.c_double_p]
    lib.jrb_group_set_tables.argtypes = [vp, C.POINTER(abi.CtlView), C.POINTER(abi.TblView)]
    lib.jrb_group_formod_batch.argtypes = [vp, C.POINTER(abi.CtlView), C.c_int, C.POINTER(abi.AtmView), C.POINTER(abi.ObsView), C.c_int]
    lib.jrb_group_get_stats.argtypes = [vp, C.POINTER(abi.GroupStats)]
    lib.jrb_dist_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    lib.jrb_group_dist_init.argtypes = [vp, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.jrb_group_dist_set_tables.argtypes = [vp, C.POINTER(abi.CtlView), C.POINTER(abi.TblView), C.c_int]
    lib.jrb_group_dist_gather.argtypes = [vp, C.c_int, abi.c_int_p, C.c_int, C.POINTER(abi.ObsView)]
    lib.jrb_shared_alloc.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(vp)]
    lib.jrb_shared_free.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]
    for f in ("jrb_context_device", "jrb_tables_share", "jrb_set_los_limit_gb", "jrb_host_register", "jrb_host_unregister",
              "jrb_host_unregister_all", "jrb_host_is_registered", "jrb_staged_results_blob", "jrb_group_create", "jrb_group_size",
              "jrb_group_set_control", "jrb_group_set_fov", "jrb_group_set_tables", "jrb_group_formod_batch", "jrb_group_get_stats",
              "jrb_dist_unique_id", "jrb_group_dist_init", "jrb_group_dist_set_tables", "jrb_group_dist_gather", "jrb_shared_alloc",
              "jrb_shared_free"):
        getattr(lib, f).restype = C.c_int
    for f in ("jrb_create", "jrb_set_control", "jrb_set_tables", "jrb_tables_blob", "jrb_tables_alloc_blob",
              "jrb_tables_adopt_blob", "jrb_set_kernel_variant", "jrb_set_fov", "jrb_formod_batch", "jrb_stage", "jrb_run_staged",
              "jrb_fetch_staged", "jrb_staged_results", "jrb_debug_los", "jrb_get_stats"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["jrb_binary_tables_filename", "jrb_binary_tables_size", "jrb_tables_read_binary", "jrb_tables_write_binary",
                    "jrb_tables_read_ascii", "jrb_host_tables_view", "jrb_host_tables_free", "jrb_ingest_last_error",
                    "jrb_tables_pack_host", "jrb_tables_upload_blob", "jrb_tables_pack_info", "jrb_version", "jrb_device_count", "jrb_create", "jrb_destroy", "jrb_last_error",
                    "jrb_set_control", "jrb_set_tables", "jrb_tables_blob", "jrb_tables_alloc_blob",
                    "jrb_tables_adopt_blob", "jrb_set_kernel_variant", "jrb_set_fov", "jrb_formod_batch", "jrb_stage",
                    "jrb_run_staged", "jrb_fetch_staged", "jrb_staged_results", "jrb_debug_los", "jrb_get_stats",
                    "jrb_context_device", "jrb_tables_share", "jrb_set_los_limit_gb", "jrb_host_register", "jrb_host_unregister",
                    "jrb_host_unregister_all", "jrb_host_is_registered", "jrb_staged_results_blob", "jrb_group_create",
                    "jrb_group_destroy", "jrb_group_last_error", "jrb_group_size", "jrb_group_context", "jrb_group_set_control",
                    "jrb_group_set_fov", "jrb_group_set_tables", "jrb_group_formod_batch", "jrb_group_get_stats",
                    "jrb_dist_unique_id", "jrb_group_dist_init", "jrb_group_dist_set_tables", "jrb_group_dist_gather",
                    "jrb_shared_alloc", "jrb_shared_free"]


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p)


def tables_pack_info(tbl, ng, nd):
    """Host-only: size and properties of the packed device form of `tbl` (works without a GPU)."""
    lib = load_core()
    v = tbl.view()
    n, sh, mo, ne, gs = C.c_size_t(), C.c_int(), C.c_int(), C.c_ulonglong(), C.c_int()
    rc = lib.jrb_tables_pack_info(C.byref(v), ng, nd, C.byref(n), C.byref(sh), C.byref(mo), C.byref(ne), C.byref(gs))
    if rc != 0:
        raise JrbError(f"jrb_tables_pack_info failed ({rc}): {lib.jrb_last_error(None).decode()}")
    return {"nbytes": n.value, "all_shared": sh.value, "monotone": mo.value, "n_entries": ne.value,
            "gas_axes_same": gs.value}


def read_ascii_tables(ctl, tblbase):
    """Native ingest of the reference's ASCII .tab/.filt files -> Tables container (host only, no GPU needed)."""
    lib = load_core()
    names = (C.c_char_p * max(ctl.ng, 1))(*[e.encode() for e in ctl.emitters])
    h = C.c_void_p()
    rc = lib.jrb_tables_read_ascii(tblbase.encode(), ctl.ng, names, ctl.nd, _dp(ctl.nu), 0, 0, 0, C.byref(h))
    if rc != 0:
        raise JrbError(f"jrb_tables_read_ascii failed ({rc}): {lib.jrb_ingest_last_error().decode()}")
    v, miss = abi.TblView(), C.c_int()
    lib.jrb_host_tables_view(h, C.byref(v), C.byref(miss))
    t = Tables(v.dim_g, v.dim_d, v.dim_p, v.dim_t, v.dim_u)
    for name in ("np", "nt", "nu", "p", "t", "u", "eps", "sr", "st"):
        dst = getattr(t, name)
        src = np.ctypeslib.as_array(getattr(v, name), shape=dst.shape)
        dst[...] = src
    lib.jrb_host_tables_free(h)
    t.n_missing = miss.value
    return t


class HostTables:
    """Tables held by the core library (here: the mapped binary cache file).  Has .view() like Tables, so it can be
    handed to Context.set_tables / tables_pack_host / tables_pack_info without copying the (up to 8.8 GB) arrays."""

    def __init__(self, handle, lib):
        self.h, self.lib = handle, lib

    def view(self):
        v = abi.TblView()
        self.lib.jrb_host_tables_view(self.h, C.byref(v), None)
        return v

    @property
    def extents(self):
        v = self.view()
        return dict(NG=v.dim_g, TBLNP=v.dim_p, TBLNT=v.dim_t, TBLNU=v.dim_u, ND=v.dim_d)

    def compact(self, ng, nd):
        """copy of the populated part as a Tables container"""
        v = self.view()
        G, P, T, U, D = v.dim_g, v.dim_p, v.dim_t, v.dim_u, v.dim_d
        arr = lambda name, shape: np.ctypeslib.as_array(getattr(v, name), shape=shape)
        n_p = arr("np", (G, D))[:ng, :nd]
        mp = max(int(n_p.max(initial=0)), 1)
        n_t = arr("nt", (G, P, D))[:ng, :mp, :nd]
        mt = max(int(n_t.max(initial=0)), 1)
        n_u = arr("nu", (G, P, T, D))[:ng, :mp, :mt, :nd]
        mu = max(int(n_u.max(initial=0)), 1)
        t = Tables(max(ng, 1), nd, mp, mt, mu)
        t.np[:ng], t.nt[:ng], t.nu[:ng] = n_p, n_t, n_u
        t.p[:ng] = arr("p", (G, P, D))[:ng, :mp, :nd]
        t.t[:ng] = arr("t", (G, P, T, D))[:ng, :mp, :mt, :nd]
        t.u[:ng] = arr("u", (G, P, T, U, D))[:ng, :mp, :mt, :mu, :nd]
        t.eps[:ng] = arr("eps", (G, P, T, U, D))[:ng, :mp, :mt, :mu, :nd]
        t.sr[...] = arr("sr", (abi.TBLNS, D))[:, :nd]
        t.st[...] = arr("st", (abi.TBLNS,))
        return t

    def close(self):
        if self.h:
            self.lib.jrb_host_tables_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def binary_tables_filename(NG, ND, TBLNP=abi.TBLNP, TBLNT=abi.TBLNT, TBLNU=abi.TBLNU):
    """the reference's name of the binary cache for a build with these extents (src/jr_binary_tables_io.h:12-16)"""
    buf = C.create_string_buffer(256)
    if load_core().jrb_binary_tables_filename(buf, 256, NG, TBLNP, TBLNT, TBLNU, ND) != 0:
        raise JrbError("jrb_binary_tables_filename failed")
    return buf.value.decode()


def read_binary_tables(ctl, filename):
    """Map the reference's binary table cache -> HostTables (host only).  Raises if the header does not hold ctl's gases
    and channels at the same indices (rules of jr_binary_tables_check_header, src/jr_binary_tables_io.h:65-211)."""
    lib = load_core()
    names = (C.c_char_p * max(ctl.ng, 1))(*[e.encode() for e in ctl.emitters])
    h = C.c_void_p()
    rc = lib.jrb_tables_read_binary(os.fspath(filename).encode(), ctl.ng, names, ctl.nd, _dp(ctl.nu), C.byref(h))
    if rc != 0:
        raise JrbError(f"jrb_tables_read_binary failed ({rc}): {lib.jrb_ingest_last_error().decode()}")
    return HostTables(h, lib)


def write_binary_tables(filename, tbl, ctl, NG, ND, TBLNP=abi.TBLNP, TBLNT=abi.TBLNT, TBLNU=abi.TBLNU):
    """Write `tbl` as the binary cache file of a reference build with the given compile-time extents (sparse file)."""
    lib = load_core()
    names = (C.c_char_p * max(ctl.ng, 1))(*[e.encode() for e in ctl.emitters])
    v = tbl.view()
    rc = lib.jrb_tables_write_binary(os.fspath(filename).encode(), C.byref(v), ctl.ng, names, ctl.nd, _dp(ctl.nu), NG, TBLNP, TBLNT,
                                     TBLNU, ND)
    if rc != 0:
        raise JrbError(f"jrb_tables_write_binary failed ({rc}): {lib.jrb_ingest_last_error().decode()}")


def tables_pack_host(tbl, ng, nd):
    """Host-only: the packed, position-independent table blob as a numpy uint8 array (works without a GPU)."""
    lib = load_core()
    v = tbl.view()
    n = C.c_size_t()
    rc = lib.jrb_tables_pack_host(C.byref(v), ng, nd, None, 0, C.byref(n))
    if rc != 0:
        raise JrbError(f"jrb_tables_pack_host failed ({rc}): {lib.jrb_last_error(None).decode()}")
    out = np.empty(n.value, dtype=np.uint8)
    rc = lib.jrb_tables_pack_host(C.byref(v), ng, nd, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n))
    if rc != 0:
        raise JrbError(f"jrb_tables_pack_host failed ({rc}): {lib.jrb_last_error(None).decode()}")
    return out


# ---------------------------------------------------------------------------------------------------------------
# dimension-agnostic containers (numpy, C-contiguous)
# ---------------------------------------------------------------------------------------------------------------
class Control:
    """The ctl_t fields the path reads (src/jurassic.h:229-347), defaults as read_ctl (src/jurassic.c:928-1021)."""

    def __init__(self, emitters, nu, window=None, nw=1, ctm_co2=1, ctm_h2o=1, ctm_n2=1, ctm_o2=1, refrac=1,
                 rayds=10.0, raydz=0.5, hydz=-999.0, write_bbt=0, tblbase="-"):
        self.emitters = list(emitters)
        self.ng = len(self.emitters)
        self.nu = np.ascontiguousarray(nu, dtype=np.float64)
        self.nd = int(self.nu.size)
        self.nw = nw
        self.window = np.zeros(self.nd, dtype=np.int32) if window is None else np.ascontiguousarray(window, np.int32)
        self.ctm_co2, self.ctm_h2o, self.ctm_n2, self.ctm_o2 = ctm_co2, ctm_h2o, ctm_n2, ctm_o2
        self.refrac, self.rayds, self.raydz, self.hydz, self.write_bbt = refrac, rayds, raydz, hydz, write_bbt
        self.formod, self.ip = 2, 1
        self.cz, self.cx = 0.0, 0.0  # influence radii of the 3-D interpolation (ip = 3), read_ctl defaults
        self.tblbase = tblbase
        self.auto_ctm()

    def auto_ctm(self):
        """read_ctl's automatic switch-off of continua without a channel in range (src/jurassic.c:954-968)."""
        nu = self.nu
        if not np.any(nu < 4000): self.ctm_co2 = 0
        if not np.any(nu < 20000): self.ctm_h2o = 0
        if not np.any((nu >= 2120) & (nu <= 2605)): self.ctm_n2 = 0
        if not np.any((nu >= 1360) & (nu <= 1805)): self.ctm_o2 = 0

    def find_emitter(self, name):
        for i, e in enumerate(self.emitters):
            if e.lower() == name.lower():
                return i
        return -1

    def view(self):
        v = abi.CtlView()
        v.ng, v.nd, v.nw = self.ng, self.nd, self.nw
        v.nu = _dp(self.nu)
        v.window = self.window.ctypes.data_as(abi.c_int_p)
        v.ctm_co2, v.ctm_h2o, v.ctm_n2, v.ctm_o2 = self.ctm_co2, self.ctm_h2o, self.ctm_n2, self.ctm_o2
        v.ig_h2o = self.find_emitter("H2O") if self.ctm_h2o else -999
        v.ig_co2 = self.find_emitter("CO2") if self.ctm_co2 else -999
        v.refrac, v.rayds, v.raydz, v.hydz = self.refrac, self.rayds, self.raydz, self.hydz
        v.write_bbt, v.formod, v.ip = self.write_bbt, self.formod, self.ip
        v.cz, v.cx = self.cz, self.cx
        return v

    @property
    def ctm_mask(self):
        v = self.view()
        return ((v.ctm_co2 == 1 and v.ig_co2 >= 0) * 8 + (v.ctm_h2o == 1 and v.ig_h2o >= 0) * 4 +
                (v.ctm_n2 == 1) * 2 + (v.ctm_o2 == 1))


class Tables:
    """Emissivity + source tables in tbl_t's row-major [g][p][T][u][d] order with freely chosen extents."""

    def __init__(self, ng, nd, dim_p, dim_t, dim_u):
        self.dims = (ng, dim_p, dim_t, dim_u, nd)
        self.np = np.zeros((ng, nd), np.int32)
        self.nt = np.zeros((ng, dim_p, nd), np.int32)
        self.nu = np.zeros((ng, dim_p, dim_t, nd), np.int32)
        self.p = np.zeros((ng, dim_p, nd), np.float64)
        self.t = np.zeros((ng, dim_p, dim_t, nd), np.float64)
        self.u = np.zeros((ng, dim_p, dim_t, dim_u, nd), np.float32)
        self.eps = np.zeros((ng, dim_p, dim_t, dim_u, nd), np.float32)
        self.sr = np.zeros((abi.TBLNS, nd), np.float64)
        self.st = 100.0 + 0.25 * np.arange(abi.TBLNS, dtype=np.float64)

    def view(self):
        g, p, t, u, d = self.dims
        v = abi.TblView()
        v.dim_g, v.dim_p, v.dim_t, v.dim_u, v.dim_d, v.dim_s = g, p, t, u, d, abi.TBLNS
        i32 = C.POINTER(C.c_int32)
        v.np, v.nt, v.nu = (self.np.ctypes.data_as(i32), self.nt.ctypes.data_as(i32), self.nu.ctypes.data_as(i32))
        v.p, v.t = _dp(self.p), _dp(self.t)
        v.u = self.u.ctypes.data_as(C.POINTER(C.c_float))
        v.eps = self.eps.ctypes.data_as(C.POINTER(C.c_float))
        v.sr, v.st = _dp(self.sr), _dp(self.st)
        return v


class Package:
    """One (atm_t, obs_t) pair of the reference, as compact arrays."""

    def __init__(self, ng, nw, nd, n_atm, n_rays):
        self.ng, self.nw, self.nd = ng, nw, nd
        self.atm_time = np.zeros(n_atm); self.z = np.zeros(n_atm); self.lon = np.zeros(n_atm)
        self.lat = np.zeros(n_atm); self.p = np.zeros(n_atm); self.t = np.zeros(n_atm)
        self.q = np.zeros((max(ng, 1), n_atm)); self.k = np.zeros((max(nw, 1), n_atm))
        self.time = np.zeros(n_rays); self.obsz = np.zeros(n_rays); self.obslon = np.zeros(n_rays)
        self.obslat = np.zeros(n_rays); self.vpz = np.zeros(n_rays); self.vplon = np.zeros(n_rays)
        self.vplat = np.zeros(n_rays)
        self.tpz = np.zeros(n_rays); self.tplon = np.zeros(n_rays); self.tplat = np.zeros(n_rays)
        self.rad = np.zeros((n_rays, nd)); self.tau = np.zeros((n_rays, nd))

    @property
    def n_atm(self):
        return self.z.size

    @property
    def n_rays(self):
        return self.obsz.size

    def atm_view(self):
        v = abi.AtmView()
        v.np = self.n_atm
        v.time, v.z, v.lon, v.lat, v.p, v.t = (_dp(self.atm_time), _dp(self.z), _dp(self.lon), _dp(self.lat),
                                               _dp(self.p), _dp(self.t))
        v.q, v.q_stride = _dp(self.q), self.q.shape[1]
        v.k, v.k_stride = _dp(self.k), self.k.shape[1]
        return v

    def obs_view(self):
        v = abi.ObsView()
        v.nr = self.n_rays
        v.time, v.obsz, v.obslon, v.obslat = _dp(self.time), _dp(self.obsz), _dp(self.obslon), _dp(self.obslat)
        v.vpz, v.vplon, v.vplat = _dp(self.vpz), _dp(self.vplon), _dp(self.vplat)
        v.tpz, v.tplon, v.tplat = _dp(self.tpz), _dp(self.tplon), _dp(self.tplat)
        v.rad, v.tau = _dp(self.rad), _dp(self.tau)
        v.row_stride, v.nd_reset = self.rad.shape[1], self.rad.shape[1]
        return v


# ---------------------------------------------------------------------------------------------------------------
class Context:
    """One GPU context of the core library."""

    def __init__(self, device=0, handle=None):
        self.lib = load_core()
        self._keep = []
        self.owned = handle is None
        if handle is not None:  # a context owned by somebody else (a lane of a group, the drop-in layer's context)
            self.h = C.c_void_p(handle)
            return
        self.h = C.c_void_p()
        rc = self.lib.jrb_create(C.byref(self.h), device)
        if rc != 0:
            raise JrbError(f"jrb_create failed ({rc}): {self.lib.jrb_last_error(None).decode()}")

    def _check(self, rc, what):
        if rc != 0:
            raise JrbError(f"{what} failed ({rc}): {self.lib.jrb_last_error(self.h).decode()}")

    def close(self):
        if self.h and self.owned:
            self.lib.jrb_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stage_views(self, n, av, ov):
        """stage packages described by ready-made view arrays (e.g. views onto the reference's structs)"""
        self._check(self.lib.jrb_stage(self.h, n, av, ov), "jrb_stage")

    def set_control(self, ctl):
        v = ctl.view()
        self._check(self.lib.jrb_set_control(self.h, C.byref(v)), "jrb_set_control")

    def set_tables(self, tbl):
        v = tbl.view()
        self._check(self.lib.jrb_set_tables(self.h, C.byref(v)), "jrb_set_tables")

    def set_kernel_variant(self, variant):
        self._check(self.lib.jrb_set_kernel_variant(self.h, variant), "jrb_set_kernel_variant")

    def set_fov(self, dz=None, w=None):
        """field-of-view epilogue (formod_fov, src/jurassic.c:214-258): shape offsets dz [km] and weights w; None = off"""
        if dz is None or len(dz) == 0:
            self._check(self.lib.jrb_set_fov(self.h, 0, None, None), "jrb_set_fov")
            return
        dz, w = np.ascontiguousarray(dz, dtype=np.float64), np.ascontiguousarray(w, dtype=np.float64)
        assert dz.shape == w.shape and dz.ndim == 1
        self._check(self.lib.jrb_set_fov(self.h, len(dz), _dp(dz), _dp(w)), "jrb_set_fov")

    def tables_blob(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.jrb_tables_blob(self.h, C.byref(p), C.byref(n)), "jrb_tables_blob")
        return p.value, n.value

    def tables_alloc_blob(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.jrb_tables_alloc_blob(self.h, nbytes, C.byref(p)), "jrb_tables_alloc_blob")
        return p.value

    def tables_upload_blob(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._check(self.lib.jrb_tables_upload_blob(self.h, blob.ctypes.data_as(C.c_void_p), blob.size), "jrb_tables_upload_blob")

    def tables_adopt_blob(self):
        self._check(self.lib.jrb_tables_adopt_blob(self.h), "jrb_tables_adopt_blob")

    def _views(self, packages):
        n = len(packages)
        av = (abi.AtmView * n)(*[p.atm_view() for p in packages])
        ov = (abi.ObsView * n)(*[p.obs_view() for p in packages])
        return n, av, ov

    def formod_batch(self, packages):
        n, av, ov = self._views(packages)
        self._check(self.lib.jrb_formod_batch(self.h, n, av, ov), "jrb_formod_batch")

    def stage(self, packages):
        n, av, ov = self._views(packages)
        self._check(self.lib.jrb_stage(self.h, n, av, ov), "jrb_stage")

    def run_staged(self):
        self._check(self.lib.jrb_run_staged(self.h), "jrb_run_staged")

    def fetch_staged(self, packages):
        n, av, ov = self._views(packages)
        self._check(self.lib.jrb_fetch_staged(self.h, n, ov), "jrb_fetch_staged")

    def staged_results(self):
        r, t, n, nd = C.c_void_p(), C.c_void_p(), C.c_longlong(), C.c_int()
        self._check(self.lib.jrb_staged_results(self.h, C.byref(r), C.byref(t), C.byref(n), C.byref(nd)),
                    "jrb_staged_results")
        return r.value, t.value, n.value, nd.value

    def debug_los(self, ray):
        npo, rec, ts = C.c_int(), C.c_int(), C.c_double()
        buf = np.zeros(abi.NLOS * 512)
        self._check(self.lib.jrb_debug_los(self.h, ray, _dp(buf), buf.size, C.byref(npo), C.byref(rec), C.byref(ts)),
                    "jrb_debug_los")
        return buf[: npo.value * rec.value].reshape(npo.value, rec.value).copy(), ts.value

    def stats(self):
        s = abi.Stats()
        self._check(self.lib.jrb_get_stats(self.h, C.byref(s)), "jrb_get_stats")
        return {f[0]: getattr(s, f[0]) for f in abi.Stats._fields_}


def host_register(arr):
    """page-lock the memory of a numpy array (jrb_host_register); batches whose arrays are all registered run in direct mode"""
    rc = load_core().jrb_host_register(arr.ctypes.data_as(C.c_void_p), arr.nbytes)
    if rc != 0:
        raise JrbError(f"jrb_host_register failed ({rc}): {load_core().jrb_last_error(None).decode()}")


def host_unregister_all():
    load_core().jrb_host_unregister_all()


def register_package(pkg):
    """page-lock every array of a Package"""
    for name in ("atm_time", "z", "lon", "lat", "p", "t", "q", "k", "time", "obsz", "obslon", "obslat", "vpz", "vplon", "vplat",
                 "tpz", "tplon", "tplat", "rad", "tau"):
        host_register(getattr(pkg, name))


class Group:
    """devices x lanes behind one handle (jrb_group_*)"""

    def __init__(self, ndev=1, devices=None, nlanes=1):
        self.lib = load_core()
        self.h = C.c_void_p()
        dev = None if devices is None else (C.c_int * len(devices))(*devices)
        rc = self.lib.jrb_group_create(C.byref(self.h), ndev, dev, nlanes)
        if rc != 0:
            raise JrbError(f"jrb_group_create failed ({rc}): {self.lib.jrb_last_error(None).decode()}")

    def _check(self, rc, what):
        if rc != 0:
            raise JrbError(f"{what} failed ({rc}): {self.lib.jrb_group_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.jrb_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        a, b = C.c_int(), C.c_int()
        self.lib.jrb_group_size(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_tables(self, ctl, tbl):
        cv, tv = ctl.view(), tbl.view()
        self._check(self.lib.jrb_group_set_tables(self.h, C.byref(cv), C.byref(tv)), "jrb_group_set_tables")

    def set_fov(self, dz, w):
        dz, w = np.ascontiguousarray(dz, dtype=np.float64), np.ascontiguousarray(w, dtype=np.float64)
        self._check(self.lib.jrb_group_set_fov(self.h, len(dz), _dp(dz), _dp(w)), "jrb_group_set_fov")

    def formod_batch(self, packages, ctl=None, use_fov=0):
        n = len(packages)
        av = (abi.AtmView * n)(*[p.atm_view() for p in packages])
        ov = (abi.ObsView * n)(*[p.obs_view() for p in packages])
        cv = ctl.view() if ctl is not None else None
        self._check(self.lib.jrb_group_formod_batch(self.h, C.byref(cv) if cv is not None else None, n, av, ov, use_fov),
                    "jrb_group_formod_batch")

    def context_stats(self, dev=0, lane=0):
        s = abi.Stats()
        c = self.lib.jrb_group_context(self.h, dev, lane)
        self.lib.jrb_get_stats(c, C.byref(s))
        return {f[0]: getattr(s, f[0]) for f in abi.Stats._fields_}

    def stats(self):
        s = abi.GroupStats()
        self._check(self.lib.jrb_group_get_stats(self.h, C.byref(s)), "jrb_group_get_stats")
        return {f[0]: getattr(s, f[0]) for f in abi.GroupStats._fields_}

    # rank style
    def dist_init(self, rank, nranks, uid):
        buf = (C.c_char * 128).from_buffer_copy(bytes(uid)[:128].ljust(128, b"\0"))
        self._check(self.lib.jrb_group_dist_init(self.h, rank, nranks, buf, 128), "jrb_group_dist_init")

    def dist_set_tables(self, ctl, tbl, root=0):
        cv = ctl.view()
        tv = tbl.view() if tbl is not None else None
        self._check(self.lib.jrb_group_dist_set_tables(self.h, C.byref(cv), C.byref(tv) if tv is not None else None, root),
                    "jrb_group_dist_set_tables")

    def dist_gather(self, counts, packages_all=None, root=0):
        cnt = (C.c_int * len(counts))(*counts)
        if packages_all is None:
            self._check(self.lib.jrb_group_dist_gather(self.h, root, cnt, sum(counts), None), "jrb_group_dist_gather")
            return
        n = len(packages_all)
        ov = (abi.ObsView * n)(*[p.obs_view() for p in packages_all])
        self._check(self.lib.jrb_group_dist_gather(self.h, root, cnt, n, ov), "jrb_group_dist_gather")


def dist_unique_id():
    buf = (C.c_char * 128)()
    if load_core().jrb_dist_unique_id(buf, 128) != 0:
        raise JrbError("jrb_dist_unique_id failed (NCCL missing?)")
    return bytes(buf)
