"""ctypes mirrors of the reference's interface structs (src/jurassic.h:215-425) and of the view structs of
include/jurassic_b200.h.  The reference fixes its array extents at compile time (ND, NG overridable with -D), so the
struct classes are generated per dimension set."""
import ctypes as C
from functools import lru_cache

NP, NR, NW, LEN, NLOS = 9600, 1088, 1, 5000, 400
TBLNP, TBLNT, TBLNU, TBLNS = 40, 30, 304, 1201

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


@lru_cache(maxsize=None)
def structs(ND, NG):
    """Return (ctl_t, atm_t, obs_t, tbl_t) ctypes classes for the given compile-time dimensions."""

    class atm_t(C.Structure):
        _fields_ = [("time", C.c_double * NP), ("z", C.c_double * NP), ("lon", C.c_double * NP),
                    ("lat", C.c_double * NP), ("p", C.c_double * NP), ("t", C.c_double * NP),
                    ("q", (C.c_double * NP) * NG), ("k", (C.c_double * NP) * NW),
                    ("np", C.c_int), ("init", C.c_int)]

    class ctl_t(C.Structure):
        _fields_ = [("ng", C.c_int), ("emitter", (C.c_char * LEN) * NG), ("nd", C.c_int), ("nw", C.c_int),
                    ("nu", C.c_double * ND), ("window", C.c_int * ND), ("tblbase", C.c_char * LEN),
                    ("hydz", C.c_double), ("ctm_co2", C.c_int), ("ctm_h2o", C.c_int), ("ctm_n2", C.c_int),
                    ("ctm_o2", C.c_int), ("ip", C.c_int), ("cz", C.c_double), ("cx", C.c_double),
                    ("refrac", C.c_int), ("rayds", C.c_double), ("raydz", C.c_double), ("fov", C.c_char * LEN),
                    ("retp_zmin", C.c_double), ("retp_zmax", C.c_double), ("rett_zmin", C.c_double),
                    ("rett_zmax", C.c_double), ("retq_zmin", C.c_double * NG), ("retq_zmax", C.c_double * NG),
                    ("retk_zmin", C.c_double * NW), ("retk_zmax", C.c_double * NW), ("write_bbt", C.c_int),
                    ("write_matrix", C.c_int), ("formod", C.c_int), ("rfmbin", C.c_char * LEN),
                    ("rfmhit", C.c_char * LEN), ("rfmxsc", (C.c_char * LEN) * NG), ("useGPU", C.c_int),
                    ("checkmode", C.c_int), ("MPIglobrank", C.c_int), ("MPIlocalrank", C.c_int),
                    ("read_binary", C.c_int), ("write_binary", C.c_int), ("gpu_nbytes_shared_memory", C.c_int)]

    class obs_t(C.Structure):
        _fields_ = [("time", C.c_double * NR), ("obsz", C.c_double * NR), ("obslon", C.c_double * NR),
                    ("obslat", C.c_double * NR), ("vpz", C.c_double * NR), ("vplon", C.c_double * NR),
                    ("vplat", C.c_double * NR), ("tpz", C.c_double * NR), ("tplon", C.c_double * NR),
                    ("tplat", C.c_double * NR), ("tau", (C.c_double * ND) * NR), ("rad", (C.c_double * ND) * NR),
                    ("nr", C.c_int)]

    class tbl_t(C.Structure):
        _fields_ = [("np", (C.c_int32 * ND) * NG), ("nt", ((C.c_int32 * ND) * TBLNP) * NG),
                    ("nu", (((C.c_int32 * ND) * TBLNT) * TBLNP) * NG), ("p", ((C.c_double * ND) * TBLNP) * NG),
                    ("t", (((C.c_double * ND) * TBLNT) * TBLNP) * NG),
                    ("u", ((((C.c_float * ND) * TBLNU) * TBLNT) * TBLNP) * NG),
                    ("eps", ((((C.c_float * ND) * TBLNU) * TBLNT) * TBLNP) * NG),
                    ("sr", (C.c_double * ND) * TBLNS), ("st", C.c_double * TBLNS)]

    return ctl_t, atm_t, obs_t, tbl_t


# ---- views of include/jurassic_b200.h -----------------------------------------------------------------------
class CtlView(C.Structure):
    _fields_ = [("ng", C.c_int), ("nd", C.c_int), ("nw", C.c_int), ("nu", c_double_p), ("window", c_int_p),
                ("ctm_co2", C.c_int), ("ctm_h2o", C.c_int), ("ctm_n2", C.c_int), ("ctm_o2", C.c_int),
                ("ig_co2", C.c_int), ("ig_h2o", C.c_int), ("refrac", C.c_int), ("rayds", C.c_double),
                ("raydz", C.c_double), ("hydz", C.c_double), ("write_bbt", C.c_int), ("formod", C.c_int),
                ("ip", C.c_int), ("cz", C.c_double), ("cx", C.c_double)]


class AtmView(C.Structure):
    _fields_ = [("np", C.c_int), ("time", c_double_p), ("z", c_double_p), ("lon", c_double_p), ("lat", c_double_p),
                ("p", c_double_p), ("t", c_double_p), ("q", c_double_p), ("q_stride", C.c_long),
                ("k", c_double_p), ("k_stride", C.c_long), ("q_rows", C.POINTER(c_double_p)),
                ("k_rows", C.POINTER(c_double_p))]


class ObsView(C.Structure):
    _fields_ = [("nr", C.c_int), ("time", c_double_p), ("obsz", c_double_p), ("obslon", c_double_p),
                ("obslat", c_double_p), ("vpz", c_double_p), ("vplon", c_double_p), ("vplat", c_double_p),
                ("tpz", c_double_p), ("tplon", c_double_p), ("tplat", c_double_p), ("rad", c_double_p),
                ("tau", c_double_p), ("row_stride", C.c_long), ("nd_reset", C.c_int)]


class TblView(C.Structure):
    _fields_ = [("dim_g", C.c_int), ("dim_p", C.c_int), ("dim_t", C.c_int), ("dim_u", C.c_int), ("dim_d", C.c_int),
                ("dim_s", C.c_int), ("np", C.POINTER(C.c_int32)), ("nt", C.POINTER(C.c_int32)),
                ("nu", C.POINTER(C.c_int32)), ("p", c_double_p), ("t", c_double_p), ("u", C.POINTER(C.c_float)),
                ("eps", C.POINTER(C.c_float)), ("sr", c_double_p), ("st", c_double_p)]


class Stats(C.Structure):
    _fields_ = [("n_packages", C.c_longlong), ("n_rays", C.c_longlong), ("n_ray_channels", C.c_longlong),
                ("n_los_points", C.c_longlong), ("n_kernel_launches", C.c_longlong), ("ms_raytrace", C.c_float),
                ("ms_ega", C.c_float), ("ms_total_device", C.c_float), ("h2d_bytes", C.c_longlong),
                ("d2h_bytes", C.c_longlong), ("ega_kernel_variant", C.c_int), ("ega_ngb", C.c_int),
                ("ega_ctm_mask", C.c_int), ("table_blob_bytes", C.c_longlong), ("host_ms_pack", C.c_float),
                ("host_ms_h2d", C.c_float), ("host_ms_d2h", C.c_float), ("host_ms_scatter", C.c_float),
                ("n_chunks", C.c_int), ("pipelined", C.c_int), ("ega_phase_lock", C.c_int), ("ega_channels_per_warp", C.c_int),
                ("io_direct", C.c_int), ("host_ms_stage", C.c_float), ("cum_runs", C.c_longlong), ("cum_launches", C.c_longlong),
                ("cum_ega_launches", C.c_longlong), ("cum_ms_ega", C.c_double), ("cum_ms_raytrace", C.c_double),
                ("cum_ms_device", C.c_double), ("ega_per_channel_axes", C.c_int), ("ega_tiled", C.c_int), ("ega_gas_blocks", C.c_int)]


class GroupStats(C.Structure):
    _fields_ = [("ndev", C.c_int), ("nlanes", C.c_int), ("n_slices", C.c_int), ("nccl_nranks", C.c_int), ("dist_rank", C.c_int),
                ("dist_nranks", C.c_int), ("table_bytes", C.c_longlong), ("gather_bytes", C.c_longlong), ("ms_tables", C.c_float),
                ("ms_last_call", C.c_float), ("ms_gather", C.c_float), ("ms_gather_scatter", C.c_float)]


def atm_view_of(a, ND, NG):
    """AtmView onto a ctypes atm_t (no copy)"""
    cls = type(a)
    base = C.addressof(a)
    v = AtmView()
    v.np = a.np
    for name in ("time", "z", "lon", "lat", "p", "t"):
        setattr(v, name, C.cast(base + getattr(cls, name).offset, c_double_p))
    v.q = C.cast(base + cls.q.offset, c_double_p); v.q_stride = NP
    v.k = C.cast(base + cls.k.offset, c_double_p); v.k_stride = NP
    return v


def obs_view_of(o, ND):
    """ObsView onto a ctypes obs_t (no copy)"""
    cls = type(o)
    base = C.addressof(o)
    v = ObsView()
    v.nr = o.nr
    for name in ("time", "obsz", "obslon", "obslat", "vpz", "vplon", "vplat", "tpz", "tplon", "tplat", "rad", "tau"):
        setattr(v, name, C.cast(base + getattr(cls, name).offset, c_double_p))
    v.row_stride, v.nd_reset = ND, ND
    return v
