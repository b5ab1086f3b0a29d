"""Deterministic synthetic workloads for the EGA forward model.

The reference checkout ships no emissivity tables (.MISSING_LARGE_BLOBS), so every parity and throughput case uses
analytic tables generated here; geometry follows the reference's own generators (src/limb.c:50-59,
src/nadir.c:51-58), atmospheres follow src/climatology.c:66-78 (per-profile random p/T offsets) on top of the
91-level mid-latitude profile committed as tests/golden/limb/atm.tab.
"""
import os

import numpy as np

from .core import Control, Package, Tables

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(_ROOT, "tests", "golden")

RE = 6367.421  # src/jurassic.h:126
C1, C2 = 1.19104259e-8, 1.43877506  # src/jurassic.h:111-114

LIMB_GASES = ["CO2", "H2O", "O3", "F11", "CCl4"]
NADIR_GASES = ["CO2", "H2O", "O3", "N2O", "CH4", "CO", "HNO3", "SO2"]
KAPPA = {"CO2": 3e-22, "H2O": 2e-23, "O3": 5e-21, "F11": 4e-19, "CCl4": 6e-19, "N2O": 2e-21, "CH4": 1e-21,
         "CO": 3e-21, "HNO3": 5e-20, "SO2": 1e-20}


# ---------------------------------------------------------------------------------------------------------------
def read_tab(path):
    rows = [l.split() for l in open(path) if l.strip() and not l.startswith("#")]
    return np.array(rows, dtype=np.float64)


def base_profile():
    """z, p, T and vmr of CO2, H2O, O3, F11, CCl4 at 0..90 km from the committed limb example atmosphere."""
    a = read_tab(os.path.join(GOLDEN, "limb", "atm.tab"))
    prof = {"z": a[:, 1], "p": a[:, 4], "t": a[:, 5]}
    for i, g in enumerate(LIMB_GASES):
        prof[g] = a[:, 6 + i]
    z = prof["z"]
    # smooth analytic profiles for the additional nadir emitters (synthetic, order-of-magnitude realistic)
    prof["N2O"] = 3.2e-7 * np.where(z < 15, 1.0, np.exp(-(z - 15) / 6.0))
    prof["CH4"] = 1.75e-6 * np.where(z < 12, 1.0, np.exp(-(z - 12) / 14.0))
    prof["CO"] = 1.2e-7 * np.exp(-z / 9.0) + 2e-8
    prof["HNO3"] = 6e-9 * np.exp(-((z - 23.0) / 6.0) ** 2) + 5e-11
    prof["SO2"] = 1e-10 * np.exp(-z / 4.0) + 1e-11
    return prof


def make_atmosphere(pkg, gases, n_profiles, rng, perturb=True):
    """Fill pkg's atmosphere with n_profiles time-keyed 91-level profiles (time = profile index)."""
    prof = base_profile()
    nz = prof["z"].size
    for j in range(n_profiles):
        s = slice(j * nz, (j + 1) * nz)
        dp = rng.uniform(-0.05, 0.05) if perturb else 0.0
        dt = rng.uniform(-30.0, 30.0) if perturb else 0.0
        pkg.atm_time[s] = float(j)
        pkg.z[s] = prof["z"]
        pkg.lon[s] = 0.0
        pkg.lat[s] = 0.0
        pkg.p[s] = prof["p"] * (1.0 + dp)
        pkg.t[s] = prof["t"] + dt
        for ig, g in enumerate(gases):
            pkg.q[ig, s] = prof[g] if g in prof else 1e-9  # unknown emitter: constant trace amount
        pkg.k[:, s] = 0.0


def limb_package(ctl, n_profiles=17, rays_per_profile=64, z0=3.0, dz=1.0, obsz=780.0, seed=20240517, perturb=True):
    """Config-D style package: n_profiles x rays_per_profile limb rays (tangent heights z0 + k dz)."""
    rng = np.random.default_rng(seed)
    nz = base_profile()["z"].size
    pkg = Package(ctl.ng, ctl.nw, ctl.nd, n_profiles * nz, n_profiles * rays_per_profile)
    make_atmosphere(pkg, ctl.emitters, n_profiles, rng, perturb)
    for j in range(n_profiles):
        for k in range(rays_per_profile):
            r = j * rays_per_profile + k
            vpz = z0 + k * dz
            pkg.time[r] = float(j)
            pkg.obsz[r] = obsz
            pkg.vpz[r] = vpz
            pkg.vplat[r] = 180.0 / np.pi * np.arccos((RE + vpz) / (RE + obsz))
    return pkg


def track_package(ctl, n_profiles=9, rays=24, lat0=-10.0, dlat=3.0, z0=6.0, dz=1.5, obsz=780.0, seed=20240519, obslat=-24.0):
    """Atmosphere for the 2-D / 3-D interpolation (ctl.ip = 2, 3; src/jurassic.c:704-804): n_profiles columns of one time
    stamp along a meridional track (latitude lat0 + j*dlat, longitude drifting a little), each with its own pressure and
    temperature perturbation and vmr scaling, and limb rays whose paths cross several columns (the default track covers the
    rays from where they enter the atmosphere, latitude -6, to where they leave it, latitude +13)."""
    rng = np.random.default_rng(seed)
    prof = base_profile()
    nz = prof["z"].size
    pkg = Package(ctl.ng, ctl.nw, ctl.nd, n_profiles * nz, rays)
    for j in range(n_profiles):
        s = slice(j * nz, (j + 1) * nz)
        pkg.atm_time[s] = 0.0
        pkg.z[s] = prof["z"]
        pkg.lon[s] = 0.3 * j
        pkg.lat[s] = lat0 + dlat * j
        pkg.p[s] = prof["p"] * (1.0 + rng.uniform(-0.05, 0.05))
        pkg.t[s] = prof["t"] + rng.uniform(-20.0, 20.0) + 3.0 * np.sin(prof["z"] / 7.0 + j)
        scale = 1.0 + rng.uniform(-0.3, 0.3)
        for ig, g in enumerate(ctl.emitters):
            pkg.q[ig, s] = (prof[g] if g in prof else 1e-9) * scale
        pkg.k[:, s] = 1e-4 * np.exp(-prof["z"] / 8.0) * (1.0 + 0.1 * j)
    for r in range(rays):
        vpz = z0 + dz * r
        # observer south of the track looking north: tangent point near the middle of the track
        a = np.arccos((RE + vpz) / (RE + obsz)) * 180.0 / np.pi
        pkg.time[r] = 0.0
        pkg.obsz[r] = obsz
        pkg.obslat[r] = obslat
        pkg.obslon[r] = 0.3 * (n_profiles - 1) / 2
        pkg.vpz[r] = vpz
        pkg.vplat[r] = obslat + a
        pkg.vplon[r] = 0.3 * (n_profiles - 1) / 2 + 0.05 * r / max(rays - 1, 1)
    return pkg


def nadir_package(ctl, n_profiles=16, rays_per_profile=68, obsz=700.0, lat0=-6.03, dlat=0.18, seed=20240518,
                  perturb=True):
    """Config-E style package: nadir footprints looking at the ground (surface term active)."""
    rng = np.random.default_rng(seed)
    nz = base_profile()["z"].size
    pkg = Package(ctl.ng, ctl.nw, ctl.nd, n_profiles * nz, n_profiles * rays_per_profile)
    make_atmosphere(pkg, ctl.emitters, n_profiles, rng, perturb)
    for j in range(n_profiles):
        for m in range(rays_per_profile):
            r = j * rays_per_profile + m
            pkg.time[r] = float(j)
            pkg.obsz[r] = obsz
            pkg.vpz[r] = 0.0
            pkg.vplat[r] = lat0 + dlat * m
    return pkg


def example_package(case, ctl):
    """The reference's example/limb or example/nadir inputs (committed fixtures)."""
    a = read_tab(os.path.join(GOLDEN, case, "atm.tab"))
    o = read_tab(os.path.join(GOLDEN, case, "obs.tab"))
    pkg = Package(ctl.ng, ctl.nw, ctl.nd, a.shape[0], o.shape[0])
    pkg.atm_time[:], pkg.z[:], pkg.lon[:], pkg.lat[:], pkg.p[:], pkg.t[:] = a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4], a[:, 5]
    for ig in range(ctl.ng):
        pkg.q[ig, :] = a[:, 6 + ig]
    for iw in range(ctl.nw):
        pkg.k[iw, :] = a[:, 6 + ctl.ng + iw]
    pkg.time[:], pkg.obsz[:], pkg.obslon[:], pkg.obslat[:] = o[:, 0], o[:, 1], o[:, 2], o[:, 3]
    pkg.vpz[:], pkg.vplon[:], pkg.vplat[:] = o[:, 4], o[:, 5], o[:, 6]
    return pkg


# ---------------------------------------------------------------------------------------------------------------
def control_limb_example():
    return Control(LIMB_GASES, [792.0, 832.0])


def control_nadir_example():
    return Control(["CO2"], [667.7820, 668.5410, 669.8110], write_bbt=1)


def control_config_d(nd=32):
    """Synthetic limb sounder: 5 gases, channels 785+i cm^-1 (CO2 + H2O continua on, N2/O2 off -> mask 1100)."""
    return Control(LIMB_GASES, 785.0 + np.arange(nd))


def control_config_e(nd=128):
    """Synthetic AIRS-like nadir: 8 gases, three channel groups so that all four continua are active (mask 1111)."""
    i = np.arange(nd)
    a, b = nd // 2, nd // 2 + nd // 4
    nu = np.where(i < a, 650.0 + i, np.where(i < b, 1370.0 + 7.0 * (i - a), 2150.0 + 7.0 * (i - b)))
    return Control(NADIR_GASES, nu)


# ---------------------------------------------------------------------------------------------------------------
def planck(t, nu):
    return C1 * nu ** 3 / np.expm1(C2 * nu / t)  # planck(), src/jurassic.c:860


def boxcar_filter(nu0):
    """nu +- 0.5 cm^-1 boxcar, step 0.01, two zero guard points on each side."""
    nu = np.round(nu0 - 0.52 + 0.01 * np.arange(105), 4)
    f = np.ones(105)
    f[:2] = 0.0
    f[-2:] = 0.0
    return nu, f


def source_table(nus):
    """sr[TBLNS][nd]: filter-averaged Planck radiance (init_tbl, src/jurassic.c:645-667)."""
    st = 100.0 + 0.25 * np.arange(1201)
    sr = np.zeros((1201, len(nus)))
    for d, nu0 in enumerate(nus):
        nu, f = boxcar_filter(nu0)
        w = f / f.sum()
        sr[:, d] = (planck(st[:, None], nu[None, :]) * w[None, :]).sum(axis=1)
    return sr


def kappa0(gas, ig, d):
    return KAPPA.get(gas, 1e-20) * (1.0 + 0.5 * np.sin(0.37 * d + ig))


P_AXIS = 1e-3 * 10.0 ** (6.2 * np.arange(36) / 35.0)
T_AXIS = 180.0 + 12.0 * np.arange(12)
_U_GRID = 1e12 * 10.0 ** (0.05 * np.arange(460))


def make_tables(ctl, skip_pairs=(), dim_u=None, axis_jitter=False, gas_axis_shift=False):
    """Analytic emissivity tables eps(p,T,u) = 1 - exp(-k u)/2 - exp(-0.05 k u)/2 on a geometric u grid.

    skip_pairs: iterable of (ig, id) left without a table (-> gas factor 1, like a missing .tab file).
    axis_jitter: give every channel slightly different (p,T) axes (exercises the generic kernel).
    gas_axis_shift: give every gas its own (channel-independent) (p,T) grid (one table cell per gas and segment).
    Values are float32 exactly as stored in tbl_t (real_tblND_t, src/jurassic.h:387).
    """
    ng, nd = ctl.ng, ctl.nd
    NPx, NTx = P_AXIS.size, T_AXIS.size
    dim_u = dim_u or 200
    tbl = Tables(ng, nd, NPx, NTx, dim_u)
    skip = set(skip_pairs)
    ug = _U_GRID
    for ig, gas in enumerate(ctl.emitters):
        for d in range(nd):
            if (ig, d) in skip:
                continue
            pax = P_AXIS * (1.0 + (1e-3 * ((d * 7 + ig) % 5) if axis_jitter else 0.0)) * (1.0 + (0.07 * ig if gas_axis_shift else 0.0))
            tax = T_AXIS + (0.25 * ((d + ig) % 3) if axis_jitter else 0.0) + (1.5 * ig if gas_axis_shift else 0.0)
            kap = kappa0(gas, ig, d) * (0.3 + 0.7 * (pax[:, None] / 1013.25) ** 0.6) * (1.0 + 0.004 * (tax[None, :] - 250.0))
            ku = kap[:, :, None] * ug[None, None, :]
            eps = (1.0 - 0.5 * np.exp(-ku) - 0.5 * np.exp(-0.05 * ku)).astype(np.float32)
            start = np.argmax(eps > np.float32(1e-7), axis=2)
            end = np.argmax(eps > np.float32(0.99999), axis=2)  # first saturated row is kept
            nu_col = np.minimum(end - start + 1, dim_u)
            idx = start[:, :, None] + np.arange(dim_u)[None, None, :]
            valid = np.arange(dim_u)[None, None, :] < nu_col[:, :, None]
            idxc = np.minimum(idx, ug.size - 1)
            e_sel = np.take_along_axis(eps, idxc, axis=2)
            u_sel = ug.astype(np.float32)[idxc]
            tbl.np[ig, d] = NPx
            tbl.nt[ig, :, d] = NTx
            tbl.nu[ig, :, :, d] = nu_col
            tbl.p[ig, :, d] = pax
            tbl.t[ig, :, :, d] = tax[None, :]
            tbl.u[ig, :, :, :, d] = np.where(valid, u_sel, 0)
            tbl.eps[ig, :, :, :, d] = np.where(valid, e_sel, 0)
    tbl.sr[:, :] = source_table(ctl.nu)
    return tbl


def write_ascii_tables(ctl, tbl, directory, base):
    """Write tbl as the reference's ASCII inputs <base>_<nu %.4f>_<GAS>.tab and <base>_<nu %.4f>.filt
    (formats: src/jurassic.c:337,355-388 and :651-655).  %.9g reproduces every float32 exactly."""
    os.makedirs(directory, exist_ok=True)
    for d in range(ctl.nd):
        nu, f = boxcar_filter(ctl.nu[d])
        with open(os.path.join(directory, "%s_%.4f.filt" % (base, ctl.nu[d])), "w") as fh:
            fh.write("# $1 = wavenumber [cm^-1]\n# $2 = filter function\n\n")
            for a, b in zip(nu, f):
                fh.write("%.4f %g\n" % (a, b))
        for ig, gas in enumerate(ctl.emitters):
            if tbl.np[ig, d] < 1:
                continue
            with open(os.path.join(directory, "%s_%.4f_%s.tab" % (base, ctl.nu[d], gas)), "w") as fh:
                fh.write("# $1 = pressure [hPa]\n# $2 = temperature [K]\n# $3 = column density [molecules/cm^2]\n# $4 = emissivity\n")
                for ip in range(tbl.np[ig, d]):
                    for it in range(tbl.nt[ig, ip, d]):
                        fh.write("\n")
                        n = tbl.nu[ig, ip, it, d]
                        p, t = tbl.p[ig, ip, d], tbl.t[ig, ip, it, d]
                        for iu in range(n):
                            fh.write("%.17g %.17g %.9g %.9g\n" % (p, t, tbl.u[ig, ip, it, iu, d], tbl.eps[ig, ip, it, iu, d]))
    return os.path.join(directory, base)


# ---------------------------------------------------------------------------------------------------------------
# ASCII inputs of the reference's command-line tools (formats: SURVEY.md Appendix B)
def write_ctl(ctl, path, tblbase, extra=()):
    """KEY = VALUE control file as parsed by scan_ctl (src/jurassic.c:1153-1201)."""
    with open(path, "w") as f:
        f.write("TBLBASE = %s\nNG = %d\n" % (tblbase, ctl.ng))
        for i, e in enumerate(ctl.emitters):
            f.write("EMITTER[%d] = %s\n" % (i, e))
        f.write("ND = %d\n" % ctl.nd)
        for i, nu in enumerate(ctl.nu):
            f.write("NU[%d] = %.4f\n" % (i, nu))
        f.write("REFRAC = %d\nRAYDS = %g\nRAYDZ = %g\nHYDZ = %g\nWRITE_BBT = %d\n" % (ctl.refrac, ctl.rayds, ctl.raydz, ctl.hydz, ctl.write_bbt))
        f.write("READ_BINARY = 0\nWRITE_BINARY = 0\n")
        for k, v in extra:
            f.write("%s = %s\n" % (k, v))


def write_obs_tab(pkg, path):
    """read_obs format (src/jurassic.c:1041-1068): 10 geometry columns + rad[nd] + tau[nd] per ray."""
    with open(path, "w") as f:
        for r in range(pkg.n_rays):
            cols = [pkg.time[r], pkg.obsz[r], pkg.obslon[r], pkg.obslat[r], pkg.vpz[r], pkg.vplon[r], pkg.vplat[r], 0, 0, 0]
            f.write(" ".join("%.17g" % c for c in cols) + " " + " ".join(["0"] * (2 * pkg.nd)) + "\n")


def write_atm_tab(pkg, path):
    """read_atm format (src/jurassic.c:882-916): time z lon lat p T q[ng] k[nw]."""
    with open(path, "w") as f:
        for i in range(pkg.n_atm):
            cols = [pkg.atm_time[i], pkg.z[i], pkg.lon[i], pkg.lat[i], pkg.p[i], pkg.t[i]] + list(pkg.q[:pkg.ng, i]) + list(pkg.k[:pkg.nw, i])
            f.write(" ".join("%.17g" % c for c in cols) + "\n")
