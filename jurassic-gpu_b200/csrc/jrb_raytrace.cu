// jrb_raytrace.cu -- line-of-sight ray tracer + column densities, sm_100a.
//
// Computes what the reference's traceray() computes (src/jr_common.h:585-711: profile selection :127-154,
// altitude range :411-420, observer/view-point rejection :598-599, entry search :610-621, stepping loop
// :624-691 with refraction :664-681, tangent point :502-539, trapezoid rule :437-443, column density
// :446-453), split by what is sequential and what is not:
//
//   atm_slopes_kernel    per atmosphere level: d ln p / dz of the exponential pressure interpolation (eip, :53-57) and dT/dz,
//                        so that the stepping loop needs one exp but no log and no division per interpolation;
//   ray_step_kernel      one thread per ray, the inherently sequential part: walk the ray, per point altitude,
//                        p, T (5 interpolations per step with refraction), raw step length, Cartesian position.
//                        Longitude/latitude (asin, atan2) are not needed along the way -- the 1-D atmosphere is a
//                        function of altitude only -- they are evaluated for the tangent point only;
//   los_finalize_kernel  one thread per (ray, segment), fully parallel: vmr/extinction interpolation, trapezoid
//                        segment lengths, column densities and -- for tables whose (p,T) axes do not depend on the
//                        channel -- the table cell and interpolation weights per gas, so that the EGA kernel does
//                        not search axes per channel.
//
// 2-D / 3-D atmospheres (ctl->ip = 2, 3; intpol_atm_2d / _3d, src/jurassic.c:704-804) go through ray_geo_kernel instead of
// ray_step_kernel: there the atmosphere depends on longitude and latitude as well, so every evaluation point is converted
// to geographic coordinates and interpolated between the two nearest columns (2-D) or averaged over the points inside the
// influence sphere (3-D), as the dispatch of src/jurassic.c:685-691 does.  The reference's own formod() stops at an assert
// for ip != 1 (src/jr_common.h:573,581); what is computed here is its tracer with those two calls replaced by that dispatch,
// applied to the atmosphere slice the ray's time selects (oracle/jr_oracle.c restates exactly this).
#include "jrb_internal.h"
#include <jurassic_b200.h> // JRB_MAX_NG, JRB_MAX_NW
#include <cstdlib>

namespace jrb {

namespace {

__device__ __forceinline__ double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

__device__ __forceinline__ void geo_to_cart(double alt, double lon, double lat, double x[3]) {
  const double d2r = M_PI / 180.0;
  const double radius = alt + kRE, clat = cos(lat * d2r);
  x[0] = radius * clat * cos(lon * d2r);
  x[1] = radius * clat * sin(lon * d2r);
  x[2] = radius * sin(lat * d2r);
}
__device__ __forceinline__ void cart_to_lonlat(const double x[3], double *lon, double *lat) {
  const double r2d = 180.0 / M_PI;
  *lat = asin(x[2] / norm3(x)) * r2d;
  *lon = atan2(x[1], x[0]) * r2d;
}

// One vertical profile as seen by a ray.  `locate` of the reference (src/jr_common.h:87-104) returns, for an
// ascending axis, max{i <= n-2 : z[i] <= x} (0 if x < z[0]); for a descending one max{i <= n-2 : z[i] > x}.
// The level found for the previous evaluation is tried first; the bisection runs only when it does not bracket x.
struct Profile {
  const double *__restrict__ z, *__restrict__ p, *__restrict__ t, *__restrict__ pslope, *__restrict__ tslope;
  int n;
  bool asc;

  __device__ __forceinline__ bool brackets(double x, int i) const {
    const double zl = z[i], zh = z[i + 1];
    return asc ? ((zl <= x || i == 0) && (zh > x || i == n - 2)) : ((zl > x || i == 0) && (zh <= x || i == n - 2));
  }
  __device__ __forceinline__ int locate(double x, int hint) const {
    if (brackets(x, hint)) return hint;
    // a step moves the point by at most one level in practice: try the neighbour before bisecting
    const bool up = asc ? (x >= z[hint + 1]) : (x < z[hint + 1]);
    const int nb = up ? min(hint + 1, n - 2) : max(hint - 1, 0);
    if (brackets(x, nb)) return nb;
    int ilo = 0, ihi = n - 1;
    if (asc) { while (ihi > ilo + 1) { const int i = (ihi + ilo) >> 1; if (z[i] > x) ihi = i; else ilo = i; } }
    else     { while (ihi > ilo + 1) { const int i = (ihi + ilo) >> 1; if (z[i] <= x) ihi = i; else ilo = i; } }
    return ilo;
  }
  // intpol_atm_1d_pt (:549-555): p exponential (eip :53-57), T linear (lip :48-50), with per-level slopes
  __device__ __forceinline__ void eval(double x, int level, double *pp, double *tt) const {
    const double dx = x - z[level], s = pslope[level];
    if (s == s) *pp = p[level] * exp(s * dx);                                     // both pressures positive
    else *pp = p[level] + dx * (p[level + 1] - p[level]) / (z[level + 1] - z[level]); // linear fallback of eip
    *tt = fma(dx, tslope[level], t[level]);
  }
  __device__ __forceinline__ void pt(double x, int &level, double *pp, double *tt) const {
    level = locate(x, level);
    eval(x, level, pp, tt);
  }
};

__device__ __forceinline__ double refractivity(double p, double t) { return 7.753e-05 * p / t; }
// the stepping loop of ray_step_kernel is one long dependent FP64 chain per ray: divisions and square roots there use the
// Newton forms of jrb_device.cuh (1e-16 relative, no special-case branches) -- about a third of the chain
// (explicitly rounded products: the two forms of the kernel use these values in different expressions, and a product that the
//  compiler contracts into a following add in one form only would break their bit-identity)
__device__ __forceinline__ double refractivity_fast(double p, double t) { return __dmul_rn(7.753e-05 * p, fast_rcp(t)); }
__device__ __forceinline__ double fast_sqrt(double x) { return __dmul_rn(x, fast_rsqrt(x)); }

} // namespace

__global__ void atm_slopes_kernel(const double *__restrict__ z, const double *__restrict__ p, const double *__restrict__ t,
                                  double *__restrict__ pslope, double *__restrict__ tslope, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sp = __longlong_as_double(0x7ff8000000000000ll), st = 0.0;
  if (i + 1 < n) {
    const double y0 = p[i], y1 = p[i + 1], dz = z[i + 1] - z[i];
    if (y0 > 0 && y1 > 0) sp = log(y1 / y0) / dz;
    st = (t[i + 1] - t[i]) / dz;
  }
  pslope[i] = sp;
  tslope[i] = st;
}

// LPR = lanes per ray.  1: one thread walks one ray (throughput form, large batches).  8: a group of 8 lanes walks one
// ray together -- the five (p,T) evaluations of a step (the point itself and the four probe points of the refractivity
// gradient) are independent and run on five lanes side by side, everything else is computed redundantly by all lanes of
// the group; the per-step instruction count drops ~3x and a batch has 8x more warps.  That is what a small batch needs
// (a single 1088-ray package is 34 warps in the throughput form: one warp per scheduler on 9 SMs, every dependent FP64
// instruction exposed).  Every evaluation uses the same expressions in both forms: bit-identical results.
template <int LPR>
__global__ void __launch_bounds__(128) ray_step_kernel(TraceArgs a) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = gtid / LPR;
  const int role = LPR > 1 ? (int)(threadIdx.x & (LPR - 1)) : 0;                                   // lane within the ray's group
  const unsigned gmask = LPR > 1 ? ((LPR >= 32 ? 0xffffffffu : ((1u << LPR) - 1u)) << ((threadIdx.x & 31) & ~(LPR - 1))) : 0u; // the group's lanes
  if (r >= a.n_rays) return;

  double *__restrict__ raw0 = a.raw + (size_t)r * kNLOS * kRaw; // this ray's raw points

  const int pk = a.ray_pkg[r];
  const long long abase = a.pkg_atm_off[pk];
  const int anp = a.pkg_atm_np[pk];
  const double *__restrict__ atime = a.atm_time + abase;

  const double obsz = a.geo[0 * a.geo_stride + r], obslon = a.geo[1 * a.geo_stride + r],
               obslat = a.geo[2 * a.geo_stride + r], vpz = a.geo[3 * a.geo_stride + r],
               vplon = a.geo[4 * a.geo_stride + r], vplat = a.geo[5 * a.geo_stride + r],
               rtime = a.geo[6 * a.geo_stride + r];

  // defaults (src/jr_common.h:589-594)
  double tsurf = -999.0;
  double tpz = vpz, tplon = vplon, tplat = vplat;
  int np = 0;

  // ---- profile selection by time (locate_atm, :127-154) ----
  int lo = 0, hi = anp - 1;
  while (hi > lo + 1) { int i = (lo + hi) / 2; if (atime[i] < rtime) lo = i; else hi = i; }
  const int lower = (0 == lo) ? lo : hi;
  lo = lower; hi = anp - 1;
  while (hi > lo + 1) { int i = (lo + hi) / 2; if (atime[i] > rtime) hi = i; else lo = i; }
  const int upper = (hi == anp - 1) ? anp : hi;
  Profile P;
  P.n = upper - lower;
  P.z = a.atm_z + abase + lower;
  P.p = a.atm_p + abase + lower;
  P.t = a.atm_t + abase + lower;
  P.pslope = a.atm_lnp_slope + abase + lower;
  P.tslope = a.atm_lnp_slope + a.atm_stride + abase + lower;
  { const int m = (P.n - 1) >> 1; P.asc = P.z[m] < P.z[m + 1]; }
  const double *__restrict__ alon = a.atm_lon + abase + lower;
  const double *__restrict__ alat = a.atm_lat + abase + lower;

  // ---- altitude range of the profile (altitude_range_nn, :411-420) ----
  double zmin = P.z[0], zmax = P.z[0];
  for (int i = 0; i < P.n && alon[i] == alon[0] && alat[i] == alat[0]; ++i) {
    zmax = fmax(zmax, P.z[i]);
    zmin = fmin(zmin, P.z[i]);
  }

  // (a ray whose time selects fewer than two levels -- no profile matches -- is undefined behaviour in the reference,
  //  src/jr_common.h:640; here it is rejected like a ray that misses the atmosphere: np = 0, rad = 0, tau = 1)
  const bool rejected = (obsz < zmin) || (vpz > zmax - 0.001) || (P.n < 2);
  int z_low_idx = -1;
  if (!rejected) {
    double xobs[3], xvp[3], ex0[3], x[3];
    geo_to_cart(obsz, obslon, obslat, xobs);
    geo_to_cart(vpz, vplon, vplat, xvp);
    for (int i = 0; i < 3; i++) ex0[i] = xvp[i] - xobs[i];
    const double norm = norm3(ex0);
    for (int i = 0; i < 3; i++) { ex0[i] /= norm; x[i] = xobs[i]; }
    double z = 1e99;
    if (obsz > zmax) { // observer above the atmosphere: bisect for the entry point (:610-621)
      double dmax = norm, dmin = 0.0;
      while (fabs(dmin - dmax) > 0.001) {
        const double d = 0.5 * (dmax + dmin);
        for (int i = 0; i < 3; i++) x[i] = xobs[i] + d * ex0[i];
        z = norm3(x) - kRE;
        if ((z <= zmax) && (z > zmax - 0.001)) break;
        if (z < zmax - 0.0005) dmax = d; else dmin = d;
      }
    }

    double z_low = 1e99, p, t;
    double xprev[3] = {0, 0, 0}, zprev = 0; // previous LOS point
    int level = 0, stop = 0;
    for (; np < kNLOS; ++np) {
      const double rr = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
      const double inv = fast_rsqrt(rr), rn = __dmul_rn(rr, inv); // |x| and 1/|x| from one chain
      double ds = a.rayds;
      if (a.raydz > 0.0) { // step length from the angle to the local vertical (:625-635)
        double dot = 0.0;
        for (int i = 0; i < 3; i++) dot += ex0[i] * x[i] * inv;
        const double cosa = fabs(dot);
        if (cosa != 0.0) ds = fmin(ds, a.raydz * fast_rcp(cosa));
      }
      z = rn - kRE;
      if ((z < zmin) || (z > zmax)) { // left the atmosphere: clip the last segment (:637-648)
        if (np == 0) break;           // (the reference would read los[-1]; unreachable after the entry search)
        stop = (z < zmin) ? 2 : 1;
        const double zfrac = (z < zmin) ? zmin : zmax;
        const double frac = (zfrac - zprev) / (z - zprev);
        for (int i = 0; i < 3; i++) x[i] = xprev[i] + frac * (x[i] - xprev[i]);
        z = norm3(x) - kRE;
        if (role == 0) raw0[(size_t)(np - 1) * kRaw + kRawTail + LT_DSRAW] = ds * frac;
        ds = 0.0;
      }
      double nref = 1.0, ngr[3] = {0.0, 0.0, 0.0};
      const double h = 0.02;
      if (LPR == 1) {
        P.pt(z, level, &p, &t);
      } else {
        // lane 0 of the group: the point itself; lane 1: the half-step point; lanes 2..4: half step + h along each axis.
        // (the probe points are only needed below 60 km with refraction on and not at the last point; evaluating them
        //  anyway costs nothing -- the lanes would idle -- and keeps the group convergent)
        double zj = z;
        if (role >= 1 && role <= 4) {
          double xh[3];
          for (int i = 0; i < 3; i++) xh[i] = x[i] + 0.5 * ds * ex0[i];
          const double r2 = xh[0] * xh[0] + xh[1] * xh[1] + xh[2] * xh[2];
          const double xi = role == 2 ? xh[0] : (role == 3 ? xh[1] : xh[2]);
          const double s2 = role == 1 ? r2 : fma(h, fma(2.0, xi, h), r2);
          zj = fast_sqrt(s2) - kRE;
        }
        const int lvj = P.locate(zj, level);
        double pj, tj;
        P.eval(zj, lvj, &pj, &tj);
        const double nj = refractivity_fast(pj, tj);
        p = __shfl_sync(gmask, pj, 0, LPR); t = __shfl_sync(gmask, tj, 0, LPR);
        level = __shfl_sync(gmask, lvj, 0, LPR);
        const double n2 = __shfl_sync(gmask, nj, 1, LPR);
        const double g0 = __shfl_sync(gmask, nj, 2, LPR), g1 = __shfl_sync(gmask, nj, 3, LPR), g2 = __shfl_sync(gmask, nj, 4, LPR);
        if (a.refrac && z <= 60.0) {
          nref += refractivity_fast(p, t);
          ngr[0] = (g0 - n2) * (1.0 / h); ngr[1] = (g1 - n2) * (1.0 / h); ngr[2] = (g2 - n2) * (1.0 / h);
        }
      }
      if (role == 0) { // one 64-byte raw point: {p, t, z, ds} {level, x, y, z}
        double4 *__restrict__ pt = reinterpret_cast<double4 *>(raw0 + (size_t)np * kRaw);
        pt[0] = make_double4(p, t, z, ds);
        pt[1] = make_double4((double)level, x[0], x[1], x[2]);
      }
      for (int i = 0; i < 3; i++) xprev[i] = x[i];
      zprev = z;
      if (z < z_low) { z_low = z; z_low_idx = np; }

      if (stop) { tsurf = (stop == 2 ? t : -999.0); break; }

      if (LPR == 1 && a.refrac && z <= 60.0) { // refractivity gradient by finite differences at the half step (:664-681)
        nref += refractivity_fast(p, t);
        // the four probe points (half step, and half step + h along each axis) are independent: altitudes, level
        // searches and the exponentials are evaluated side by side
        double xh[3], zz[4], ph[4], th[4];
        int lv[4];
        for (int i = 0; i < 3; i++) xh[i] = x[i] + 0.5 * ds * ex0[i];
        const double r2 = xh[0] * xh[0] + xh[1] * xh[1] + xh[2] * xh[2];
        zz[0] = fast_sqrt(r2) - kRE;
#pragma unroll
        for (int i = 0; i < 3; i++) { const double s2 = fma(h, fma(2.0, xh[i], h), r2); zz[1 + i] = fast_sqrt(s2) - kRE; } // |xh + h e_i|
#pragma unroll
        for (int j = 0; j < 4; j++) lv[j] = P.locate(zz[j], level);
#pragma unroll
        for (int j = 0; j < 4; j++) P.eval(zz[j], lv[j], &ph[j], &th[j]);
        const double n2 = refractivity_fast(ph[0], th[0]);
#pragma unroll
        for (int i = 0; i < 3; i++) ngr[i] = (refractivity_fast(ph[1 + i], th[1 + i]) - n2) * (1.0 / h);
      }
      double ex1[3];
      for (int i = 0; i < 3; i++) ex1[i] = ex0[i] * nref + ds * ngr[i];
      const double in1 = fast_rsqrt(ex1[0] * ex1[0] + ex1[1] * ex1[1] + ex1[2] * ex1[2]);
      for (int i = 0; i < 3; i++) {
        ex1[i] *= in1;
        x[i] += 0.5 * ds * (ex0[i] + ex1[i]);
        ex0[i] = ex1[i];
      }
    }
    if (np > 0 || stop) ++np;   // the reference increments after the loop (:692)
    // the reference's CPU path is fatal here ("Too many LOS points!" when NLOS <= np, src/jr_common.h:693-695): reported
    // through the error word, the run then fails instead of returning a truncated ray
    if (np >= kNLOS && a.error_flag) a.error_flag[0] = 1;
    if (np > kNLOS) np = kNLOS;
  }

  if (role != 0) return; // the rest is done by the first lane of the group (it wrote the records it reads back here)
  // ---- tangent point (from the raw step lengths; :502-539) ----
  if (np > 0) {
    const int ip = z_low_idx;
    if (ip <= 0 || ip >= np - 1) { // nadir or zenith: last point
      const double *tl = raw0 + (size_t)(np - 1) * kRaw + kRawTail;
      tpz = tl[LT_Z];
      cart_to_lonlat(tl + LT_X, &tplon, &tplat);
    } else {
      const double *t0 = raw0 + (size_t)(ip - 1) * kRaw + kRawTail, *t1 = t0 + kRaw, *t2 = t1 + kRaw;
      const double yy0 = t0[LT_Z], yy1 = t1[LT_Z], yy2 = t2[LT_Z];
      const double ds0 = t1[LT_DSRAW], ds1 = t2[LT_DSRAW];
      const double dyy10 = yy1 - yy0, dyy21 = yy2 - yy1;
      const double x1 = sqrt(ds0 * ds0 - dyy10 * dyy10);
      const double x2 = x1 + sqrt(ds1 * ds1 - dyy21 * dyy21);
      const double dx12 = x1 - x2;
      const double qa = (dyy10 * x2 + (yy0 - yy2) * x1) / (x1 * x2 * dx12);
      const double qb = dyy10 / x1 - qa * x1;
      const double xt = -qb / (2 * qa);
      tpz = (qa * xt + qb) * xt + yy0;
      double v[3];
      for (int i = 0; i < 3; i++) v[i] = lerp_div(0.0, t0[LT_X + i], x2, t2[LT_X + i], xt);
      cart_to_lonlat(v, &tplon, &tplat);
    }
  }

  a.ray_np[r] = np;
  a.ray_tsurf[r] = tsurf;
  a.ray_level0[r] = lower;
  a.tp[0 * a.geo_stride + r] = tpz;
  a.tp[1 * a.geo_stride + r] = tplon;
  a.tp[2 * a.geo_stride + r] = tplat;
  if (a.tp_host) { // the caller's obs_t (or the pinned result buffer), host-mapped: no copy phase afterwards
    *a.tp_host[0 * a.geo_stride + r] = tpz;
    *a.tp_host[1 * a.geo_stride + r] = tplon;
    *a.tp_host[2 * a.geo_stride + r] = tplat;
  }
}

// ---- 2-D / 3-D atmospheres ------------------------------------------------------------------------------------------------
// per atmosphere point: Cartesian position at altitude 0 (geo2cart(0, lon, lat), the x1[] of intpol_atm_2d / _3d) and the index
// of the next point whose (lon, lat) differs (the end of the column the point belongs to)
__global__ void atm_geo_kernel(const double *__restrict__ lon, const double *__restrict__ lat, double *__restrict__ cart,
                               int *__restrict__ next, long long stride, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[3];
  geo_to_cart(0.0, lon[i], lat[i], x);
  cart[i] = x[0]; cart[stride + i] = x[1]; cart[2 * stride + i] = x[2];
  long long j = i + 1;
  while (j < n && lon[j] == lon[i] && lat[j] == lat[i]) ++j;
  next[i] = (int)j;
}

namespace {

__device__ __forceinline__ double dist2(const double a[3], double b0, double b1, double b2) {
  return (a[0] - b0) * (a[0] - b0) + (a[1] - b1) * (a[1] - b1) + (a[2] - b2) * (a[2] - b2);
}

// locate (src/jr_common.h:87-104) on an ascending or descending axis
__device__ __forceinline__ int locate_plain(const double *__restrict__ xx, int n, double x) {
  int ilo = 0, ihi = n - 1, i = (n - 1) >> 1;
  if (xx[i] < xx[i + 1]) { while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (xx[i] > x) ihi = i; else ilo = i; } }
  else                   { while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (xx[i] <= x) ihi = i; else ilo = i; } }
  return ilo;
}

struct GeoAtm {
  const TraceArgs &a;
  long long lo, hi; // the ray's atmosphere slice (global point indices)
  int ng, nw;

  // intpol_atm_1d (src/jurassic.c:694-701) in the column [s, s+n); q/k only where wanted (the probe points need p, T)
  __device__ void column(long long s, int n, double z0, double *p, double *t, double *q, double *k, bool qk) const {
    const long long i = s + locate_plain(a.atm_z + s, n, z0);
    const double x0 = a.atm_z[i], x1 = a.atm_z[i + 1], y0 = a.atm_p[i], y1 = a.atm_p[i + 1];
    *p = (y0 > 0 && y1 > 0) ? y0 * exp(log(y1 / y0) / (x1 - x0) * (z0 - x0)) : lerp_div(x0, y0, x1, y1, z0);
    *t = lerp_div(x0, a.atm_t[i], x1, a.atm_t[i + 1], z0);
    if (!qk) return;
    for (int ig = 0; ig < ng; ig++) { const double *v = a.atm_q + (size_t)ig * a.atm_stride + i; q[ig] = lerp_div(x0, v[0], x1, v[1], z0); }
    for (int iw = 0; iw < nw; iw++) { const double *v = a.atm_k + (size_t)iw * a.atm_stride + i; k[iw] = lerp_div(x0, v[0], x1, v[1], z0); }
  }

  // intpol_atm_2d (src/jurassic.c:704-760): the two nearest columns, blended by the projected position between them
  __device__ void eval2d(double z0, double lon0, double lat0, double *p, double *t, double *q, double *k, bool qk) const {
    const double dlat = 10;
    double x0[3], dhmin0 = 1e99, dhmin1 = 1e99;
    geo_to_cart(0.0, lon0, lat0, x0);
    const long long first_end = min((long long)a.atm_next[lo], hi);
    long long s0 = lo, s1 = lo, e0 = first_end, e1 = first_end; // ix0 = ix1 = 0 (:710)
    for (long long s = lo; s < hi;) {
      const long long e = min((long long)a.atm_next[s], hi);
      if (fabs(lat0 - a.atm_lat[s]) <= dlat) {
        const double dh = dist2(x0, a.atm_cart[s], a.atm_cart[a.atm_stride + s], a.atm_cart[2 * a.atm_stride + s]);
        if (dh <= dhmin0) { dhmin1 = dhmin0; s1 = s0; e1 = e0; dhmin0 = dh; s0 = s; e0 = e; }
        else if (dh <= dhmin1) { dhmin1 = dh; s1 = s; e1 = e; }
      }
      s = e;
    }
    double p0, p1, t0, t1, q1[JRB_MAX_NG], k1[JRB_MAX_NW];
    column(s0, (int)(e0 - s0), z0, &p0, &t0, q, k, qk);
    column(s1, (int)(e1 - s1), z0, &p1, &t1, q1, k1, qk);
    const double xa[3] = {a.atm_cart[s0], a.atm_cart[a.atm_stride + s0], a.atm_cart[2 * a.atm_stride + s0]};
    const double x2 = dist2(xa, a.atm_cart[s1], a.atm_cart[a.atm_stride + s1], a.atm_cart[2 * a.atm_stride + s1]);
    const double x = sqrt(x2), r0 = (dhmin0 - dhmin1 + x2) / (2 * x), r1 = x - r0;
    double r;
    if (r0 <= 0) r = 0; else r = (r1 <= 0) ? 1 : r0 / (r0 + r1);
    *p = (1 - r) * p0 + r * p1;
    *t = (1 - r) * t0 + r * t1;
    if (!qk) return;
    for (int ig = 0; ig < ng; ig++) q[ig] = (1 - r) * q[ig] + r * q1[ig];
    for (int iw = 0; iw < nw; iw++) k[iw] = (1 - r) * k[iw] + r * k1[iw];
  }

  // intpol_atm_3d (src/jurassic.c:763-804): distance-weighted average over the points inside the influence sphere
  __device__ void eval3d(double z0, double lon0, double lat0, double *p, double *t, double *q, double *k, bool qk) const {
    const double rm2 = a.cx * a.cx;
    double x0[3], wsum = 0, ps = 0, ts = 0;
    geo_to_cart(0.0, lon0, lat0, x0);
    if (qk) { for (int ig = 0; ig < ng; ig++) q[ig] = 0; for (int iw = 0; iw < nw; iw++) k[iw] = 0; }
    for (long long i = lo; i < hi; i++) {
      const double dz = fabs(a.atm_z[i] - z0);
      if (dz >= a.cz) continue;
      if (fabs(a.atm_lat[i] - lat0) * 111.13 >= a.cx) continue;
      const double dx2 = dist2(x0, a.atm_cart[i], a.atm_cart[a.atm_stride + i], a.atm_cart[2 * a.atm_stride + i]);
      if (dx2 >= rm2) continue;
      const double w = (1 - dz / a.cz) * (rm2 - dx2) / (rm2 + dx2);
      wsum += w;
      ps += w * a.atm_p[i];
      ts += w * a.atm_t[i];
      if (qk) {
        for (int ig = 0; ig < ng; ig++) q[ig] += w * a.atm_q[(size_t)ig * a.atm_stride + i];
        for (int iw = 0; iw < nw; iw++) k[iw] += w * a.atm_k[(size_t)iw * a.atm_stride + i];
      }
    }
    if (wsum >= 1e-6) {
      *p = ps / wsum; *t = ts / wsum;
      if (qk) { for (int ig = 0; ig < ng; ig++) q[ig] /= wsum; for (int iw = 0; iw < nw; iw++) k[iw] /= wsum; }
    } else {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      *p = *t = nan;
      if (qk) { for (int ig = 0; ig < ng; ig++) q[ig] = nan; for (int iw = 0; iw < nw; iw++) k[iw] = nan; }
    }
  }

  __device__ void eval(double z0, double lon0, double lat0, double *p, double *t, double *q, double *k, bool qk) const {
    if (a.ip == 2) eval2d(z0, lon0, lat0, p, t, q, k, qk); else eval3d(z0, lon0, lat0, p, t, q, k, qk);
  }
};

__device__ __forceinline__ void cart_to_geo(const double x[3], double *alt, double *lon, double *lat) {
  const double radius = norm3(x), r2d = 180.0 / M_PI;
  *lat = asin(x[2] / radius) * r2d;
  *lon = atan2(x[1], x[0]) * r2d;
  *alt = radius - kRE;
}

} // namespace

// One thread per ray; the stepping loop of traceray (src/jr_common.h:585-711) over intpol_atm_geo (src/jurassic.c:685-691).
// Writes the complete record of every point (p, T, extinction, vmr in the column-density slots, altitude, raw step length,
// position); los_finalize_kernel then only applies the trapezoid rule, the column densities and the table cells.
__global__ void __launch_bounds__(128) ray_geo_kernel(TraceArgs a) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_rays) return;
  const LosLayout L = a.los;
  double *__restrict__ rec0 = a.los_data + (size_t)r * kNLOS * L.rec;
  double *__restrict__ raw0 = a.raw + (size_t)r * kNLOS * kRaw;
  const int pk = a.ray_pkg[r];
  const long long abase = a.pkg_atm_off[pk];
  const int anp = a.pkg_atm_np[pk];
  const double *__restrict__ atime = a.atm_time + abase;
  const double obsz = a.geo[0 * a.geo_stride + r], obslon = a.geo[1 * a.geo_stride + r],
               obslat = a.geo[2 * a.geo_stride + r], vpz = a.geo[3 * a.geo_stride + r],
               vplon = a.geo[4 * a.geo_stride + r], vplat = a.geo[5 * a.geo_stride + r],
               rtime = a.geo[6 * a.geo_stride + r];
  double tsurf = -999.0, tpz = vpz, tplon = vplon, tplat = vplat;
  int np = 0;
  // slice of the atmosphere with the ray's time stamp (locate_atm, :127-154)
  int lo = 0, hi = anp - 1;
  while (hi > lo + 1) { int i = (lo + hi) / 2; if (atime[i] < rtime) lo = i; else hi = i; }
  const int lower = (0 == lo) ? lo : hi;
  lo = lower; hi = anp - 1;
  while (hi > lo + 1) { int i = (lo + hi) / 2; if (atime[i] > rtime) hi = i; else lo = i; }
  const int upper = (hi == anp - 1) ? anp : hi;
  const GeoAtm G{a, abase + lower, abase + upper, L.ng, L.nw};
  const int n = upper - lower;
  // altitude range of the first column (altitude_range_nn, :411-420)
  double zmin = a.atm_z[G.lo], zmax = zmin;
  for (long long i = G.lo; i < G.hi && a.atm_lon[i] == a.atm_lon[G.lo] && a.atm_lat[i] == a.atm_lat[G.lo]; ++i) {
    zmax = fmax(zmax, a.atm_z[i]);
    zmin = fmin(zmin, a.atm_z[i]);
  }
  bool rejected = (obsz < zmin) || (vpz > zmax - 0.001) || (n < 2);
  if (a.ip == 2 && n >= 2) { // the profile list checks of intpol_atm_2d (:726-729), fatal in the reference
    double lat_prev = 0;
    for (long long s = G.lo; s < G.hi;) {
      const long long e = min((long long)a.atm_next[s], G.hi);
      if (e - s <= 1) { if (a.error_flag) a.error_flag[1] = 1; rejected = true; }
      if (s > G.lo && fabs(lat_prev - a.atm_lat[s]) > 10) { if (a.error_flag) a.error_flag[2] = 1; rejected = true; }
      lat_prev = a.atm_lat[s];
      s = e;
    }
  }
  int z_low_idx = -1;
  if (!rejected) {
    double xobs[3], xvp[3], ex0[3], x[3], q[JRB_MAX_NG], k[JRB_MAX_NW];
    geo_to_cart(obsz, obslon, obslat, xobs);
    geo_to_cart(vpz, vplon, vplat, xvp);
    for (int i = 0; i < 3; i++) ex0[i] = xvp[i] - xobs[i];
    const double norm = norm3(ex0);
    for (int i = 0; i < 3; i++) { ex0[i] /= norm; x[i] = xobs[i]; }
    double z = 1e99, lon, lat;
    if (obsz > zmax) { // entry point (:610-621)
      double dmax = norm, dmin = 0.0;
      while (fabs(dmin - dmax) > 0.001) {
        const double d = 0.5 * (dmax + dmin);
        for (int i = 0; i < 3; i++) x[i] = xobs[i] + d * ex0[i];
        z = norm3(x) - kRE;
        if ((z <= zmax) && (z > zmax - 0.001)) break;
        if (z < zmax - 0.0005) dmax = d; else dmin = d;
      }
    }
    double z_low = 1e99, p, t, xprev[3] = {0, 0, 0}, zprev = 0;
    int stop = 0;
    for (; np < kNLOS; ++np) {
      double ds = a.rayds;
      if (a.raydz > 0.0) {
        const double inv = 1.0 / norm3(x);
        double dot = 0.0;
        for (int i = 0; i < 3; i++) dot += ex0[i] * x[i] * inv;
        const double cosa = fabs(dot);
        if (cosa != 0.0) ds = fmin(ds, a.raydz / cosa);
      }
      cart_to_geo(x, &z, &lon, &lat);
      if ((z < zmin) || (z > zmax)) {
        if (np == 0) break;
        stop = (z < zmin) ? 2 : 1;
        const double zfrac = (z < zmin) ? zmin : zmax;
        const double frac = (zfrac - zprev) / (z - zprev);
        for (int i = 0; i < 3; i++) x[i] = xprev[i] + frac * (x[i] - xprev[i]);
        cart_to_geo(x, &z, &lon, &lat);
        raw0[(size_t)(np - 1) * kRaw + kRawTail + LT_DSRAW] = ds * frac;
        ds = 0.0;
      }
      G.eval(z, lon, lat, &p, &t, q, k, true);
      {
        double *__restrict__ rec = rec0 + (size_t)np * L.rec; // extinction and vmr go straight into the record
        for (int iw = 0; iw < L.nw; iw++) rec[4 + iw] = k[iw];
        for (int ig = 0; ig < L.ng; ig++) rec[L.u0 + ig] = q[ig];
        double4 *__restrict__ pt = reinterpret_cast<double4 *>(raw0 + (size_t)np * kRaw);
        pt[0] = make_double4(p, t, z, ds);
        pt[1] = make_double4(0.0, x[0], x[1], x[2]);
      }
      for (int i = 0; i < 3; i++) xprev[i] = x[i];
      zprev = z;
      if (z < z_low) { z_low = z; z_low_idx = np; }
      if (stop) { tsurf = (stop == 2 ? t : -999.0); break; }
      double nref = 1.0, ngr[3] = {0.0, 0.0, 0.0};
      if (a.refrac && z <= 60.0) { // refractivity gradient at the half step (:664-681)
        nref += refractivity(p, t);
        double xh[3], ph, th, zh, lonh, lath;
        for (int i = 0; i < 3; i++) xh[i] = x[i] + 0.5 * ds * ex0[i];
        cart_to_geo(xh, &zh, &lonh, &lath);
        G.eval(zh, lonh, lath, &ph, &th, nullptr, nullptr, false);
        const double n2 = refractivity(ph, th), h = 0.02;
        for (int i = 0; i < 3; i++) {
          xh[i] += h;
          cart_to_geo(xh, &zh, &lonh, &lath);
          G.eval(zh, lonh, lath, &ph, &th, nullptr, nullptr, false);
          ngr[i] = (refractivity(ph, th) - n2) / h;
          xh[i] -= h;
        }
      }
      double ex1[3];
      for (int i = 0; i < 3; i++) ex1[i] = ex0[i] * nref + ds * ngr[i];
      const double n1 = norm3(ex1);
      for (int i = 0; i < 3; i++) {
        ex1[i] /= n1;
        x[i] += 0.5 * ds * (ex0[i] + ex1[i]);
        ex0[i] = ex1[i];
      }
    }
    if (np > 0 || stop) ++np;
    if (np >= kNLOS && a.error_flag) a.error_flag[0] = 1; // "Too many LOS points!" (:693-695)
    if (np > kNLOS) np = kNLOS;
  }
  if (np > 0) { // tangent point (:502-539)
    const int ip = z_low_idx;
    if (ip <= 0 || ip >= np - 1) {
      const double *tl = raw0 + (size_t)(np - 1) * kRaw + kRawTail;
      tpz = tl[LT_Z];
      cart_to_lonlat(tl + LT_X, &tplon, &tplat);
    } else {
      const double *t0 = raw0 + (size_t)(ip - 1) * kRaw + kRawTail, *t1 = t0 + kRaw, *t2 = t1 + kRaw;
      const double yy0 = t0[LT_Z], yy1 = t1[LT_Z], yy2 = t2[LT_Z], ds0 = t1[LT_DSRAW], ds1 = t2[LT_DSRAW];
      const double dyy10 = yy1 - yy0, dyy21 = yy2 - yy1, x1 = sqrt(ds0 * ds0 - dyy10 * dyy10),
                   x2 = x1 + sqrt(ds1 * ds1 - dyy21 * dyy21), dx12 = x1 - x2,
                   qa = (dyy10 * x2 + (yy0 - yy2) * x1) / (x1 * x2 * dx12), qb = dyy10 / x1 - qa * x1, xt = -qb / (2 * qa);
      tpz = (qa * xt + qb) * xt + yy0;
      double v[3];
      for (int i = 0; i < 3; i++) v[i] = lerp_div(0.0, t0[LT_X + i], x2, t2[LT_X + i], xt);
      cart_to_lonlat(v, &tplon, &tplat);
    }
  }
  a.ray_np[r] = np;
  a.ray_tsurf[r] = tsurf;
  a.ray_level0[r] = lower;
  a.tp[0 * a.geo_stride + r] = tpz; a.tp[1 * a.geo_stride + r] = tplon; a.tp[2 * a.geo_stride + r] = tplat;
  if (a.tp_host) {
    *a.tp_host[0 * a.geo_stride + r] = tpz; *a.tp_host[1 * a.geo_stride + r] = tplon; *a.tp_host[2 * a.geo_stride + r] = tplat;
  }
}

// thread per (ray, segment)
__global__ void __launch_bounds__(256) los_finalize_kernel(TraceArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = idx / kNLOS;
  const int ip = (int)(idx - r * kNLOS);
  if (r >= a.n_rays) return;
  const int np = a.ray_np[r];
  if (ip >= np) return;
  const LosLayout L = a.los;
  double *__restrict__ rec = a.los_data + ((size_t)r * kNLOS + ip) * L.rec;
  const double *__restrict__ pt = a.raw + ((size_t)r * kNLOS + ip) * kRaw;
  const double *__restrict__ tail = pt + kRawTail;
  const double p = pt[0], t = pt[1], z = tail[LT_Z];
  // trapezoid rule on the raw step lengths (:437-443)
  const double ds_raw = tail[LT_DSRAW];
  const double ds = (ip == 0) ? 0.5 * ds_raw : 0.5 * (tail[LT_DSRAW - kRaw] + ds_raw);
  rec[0] = p; rec[1] = t; rec[2] = ds;
  const double dens = 10. * p / (kBoltzmann * t) * ds; // column density per unit vmr (:446-453)
  double qh2o = 0.0;
  if (a.ip != 1) { // 2-D / 3-D atmosphere: ray_geo_kernel stored extinction and vmr of the point
    for (int ig = 0; ig < L.ng; ig++) {
      const double qv = rec[L.u0 + ig];
      if (ig == a.ig_h2o) qh2o = qv;
      rec[L.u0 + ig] = qv * dens;
    }
  } else {
    // vmr / extinction at this altitude (intpol_atm_1d_qk, :557-567); same level as for p and T
    const int pk = a.ray_pkg[r];
    const long long base = a.pkg_atm_off[pk] + a.ray_level0[r] + (long long)tail[LT_LEVEL];
    const double x0 = a.atm_z[base], x1 = a.atm_z[base + 1];
    const double w = (z - x0) / (x1 - x0);
    for (int iw = 0; iw < L.nw; iw++) {
      const double *__restrict__ k = a.atm_k + (size_t)iw * a.atm_stride + base;
      rec[4 + iw] = k[0] + w * (k[1] - k[0]);
    }
    for (int ig = 0; ig < L.ng; ig++) {
      const double *__restrict__ q = a.atm_q + (size_t)ig * a.atm_stride + base;
      const double qv = q[0] + w * (q[1] - q[0]);
      if (ig == a.ig_h2o) qh2o = qv;
      rec[L.u0 + ig] = qv * dens;
    }
  }
  rec[3] = qh2o;
  if (L.fast) {
    const TblDev &T = a.tbl;
    const int ncell = L.ng == 0 ? 0 : (L.cstride ? L.ng : 1); // one cell per gas, or one for all gases sharing the (p,T) grid
    for (int ic = 0; ic < ncell; ic++) {
      int ig = ic;
      if (!L.cstride) { ig = 0; while (ig < L.ng - 1 && T.gnp[ig] < 2) ++ig; } // first gas that has a table
      double *__restrict__ c = rec + L.c0 + L.cstride * ic;
      unsigned cell = kCellInvalid;
      double wp = 0, wt0 = 0, wt1 = 0;
      const int gnp = T.gnp[ig];
      if (gnp >= 2) {
        const double *__restrict__ gp = T.gp + (size_t)ig * T.npmax;
        const int ipr = bisect_asc([&](int i) { return gp[i]; }, gnp, p);
        const int nt0 = T.gnt[ig * T.npmax + ipr], nt1 = T.gnt[ig * T.npmax + ipr + 1];
        if (nt0 >= 2 && nt1 >= 2) {
          const double *__restrict__ g0 = T.gt + ((size_t)ig * T.npmax + ipr) * T.ntmax;
          const double *__restrict__ g1 = g0 + T.ntmax;
          const int it0 = bisect_asc([&](int i) { return g0[i]; }, nt0, t);
          const int it1 = bisect_asc([&](int i) { return g1[i]; }, nt1, t);
          cell = (unsigned)ipr | ((unsigned)it0 << 8) | ((unsigned)it1 << 16);
          wp = (p - gp[ipr]) / (gp[ipr + 1] - gp[ipr]);
          wt0 = (t - g0[it0]) / (g0[it0 + 1] - g0[it0]);
          wt1 = (t - g1[it1]) / (g1[it1 + 1] - g1[it1]);
        }
      }
      c[0] = wp; c[1] = wt0; c[2] = wt1;
      c[3] = __longlong_as_double((long long)cell);
    }
  }
}

cudaError_t launch_raytrace(const TraceArgs &a, cudaStream_t stream, int *launches) {
  if (launches) *launches = 0;
  if (a.prepare_atm && a.n_atm > 0) {
    atm_slopes_kernel<<<(unsigned)((a.n_atm + 255) / 256), 256, 0, stream>>>(a.atm_z, a.atm_p, a.atm_t, a.atm_lnp_slope,
                                                                            a.atm_lnp_slope + a.atm_stride, a.n_atm);
    if (launches) ++*launches;
  }
  if (a.prepare_atm && a.n_atm > 0 && a.ip != 1) {
    atm_geo_kernel<<<(unsigned)((a.n_atm + 255) / 256), 256, 0, stream>>>(a.atm_lon, a.atm_lat, a.atm_cart, a.atm_next, a.atm_stride, a.n_atm);
    if (launches) ++*launches;
  }
  if (a.n_rays <= 0) return cudaGetLastError();
  // pipelined chunks run beside the persistent EGA CTAs of the previous chunk, which leave ~4 K registers per SM:
  // one-warp ray CTAs (104 regs x 32) and two-warp finalisation CTAs (48 regs x 64) fit into that remainder
  const int bs = a.small_blocks ? 32 : 128, bf = a.small_blocks ? 64 : 256;
  // small batches: 8 lanes per ray (see ray_step_kernel); from ~16 k rays on the throughput form fills the schedulers
  const bool coop = !a.small_blocks && a.n_rays <= 16384 && !getenv("JRB_NO_COOP_TRACER");
  if (a.ip != 1) ray_geo_kernel<<<(unsigned)((a.n_rays + bs - 1) / bs), bs, 0, stream>>>(a);
  else if (coop) ray_step_kernel<8><<<(unsigned)((a.n_rays * 8 + 127) / 128), 128, 0, stream>>>(a);
  else ray_step_kernel<1><<<(unsigned)((a.n_rays + bs - 1) / bs), bs, 0, stream>>>(a);
  const long long n = a.n_rays * kNLOS;
  los_finalize_kernel<<<(unsigned)((n + bf - 1) / bf), bf, 0, stream>>>(a);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

} // namespace jrb
