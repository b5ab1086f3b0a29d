// jrb_raytrace.cu -- line-of-sight ray tracer + column densities (one thread per ray), sm_100a.
//
// Computes what the reference's traceray() computes (src/jr_common.h:585-711: profile selection :127-154,
// altitude range :411-420, observer/view-point rejection :598-599, entry search :610-621, stepping loop
// :624-691 with refraction :664-681, tangent point :502-539, trapezoid rule :437-443, column density
// :446-453) but is organised for the device path:
//   pass 1 walks the ray and writes raw LOS records (p, T, raw ds, q, k, z/lon/lat) to the ray's record array;
//   pass 2 (same thread) finalises segment lengths, converts vmr to column densities and -- for tables whose
//   (p,T) axes do not depend on the channel -- resolves the table cell and interpolation weights per gas once
//   per segment, so that the EGA kernel does not search axes per channel.
#include "jrb_internal.h"

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
__device__ __forceinline__ void cart_to_geo(const double x[3], double *alt, double *lon, double *lat) {
  const double r2d = 180.0 / M_PI;
  const double radius = norm3(x);
  *lat = asin(x[2] / radius) * r2d;
  *lon = atan2(x[1], x[0]) * r2d;
  *alt = radius - kRE;
}

// `locate` of the reference (src/jr_common.h:87-104): ascending or descending axis
__device__ __forceinline__ int locate_z(const double *__restrict__ zz, int n, double x) {
  int ilo = 0, ihi = n - 1, i = (n - 1) >> 1;
  if (zz[i] < zz[i + 1]) {
    while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (zz[i] > x) ihi = i; else ilo = i; }
  } else {
    while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (zz[i] <= x) ihi = i; else ilo = i; }
  }
  return ilo;
}

// pressure (exponential) and temperature (linear) at altitude z0 (intpol_atm_1d_pt, :549-555; eip :53-57)
__device__ __forceinline__ void interp_pt(const double *__restrict__ az, const double *__restrict__ ap,
                                          const double *__restrict__ at, int n, double z0, double *p, double *t,
                                          int *idx_out) {
  const int ip = locate_z(az, n, z0);
  const double x0 = az[ip], x1 = az[ip + 1];
  const double y0 = ap[ip], y1 = ap[ip + 1];
  if (y0 > 0 && y1 > 0) *p = y0 * exp(log(y1 / y0) / (x1 - x0) * (z0 - x0));
  else *p = lerp_div(x0, y0, x1, y1, z0);
  *t = lerp_div(x0, at[ip], x1, at[ip + 1], z0);
  *idx_out = ip;
}

__device__ __forceinline__ double refractivity(double p, double t) { return 7.753e-05 * p / t; }

} // namespace

__global__ void __launch_bounds__(128) raytrace_kernel(TraceArgs a) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_rays) return;

  const LosLayout L = a.los;
  double *__restrict__ rec0 = a.los_data + (size_t)r * kNLOS * L.rec;

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
  const int n = upper - lower;
  const double *__restrict__ az = a.atm_z + abase + lower;
  const double *__restrict__ ap = a.atm_p + abase + lower;
  const double *__restrict__ at = a.atm_t + abase + lower;
  const double *__restrict__ alon = a.atm_lon + abase + lower;
  const double *__restrict__ alat = a.atm_lat + abase + lower;

  // ---- altitude range of the profile (altitude_range_nn, :411-420) ----
  double zmin = az[0], zmax = az[0];
  for (int i = 0; i < n && alon[i] == alon[0] && alat[i] == alat[0]; ++i) {
    zmax = fmax(zmax, az[i]);
    zmin = fmin(zmin, az[i]);
  }

  const bool rejected = (obsz < zmin) || (vpz > zmax - 0.001);
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

    double z_low = 1e99, lon, lat, p, t;
    double pz = 0, plon = 0, plat = 0; // previous point
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
      if ((z < zmin) || (z > zmax)) { // left the atmosphere: clip the last segment (:637-648)
        if (np == 0) break;           // (reference would read los[-1]; cannot happen after the entry search)
        double xh[3];
        stop = (z < zmin) ? 2 : 1;
        geo_to_cart(pz, plon, plat, xh);
        const double zfrac = (z < zmin) ? zmin : zmax;
        const double frac = (zfrac - pz) / (z - pz);
        for (int i = 0; i < 3; i++) x[i] = xh[i] + frac * (x[i] - xh[i]);
        cart_to_geo(x, &z, &lon, &lat);
        rec0[(size_t)(np - 1) * L.rec + 2] = ds * frac;
        ds = 0.0;
      }
      int ia;
      interp_pt(az, ap, at, n, z, &p, &t, &ia);
      double *__restrict__ rec = rec0 + (size_t)np * L.rec;
      rec[0] = p; rec[1] = t; rec[2] = ds;
      {
        const double x0 = az[ia], x1 = az[ia + 1];
        for (int ig = 0; ig < L.ng; ig++) { // vmr goes to the u slot for now (intpol_atm_1d_qk, :557-567)
          const double *__restrict__ q = a.atm_q + (size_t)ig * a.atm_stride + abase + lower;
          rec[L.u0 + ig] = lerp_div(x0, q[ia], x1, q[ia + 1], z);
        }
        for (int iw = 0; iw < L.nw; iw++) {
          const double *__restrict__ k = a.atm_k + (size_t)iw * a.atm_stride + abase + lower;
          rec[4 + iw] = lerp_div(x0, k[ia], x1, k[ia + 1], z);
        }
      }
      rec[L.z0 + 0] = z; rec[L.z0 + 1] = lon; rec[L.z0 + 2] = lat;
      pz = z; plon = lon; plat = lat;
      if (z < z_low) { z_low = z; z_low_idx = np; }

      if (stop) { tsurf = (stop == 2 ? t : -999.0); break; }

      double nref = 1.0, ngr[3] = {0.0, 0.0, 0.0};
      if (a.refrac && z <= 60.0) { // refractivity gradient by finite differences at the half step (:664-681)
        nref += refractivity(p, t);
        double xh[3], ph, th; int dummy;
        for (int i = 0; i < 3; i++) xh[i] = x[i] + 0.5 * ds * ex0[i];
        interp_pt(az, ap, at, n, norm3(xh) - kRE, &ph, &th, &dummy);
        const double n2 = refractivity(ph, th);
        for (int i = 0; i < 3; i++) {
          const double h = 0.02;
          xh[i] += h;
          interp_pt(az, ap, at, n, norm3(xh) - kRE, &ph, &th, &dummy);
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
    if (np > 0 || stop) ++np; // the reference increments after the loop (:692)
    if (np > kNLOS) np = kNLOS; // (reference: fatal "Too many LOS points" on CPU when np >= NLOS)
  }

  // ---- tangent point (before changing segment lengths; :502-539) ----
  if (np > 0) {
    const int ip = z_low_idx;
    if (ip <= 0 || ip >= np - 1) {
      const double *rl = rec0 + (size_t)(np - 1) * L.rec + L.z0;
      tpz = rl[0]; tplon = rl[1]; tplat = rl[2];
    } else {
      const double *r0 = rec0 + (size_t)(ip - 1) * L.rec, *r1 = rec0 + (size_t)ip * L.rec,
                   *r2 = rec0 + (size_t)(ip + 1) * L.rec;
      const double yy0 = r0[L.z0], yy1 = r1[L.z0], yy2 = r2[L.z0];
      const double ds0 = r1[2], ds1 = r2[2];
      const double dyy10 = yy1 - yy0, dyy21 = yy2 - yy1;
      const double x1 = sqrt(ds0 * ds0 - dyy10 * dyy10);
      const double x2 = x1 + sqrt(ds1 * ds1 - dyy21 * dyy21);
      const double dx12 = x1 - x2;
      const double qa = (dyy10 * x2 + (yy0 - yy2) * x1) / (x1 * x2 * dx12);
      const double qb = dyy10 / x1 - qa * x1;
      const double qc = yy0;
      const double xt = -qb / (2 * qa);
      tpz = (qa * xt + qb) * xt + qc;
      double v[3], v0[3], v2[3], dummy;
      geo_to_cart(r0[L.z0], r0[L.z0 + 1], r0[L.z0 + 2], v0);
      geo_to_cart(r2[L.z0], r2[L.z0 + 1], r2[L.z0 + 2], v2);
      for (int i = 0; i < 3; i++) v[i] = lerp_div(0.0, v0[i], x2, v2[i], xt);
      cart_to_geo(v, &dummy, &tplon, &tplat);
    }
  }

  // ---- pass 2: trapezoid rule, column densities, table cells ----
  {
    double ds_prev = 0.0;
    for (int ip = 0; ip < np; ip++) {
      double *__restrict__ rec = rec0 + (size_t)ip * L.rec;
      const double ds_raw = rec[2];
      const double ds = (ip == 0) ? 0.5 * ds_raw : 0.5 * (ds_prev + ds_raw); // (:437-443)
      ds_prev = ds_raw;
      rec[2] = ds;
      const double p = rec[0], t = rec[1];
      rec[3] = (a.ig_h2o >= 0) ? rec[L.u0 + a.ig_h2o] : 0.0;
      for (int ig = 0; ig < L.ng; ig++) {
        const double q = rec[L.u0 + ig];
        rec[L.u0 + ig] = 10. * q * p / (kBoltzmann * t) * ds; // (:446-453)
      }
      if (L.fast) {
        const TblDev &T = a.tbl;
        for (int ig = 0; ig < L.ng; ig++) {
          double *__restrict__ c = rec + L.c0 + 4 * ig;
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
  }

  a.ray_np[r] = np;
  a.ray_tsurf[r] = tsurf;
  a.tp[0 * a.geo_stride + r] = tpz;
  a.tp[1 * a.geo_stride + r] = tplon;
  a.tp[2 * a.geo_stride + r] = tplat;
}

cudaError_t launch_raytrace(const TraceArgs &a, cudaStream_t stream) {
  if (a.n_rays <= 0) return cudaSuccess;
  const int block = 128;
  const long long grid = (a.n_rays + block - 1) / block;
  raytrace_kernel<<<(unsigned)grid, block, 0, stream>>>(a);
  return cudaGetLastError();
}

} // namespace jrb
