// jrb_ega_generic.cu -- reference-semantics EGA kernel: any number of gases, per-channel (p,T) axes, plain
// bisections exactly as locate_id/locate_tbl_id (src/jr_common.h:106-125).  One thread per (ray, channel).
// It is the fallback for table sets the specialised kernels do not accept (channel-dependent axes, columns longer
// than 1023 entries, ng > 32) and doubles as an on-device cross-check of the specialised kernels.
#include "jrb_ega_common.cuh"
#include <jurassic_b200.h>

namespace jrb {

namespace {

struct Column { unsigned first, nu; };

__device__ __forceinline__ Column load_col(const TblDev &T, int ig, int ip, int it, int id) {
  const uint2 c = T.col[(((size_t)ig * T.npmax + ip) * T.ntmax + it) * T.nd + id];
  return Column{c.x, c.y & ~kColNonMonotone};
}

// get_u (src/jr_common.h:179-185): column density giving emissivity `eps` in this column (may extrapolate)
__device__ __forceinline__ double column_u_of_eps(const float4 *__restrict__ brk, const Column c, const double eps) {
  const float4 *__restrict__ b0 = brk + c.first;
  const int idx = bisect_asc([&](int i) { return (double)b0[i].y; }, (int)c.nu, eps);
  const float4 b = b0[idx];
  return lerp_div((double)b.y, (double)b.x, (double)b.w, (double)b.z, eps);
}
// get_eps (:156-177)
__device__ __forceinline__ double column_eps_of_u(const float4 *__restrict__ brk, const Column c, const double u) {
  const float4 *__restrict__ b0 = brk + c.first;
  const int idx = bisect_asc([&](int i) { return (double)b0[i].x; }, (int)c.nu, u);
  const float4 b = b0[idx];
  return lerp_div((double)b.x, (double)b.y, (double)b.z, (double)b.w, u);
}

// ega_eps (src/jr_common.h:237-268)
__device__ double ega_factor_generic(const TblDev &T, const double tau, const double t, const double u,
                                     const double p, const int ig, const int id) {
  if (tau < 1e-9) return 0.;
  const int np = T.np[ig * T.nd + id];
  if (np < 2) return 1.;
  const size_t gbase = (size_t)ig * T.npmax;
  const double *__restrict__ pax = T.pax + gbase * T.nd + id;
  const int ipr = bisect_asc([&](int i) { return pax[(size_t)i * T.nd]; }, np, p);
  const int nt0 = T.nt[(gbase + ipr) * T.nd + id], nt1 = T.nt[(gbase + ipr + 1) * T.nd + id];
  if (nt0 < 2 || nt1 < 2) return 1.;
  const double *__restrict__ t0ax = T.tax + (gbase + ipr) * T.ntmax * T.nd + id;
  const double *__restrict__ t1ax = t0ax + (size_t)T.ntmax * T.nd;
  const int it0 = bisect_asc([&](int i) { return t0ax[(size_t)i * T.nd]; }, nt0, t);
  const Column c00 = load_col(T, ig, ipr, it0, id), c01 = load_col(T, ig, ipr, it0 + 1, id);
  if (c00.nu < 2 || c01.nu < 2) return 1.;
  const int it1 = bisect_asc([&](int i) { return t1ax[(size_t)i * T.nd]; }, nt1, t);
  const Column c10 = load_col(T, ig, ipr + 1, it1, id), c11 = load_col(T, ig, ipr + 1, it1 + 1, id);
  if (c10.nu < 2 || c11.nu < 2) return 1.;

  const double eps = 1 - tau;
  const double u00 = column_u_of_eps(T.brk, c00, eps), u01 = column_u_of_eps(T.brk, c01, eps);
  const double u10 = column_u_of_eps(T.brk, c10, eps), u11 = column_u_of_eps(T.brk, c11, eps);
  const double e00 = clamp01(column_eps_of_u(T.brk, c00, u00 + u)), e01 = clamp01(column_eps_of_u(T.brk, c01, u01 + u));
  const double e10 = clamp01(column_eps_of_u(T.brk, c10, u10 + u)), e11 = clamp01(column_eps_of_u(T.brk, c11, u11 + u));
  const double ep0 = clamp01(lerp_div(t0ax[(size_t)it0 * T.nd], e00, t0ax[(size_t)(it0 + 1) * T.nd], e01, t));
  const double ep1 = clamp01(lerp_div(t1ax[(size_t)it1 * T.nd], e10, t1ax[(size_t)(it1 + 1) * T.nd], e11, t));
  const double ept = clamp01(lerp_div(pax[(size_t)ipr * T.nd], ep0, pax[(size_t)(ipr + 1) * T.nd], ep1, p));
  return (1. - ept) / tau;
}

__global__ void __launch_bounds__(128) ega_generic_kernel(EgaArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_rays * a.nd) return;
  const long long ir = idx / a.nd;
  const int id = (int)(idx - ir * a.nd);
  const LosLayout L = a.los;
  const double *__restrict__ rec = a.los_data + (size_t)ir * kNLOS * L.rec;
  const int np = a.ray_np[ir];
  const int win = a.window[id];

  double tau_path[JRB_MAX_NG];
  for (int ig = 0; ig < a.ng; ig++) tau_path[ig] = 1.0;
  double rad = 0.0, tau = 1.0;

  for (int ip = 0; ip < np; ++ip, rec += L.rec) {
    const double p = rec[0], t = rec[1], ds = rec[2];
    const double u_co2 = (a.ctm_mask & 8) ? rec[L.u0 + a.ig_co2] : 0.0;
    const double u_h2o = (a.ctm_mask & 4) ? rec[L.u0 + a.ig_h2o] : 0.0;
    const double beta_ds = continuum_beta_ds(a.ctm_mask, a.chan, a.nd, id, p, t, ds, a.nw > 0 ? rec[4 + win] : 0.0, u_co2, u_h2o, rec[3]);
    double tau_gas = 1.0;
    for (int ig = 0; ig < a.ng; ig++) { // apply_ega_core (src/jr_common.h:270-280)
      const double f = ega_factor_generic(a.tbl, tau_path[ig], t, rec[L.u0 + ig], p, ig, id);
      tau_path[ig] *= f;
      tau_gas *= f;
    }
    const double src = planck_source(a.tbl.sr, a.nd, id, t);
    accumulate(rad, tau, beta_ds, src, tau_gas);
  }
  epilogue(rad, tau, a.ray_tsurf[ir], a.tbl.sr, a.nd, id, a.write_bbt, a.chan[CH_NU * a.nd + id]);
  a.rad[idx] = rad;
  a.tau[idx] = tau;
  if (a.rad_host) {
    a.rad_host[ir][id] = rad;
    a.tau_host[ir][id] = tau;
  }
}

} // namespace

cudaError_t launch_ega_generic(const EgaArgs &a, cudaStream_t stream) {
  const long long n = a.n_rays * a.nd;
  if (n <= 0) return cudaSuccess;
  const int block = 128;
  ega_generic_kernel<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(a);
  return cudaGetLastError();
}

} // namespace jrb
