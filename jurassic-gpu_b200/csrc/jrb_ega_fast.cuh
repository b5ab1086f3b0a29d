// jrb_ega_fast.cuh -- specialised EGA kernels, template <NGB gases held in registers, continuum MASK>.
//
// Mapping: one warp = one ray x 32 consecutive channels (lane = channel).  The LOS record of a segment is read as a
// warp-uniform broadcast; per-gas path transmittances tau_path[NGB] live in registers for the whole ray
// (the reference keeps tau_path[NG] in local memory, src/jr_fusion_kernel.mv4g.cu:6,12-14); rad/tau are
// accumulated in registers and stored once (the reference read-modify-writes obs_t in global memory every
// segment, src/jr_common.h:293-300).  The 16 continuum variants of src/jr_multiversion4gases.h are the MASK
// template parameter.  Warps fetch rays from a global work counter, so rays of different length (130..393
// segments) balance themselves.
//
// Search-free table access: every (gas, column-slot) keeps the bracket index it ended on in the previous segment
// as a 16-bit hint.  The next lookup loads that bracket first (one aligned 16-byte load) and only walks / bisects
// when the hint is off.  For monotone columns the index found is identical to the reference's full bisection
// (locate_tbl_id, src/jr_common.h:116-125), so results do not depend on the hints.
#pragma once
#include "jrb_ega_common.cuh"

namespace jrb {

namespace fast {

template <bool ON_EPS>
__device__ __forceinline__ float lo_of(const float4 b) { return ON_EPS ? b.y : b.x; }
template <bool ON_EPS>
__device__ __forceinline__ float hi_of(const float4 b) { return ON_EPS ? b.w : b.z; }

// Move (k, b) so that  val[k] <= x < val[k+1]  with k clipped to [0, nu-2]  (== reference bisection result).
template <bool ON_EPS>
__device__ __forceinline__ void relocate(const float4 *__restrict__ col, const int nu, const double x, int &k,
                                         float4 &b) {
  if (x < (double)lo_of<ON_EPS>(b)) {
    if (k == 0) return;
    int steps = 0;
    do { --k; b = col[k]; ++steps; } while (k > 0 && x < (double)lo_of<ON_EPS>(b) && steps < 2);
    if (k > 0 && x < (double)lo_of<ON_EPS>(b)) {
      int ilo = 0, ihi = k;
      while (ihi > ilo + 1) {
        const int i = (ihi + ilo) >> 1;
        const float v = ON_EPS ? col[i].y : col[i].x;
        if ((double)v > x) ihi = i; else ilo = i;
      }
      k = ilo; b = col[k];
    }
  } else if (x >= (double)hi_of<ON_EPS>(b)) {
    if (k >= nu - 2) return;
    int steps = 0;
    do { ++k; b = col[k]; ++steps; } while (k < nu - 2 && x >= (double)hi_of<ON_EPS>(b) && steps < 2);
    if (k < nu - 2 && x >= (double)hi_of<ON_EPS>(b)) {
      int ilo = k, ihi = nu - 1;
      while (ihi > ilo + 1) {
        const int i = (ihi + ilo) >> 1;
        const float v = ON_EPS ? col[i].y : col[i].x;
        if ((double)v > x) ihi = i; else ilo = i;
      }
      k = ilo; b = col[k];
    }
  }
}

// One table column of the EGA step: u* = u(eps) (get_u, may extrapolate), then eps(u* + u_seg) clamped to [0,1]
// (get_eps + c01, src/jr_common.h:249-257).  `hint` is updated to the bracket the lookup ended on.
__device__ __forceinline__ double column_step(const float4 *__restrict__ brk, const uint2 c, const double eps,
                                              const double useg, unsigned &hint) {
  const float4 *__restrict__ col = brk + c.x;
  const int nu = (int)c.y;
  int k = min((int)hint, nu - 2);
  float4 b = col[k];
  relocate<true>(col, nu, eps, k, b);
  const double ustar = lerp_fast((double)b.y, (double)b.x, (double)b.w, (double)b.z, eps);
  const double x = ustar + useg;
  relocate<false>(col, nu, x, k, b);
  hint = (unsigned)k;
  return clamp01(lerp_fast((double)b.x, (double)b.y, (double)b.z, (double)b.w, x));
}

} // namespace fast

template <int NGB, int MASK>
__global__ void __launch_bounds__(256, 2) ega_fast_kernel(const EgaArgs a) {
  const int lane = threadIdx.x & 31;
  const int ngroups = (a.nd + 31) >> 5;
  const unsigned long long n_items = (unsigned long long)a.n_rays * ngroups;
  const LosLayout L = a.los;
  const TblDev &T = a.tbl;
  const int nd = a.nd;

  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.work_counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const long long ir = (long long)(item / ngroups);
    const int grp = (int)(item - (unsigned long long)ir * ngroups);
    const int id_raw = grp * 32 + lane;
    const bool lane_on = id_raw < nd;
    const int id = lane_on ? id_raw : nd - 1;

    const double *__restrict__ rec = a.los_data + (size_t)ir * kNLOS * L.rec;
    const int np = a.ray_np[ir];
    const int win = a.window[id];

    unsigned valid = 0; // bit ig: this (gas, channel) pair has a table (np >= 2)
#pragma unroll
    for (int ig = 0; ig < NGB; ig++)
      if (ig < a.ng && T.np[ig * nd + id] >= 2) valid |= 1u << ig;

    double tau_path[NGB];
    unsigned h01[NGB], h23[NGB];
#pragma unroll
    for (int ig = 0; ig < NGB; ig++) { tau_path[ig] = 1.0; h01[ig] = 0; h23[ig] = 0; }
    double rad = 0.0, tau = 1.0;
    bool dead = false; // a gas went opaque (tau_path < 1e-9): nothing changes any more (src/jr_common.h:239,295)

    for (int ip = 0; ip < np; ++ip, rec += L.rec) {
      if (__all_sync(0xffffffffu, dead)) break;
      if (dead) continue;
      const double p = rec[0], t = rec[1], ds = rec[2];
      const double u_co2 = (MASK & 8) ? rec[L.u0 + a.ig_co2] : 0.0;
      const double u_h2o = (MASK & 4) ? rec[L.u0 + a.ig_h2o] : 0.0;
      const double beta_ds = continuum_beta_ds(MASK, a.chan, nd, id, p, t, ds, rec[4 + win], u_co2, u_h2o, rec[3]);

      double tau_gas = 1.0;
#pragma unroll
      for (int ig = 0; ig < NGB; ig++) {
        if (ig < a.ng) {
          double f;
          const double tp = tau_path[ig];
          if (tp < 1e-9) {
            f = 0.0;
          } else {
            f = 1.0;
            const double *__restrict__ cw = rec + L.c0 + 4 * ig;
            const unsigned cell = (unsigned)__double_as_longlong(cw[3]);
            if (((valid >> ig) & 1u) && cell != kCellInvalid) {
              const int ipr = cell & 0xff, it0 = (cell >> 8) & 0xff, it1 = (cell >> 16) & 0xff;
              const size_t g0 = (((size_t)ig * T.npmax + ipr) * T.ntmax + it0) * nd + id;
              const size_t g1 = (((size_t)ig * T.npmax + ipr + 1) * T.ntmax + it1) * nd + id;
              const uint2 c00 = T.col[g0], c01 = T.col[g0 + nd], c10 = T.col[g1], c11 = T.col[g1 + nd];
              if (c00.y >= 2 && c01.y >= 2 && c10.y >= 2 && c11.y >= 2) {
                const double eps = 1 - tp, useg = rec[L.u0 + ig];
                unsigned ha = h01[ig] & 0xffffu, hb = h01[ig] >> 16, hc = h23[ig] & 0xffffu, hd = h23[ig] >> 16;
                const double e00 = fast::column_step(T.brk, c00, eps, useg, ha);
                const double e01 = fast::column_step(T.brk, c01, eps, useg, hb);
                const double e10 = fast::column_step(T.brk, c10, eps, useg, hc);
                const double e11 = fast::column_step(T.brk, c11, eps, useg, hd);
                h01[ig] = ha | (hb << 16);
                h23[ig] = hc | (hd << 16);
                const double ep0 = clamp01(fma(cw[1], e01 - e00, e00));
                const double ep1 = clamp01(fma(cw[2], e11 - e10, e10));
                const double ept = clamp01(fma(cw[0], ep1 - ep0, ep0));
                f = (1. - ept) * fast_rcp(tp);
              }
            }
          }
          tau_path[ig] = tp * f;
          tau_gas *= f;
        }
      }
      if (tau_gas == 0.0) {
        // some gas is opaque: its factor stays 0 for the rest of the ray, so tau_gas stays 0 and no
        // further segment can change rad or tau (accumulate() requires tau_gas > 1e-50)
        bool any_opaque = false;
#pragma unroll
        for (int ig = 0; ig < NGB; ig++) any_opaque |= (ig < a.ng) && (tau_path[ig] < 1e-9);
        dead = any_opaque;
      }
      const double src = planck_source(T.sr, nd, id, t);
      accumulate(rad, tau, beta_ds, src, tau_gas);
    }
    epilogue(rad, tau, a.ray_tsurf[ir], T.sr, nd, id, a.write_bbt, a.chan[CH_NU * nd + id]);
    if (lane_on) {
      a.rad[(size_t)ir * nd + id] = rad;
      a.tau[(size_t)ir * nd + id] = tau;
    }
  }
}

template <int NGB, int MASK>
cudaError_t launch_ega_fast_t(const EgaArgs &a, cudaStream_t stream, int sm_count) {
  int blocks_per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, ega_fast_kernel<NGB, MASK>, 256, 0);
  if (e != cudaSuccess) return e;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  const int ngroups = (a.nd + 31) >> 5;
  const long long n_items = a.n_rays * ngroups;
  long long grid = (long long)sm_count * blocks_per_sm; // persistent: a whole number of CTAs per SM
  const long long need = (n_items + 7) / 8;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  ega_fast_kernel<NGB, MASK><<<(unsigned)grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// one translation unit per MASK instantiates NGB = 1..8
template <int MASK>
cudaError_t launch_ega_fast_mask(const EgaArgs &a, cudaStream_t stream, int sm_count, int *ngb_out);

} // namespace jrb
