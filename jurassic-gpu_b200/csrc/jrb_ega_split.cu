// jrb_ega_split.cu -- split mode of the EGA step: gas-block passes (ega_fast_kernel<.., SPLIT = true>, instantiated here once,
// the pass does not depend on the continuum mask) and the combine kernel.
//
// Why split: the gases of a ray do not depend on each other (tau_path[ig] is a per-gas recurrence,
// src/jr_common.h:270-280); only their product enters the radiance update.  Cutting the gas loop into blocks
//   * gives a single 1088-ray package (what every unmodified formod() caller hands over) ng times more independent warps,
//     so the sequential chain a warp walks is one gas long instead of ng;
//   * keeps the per-thread state (16 B per gas) and the hot table set (channels per warp x gases per block) bounded for
//     many-gas set-ups (30-gas refspec shape): 24 warps per SM instead of 8, tables of one pass fit the L2.
// The price is the scratch traffic of the block products: 8 B per (block, segment, channel) written and read once, against
// 176 B of table gathers per (gas, segment, channel).
#include "jrb_ega_fast.cuh"
#include "jrb_ega_tiled.cuh"
#include <cstdlib>

namespace jrb {

namespace {

// Everything of a segment except the along-ray recurrence, fully parallel: thread per (ray, segment, channel).
//   tau_gas = product of the block products in the canonical order (groups in order, inside a group in gas order; a block
//             whose gas went opaque at an earlier segment contributes 0)
//   continua / extinction (continua_core_bbbb), band Planck source (src_planck_core), and the two terms of new_obs_core:
//   a.seg_pre[ray][segment][channel] = {src * eps, 1 - eps}  with eps = 1 - tau_gas exp(-beta_ds), or {0, 1} where the
//   reference skips the segment (tau_gas <= 1e-50, src/jr_common.h:295).
// This takes all loads of block products and all transcendentals out of the sequential loop of the combine kernel.
__global__ void __launch_bounds__(256) ega_segment_kernel(const EgaArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nd = a.nd;
  const long long seg = idx / nd; // (ray, segment) pair
  const int id = (int)(idx - seg * nd);
  const long long ir = seg / kNLOS;
  const int ip = (int)(seg - ir * kNLOS);
  if (ir >= a.n_rays || ip >= a.ray_np[ir]) return;
  const int nb = a.n_gas_blocks, bpg = max(a.blocks_per_group, 1);
  const size_t bstride = (size_t)a.n_rays * kNLOS * nd, lstride = (size_t)a.n_rays * nd;
  const double *__restrict__ part = a.partial + idx;
  const int *__restrict__ plen = a.partial_len + (size_t)ir * nd + id;
  double tau_gas = 1.0;
  for (int b0 = 0; b0 < nb; b0 += bpg) {
    double pg = 1.0;
    for (int b = b0; b < min(b0 + bpg, nb); b++) pg *= (ip < plen[(size_t)b * lstride]) ? part[(size_t)b * bstride] : 0.0;
    tau_gas *= pg;
  }
  double2 out = make_double2(0.0, 1.0);
  if (tau_gas > 1e-50) { // new_obs_core (src/jr_common.h:293-300)
    const LosLayout L = a.los;
    const double *__restrict__ rec = a.los_data + ((size_t)ir * kNLOS + ip) * L.rec;
    const double p = rec[0], t = rec[1], ds = rec[2];
    const double u_co2 = (a.ctm_mask & 8) ? rec[L.u0 + a.ig_co2] : 0.0;
    const double u_h2o = (a.ctm_mask & 4) ? rec[L.u0 + a.ig_h2o] : 0.0;
    const double beta_ds = continuum_beta_ds(a.ctm_mask, a.chan, nd, id, p, t, ds, a.nw > 0 ? rec[4 + a.window[id]] : 0.0, u_co2, u_h2o, rec[3]);
    const double src = planck_source(a.tbl.sr, nd, id, t);
    const double eps = 1. - tau_gas * exp(-beta_ds);
    out = make_double2(src * eps, 1. - eps);
  }
  a.seg_pre[idx] = out;
}

// thread per (ray, channel): the recurrence rad += (src eps) tau; tau *= (1 - eps) over the segments (same operations in the
// same order as the fused kernel), then the per-ray epilogues (add_surface_core, brightness_core).  One 16-byte load per
// segment that does not depend on the recurrence: unrolled, eight segments in flight.
__global__ void __launch_bounds__(128) ega_combine_kernel(const EgaArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_rays * a.nd) return;
  const long long ir = idx / a.nd;
  const int id = (int)(idx - ir * a.nd), nd = a.nd;
  const int np = a.ray_np[ir];
  const double2 *__restrict__ pre = a.seg_pre + ((size_t)ir * kNLOS) * nd + id;
  double rad = 0.0, tau = 1.0;
  int ip = 0;
  for (; ip + 8 <= np; ip += 8) {
    double2 e[8];
#pragma unroll
    for (int j = 0; j < 8; j++) e[j] = pre[(size_t)(ip + j) * nd];
#pragma unroll
    for (int j = 0; j < 8; j++) { rad += e[j].x * tau; tau *= e[j].y; }
  }
  for (; ip < np; ++ip) {
    const double2 e = pre[(size_t)ip * nd];
    rad += e.x * tau;
    tau *= e.y;
  }
  epilogue(rad, tau, a.ray_tsurf[ir], a.tbl.sr, nd, id, a.write_bbt, a.chan[CH_NU * nd + id]);
  a.rad[idx] = rad;
  a.tau[idx] = tau;
  if (a.rad_host) {
    a.rad_host[ir][id] = rad;
    a.tau_host[ir][id] = tau;
  }
}

} // namespace

cudaError_t launch_ega_split(const EgaArgs &a, cudaStream_t stream, int sm_count) {
  const bool multi = a.cpw < 32;
  if (a.use_tiled && !multi && !a.per_channel_axes) // segment-tiled pass: brackets stay in registers across the segments of a tile
    return a.unsorted_columns ? launch_ega_tiled_tm<0, true, true>(a, stream, sm_count) : launch_ega_tiled_tm<0, false, true>(a, stream, sm_count);
  if (a.per_channel_axes)
    return multi ? launch_ega_fast_tm<0, true, true, true, true>(a, stream, sm_count) : launch_ega_fast_tm<0, false, true, true, true>(a, stream, sm_count);
  if (a.unsorted_columns)
    return multi ? launch_ega_fast_tm<0, true, true, true>(a, stream, sm_count) : launch_ega_fast_tm<0, false, true, true>(a, stream, sm_count);
  return multi ? launch_ega_fast_tm<0, true, false, true>(a, stream, sm_count) : launch_ega_fast_tm<0, false, false, true>(a, stream, sm_count);
}

cudaError_t launch_ega_split_passes(const EgaArgs &a, cudaStream_t stream) {
  if (a.n_gas_blocks < 1 || a.gases_per_block < 1 || !a.partial || !a.partial_len) return cudaErrorInvalidValue;
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  return launch_ega_split(a, stream, sm);
}

cudaError_t launch_ega_segments(const EgaArgs &a, cudaStream_t stream) {
  const long long n = a.n_rays * kNLOS * a.nd;
  if (n <= 0) return cudaSuccess;
  if (!a.seg_pre) return cudaErrorInvalidValue;
  ega_segment_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_ega_combine(const EgaArgs &a, cudaStream_t stream) {
  const long long n = a.n_rays * a.nd;
  if (n <= 0) return cudaSuccess;
  ega_combine_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(a);
  return cudaGetLastError();
}

} // namespace jrb
