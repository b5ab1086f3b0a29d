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

namespace jrb {

namespace {

// thread per (ray, channel): tau_gas = product of the block products in gas order, then what the fused kernel does per
// segment (continua_core_bbbb, src_planck_core, new_obs_core) and per ray (add_surface_core, brightness_core)
__global__ void __launch_bounds__(128) ega_combine_kernel(const EgaArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_rays * a.nd) return;
  const long long ir = idx / a.nd;
  const int id = (int)(idx - ir * a.nd), nd = a.nd;
  const LosLayout L = a.los;
  const double *__restrict__ rec = a.los_data + (size_t)ir * kNLOS * L.rec;
  const int np = a.ray_np[ir];
  const int win = a.window[id];
  const int nb = a.n_gas_blocks;
  // segments every block has a product for; beyond it some block's factor is 0 and nothing is accumulated any more
  int n_live = np;
  for (int b = 0; b < nb; b++) n_live = min(n_live, a.partial_len[((size_t)b * a.n_rays + ir) * nd + id]);
  const size_t bstride = (size_t)a.n_rays * kNLOS * nd;
  const double *__restrict__ part = a.partial + ((size_t)ir * kNLOS) * nd + id;
  double rad = 0.0, tau = 1.0;
  for (int ip = 0; ip < n_live; ++ip, rec += L.rec) {
    double tau_gas = 1.0;
    for (int b = 0; b < nb; b++) tau_gas *= part[(size_t)b * bstride + (size_t)ip * nd];
    const double p = rec[0], t = rec[1], ds = rec[2];
    const double u_co2 = (a.ctm_mask & 8) ? rec[L.u0 + a.ig_co2] : 0.0;
    const double u_h2o = (a.ctm_mask & 4) ? rec[L.u0 + a.ig_h2o] : 0.0;
    const double beta_ds = continuum_beta_ds(a.ctm_mask, a.chan, nd, id, p, t, ds, a.nw > 0 ? rec[4 + win] : 0.0, u_co2, u_h2o, rec[3]);
    const double src = planck_source(a.tbl.sr, nd, id, t);
    accumulate(rad, tau, beta_ds, src, tau_gas);
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

cudaError_t launch_ega_combine(const EgaArgs &a, cudaStream_t stream) {
  const long long n = a.n_rays * a.nd;
  if (n <= 0) return cudaSuccess;
  ega_combine_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(a);
  return cudaGetLastError();
}

} // namespace jrb
