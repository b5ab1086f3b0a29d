// jrb_ega_tiled.cuh -- segment-tiled form of the specialised EGA kernel (one ray x 32 channels per warp).
//
// ega_fast_kernel walks a ray segment by segment and, inside a segment, gas by gas.  Every (segment, gas) step starts
// cold: per-gas state from shared memory, four column descriptors, four hinted brackets from L2/L1.  But the gases of a ray
// do not depend on each other, and consecutive segments of a ray mostly stay in the same table cell and the same brackets
// (measured on the Config-D package with the CPU restatement: same (p,T) cell in 80 % of the steps, same bracket in the
// same column in 60 %).  The loop nest is therefore turned inside out over a TILE of T consecutive segments:
//
//     for tile:   (T line-of-sight records arrive by ONE TMA bulk copy; the next tile is in flight meanwhile)
//       for gas:     per-gas state, column descriptors and the four current brackets live in REGISTERS across ...
//         for segment of the tile:   ... the segments: descriptors are reloaded only when the ray enters another table
//                                    cell, a bracket only when the search moves off it
//            prod[segment] *= factor                       (shared memory, [T][thread])
//       for segment of the tile:  continuum, Planck source, radiance update with prod[segment]
//
// The product of the gas factors of a segment is formed in gas order, the segments are accumulated in order: results are
// bit-identical to ega_fast_kernel.  Used whenever a warp handles one ray x 32 channels and the (p,T) axes are shared by the
// channels; also as the gas-block pass of the split mode (SPLIT).
#pragma once
#include "jrb_ega_fast.cuh"

namespace jrb {

#ifndef JRB_EGA_TILE
#define JRB_EGA_TILE 6
#endif
constexpr int kEgaTile = JRB_EGA_TILE;

__host__ __device__ inline size_t ega_tiled_smem_bytes(int ng, int rec, int threads) {
  const int nwarps = threads / 32;
  return 16 + (size_t)nwarps * 2 * 8                 // work-chunk state, mbarriers
         + (size_t)nwarps * 2 * kEgaTile * rec * 8   // record tiles, double buffered per warp (whole records: one copy per tile)
         + (size_t)ng * threads * 16                 // tau_path + hints
         + (size_t)kEgaTile * threads * 8;           // per-segment products of the gas factors
}

// SPLIT = true: gas-block pass of the split mode (see ega_fast_kernel): an item is (gas block, channel group, ray), the
// products of the block's factors go to a.partial instead of into the radiance update (MASK plays no role then).
template <int MASK, bool ROBUST, bool SPLIT = false>
__global__ void __launch_bounds__(kEgaBlock, JRB_EGA_MINBLOCKS) ega_tiled_kernel(const EgaArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int T = kEgaTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const LosLayout L = a.los;
  const TblDev &Tb = a.tbl;
  const int nd = a.nd, ng = a.ng;
  const int ngs = SPLIT ? a.gases_per_block : ng; // gases whose state a thread holds
  const int sstride = blockDim.x;

  unsigned long long *chunk_state = reinterpret_cast<unsigned long long *>(smem_raw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + 16) + warp * 2;
  double *recbuf = reinterpret_cast<double *>(smem_raw + 16 + (size_t)nwarps * 16) + (size_t)warp * 2 * T * L.rec;
  double *tau_s = reinterpret_cast<double *>(smem_raw + 16 + (size_t)nwarps * 16 + (size_t)nwarps * 2 * T * L.rec * 8) + tid;
  unsigned long long *hint_s = reinterpret_cast<unsigned long long *>(tau_s - tid + (size_t)ngs * sstride) + tid;
  double *prod_s = reinterpret_cast<double *>(hint_s - tid + (size_t)ngs * sstride) + tid;

  if (lane == 0) { fast::mbar_init(&bars[0], 1); fast::mbar_init(&bars[1], 1); }
  if (tid == 0) *chunk_state = fast::kChunkEmpty;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned parity0 = 0, parity1 = 0;
  // lock step (see ega_fast_kernel): the warps of a CTA start their rays together when the rays of a chunk are equally long
  const bool phase_lock = a.phase_lock_mode == 1 || (a.phase_lock_mode < 0 && a.balance != nullptr && a.balance[0] * 32ull < a.balance[1]);

  const int ngroups = (nd + 31) / 32;
  const unsigned long long n_items_blk = (unsigned long long)a.n_rays * ngroups; // channel-group major
  const unsigned long long n_items = SPLIT ? n_items_blk * (unsigned long long)a.n_gas_blocks : n_items_blk; // gas-block major

  for (;;) {
    unsigned long long item = 0;
    if (phase_lock) {
      __syncthreads();
      if (tid == 0) *chunk_state = atomicAdd(a.work_counter, (unsigned long long)nwarps);
      __syncthreads();
      const unsigned long long base = *chunk_state;
      if (base >= n_items) break;
      item = base + warp;
      if (item >= n_items) continue;
    } else {
      if (lane == 0) item = fast::next_item(chunk_state, a.work_counter, (unsigned)a.work_chunk);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= n_items) break;
    }
    int g0 = 0, g1 = ng, gblk = 0, grp; // gases and channel group of this item
    long long ir;
    if (SPLIT && a.tail_n > 0) {
      // latency mode: ALL rays longest first (tail_perm covers the whole launch), the gas blocks and channel groups of a ray
      // next to each other -- the items that start last are the shortest ones, which is what sets the duration of a launch
      // that is only one to two rounds of warps long
      const unsigned long long per_ray = (unsigned long long)a.n_gas_blocks * ngroups;
      const unsigned long long rr = item / per_ray;
      const unsigned rest = (unsigned)(item - rr * per_ray);
      gblk = (int)(rest % (unsigned)a.n_gas_blocks);
      grp = (int)(rest / (unsigned)a.n_gas_blocks);
      ir = a.tail_perm[rr];
    } else {
      if (SPLIT) {
        gblk = (int)(item / n_items_blk);
        item -= (unsigned long long)gblk * n_items_blk;
      }
      grp = (int)(item / (unsigned long long)a.n_rays);
      ir = (long long)(item - (unsigned long long)grp * (unsigned long long)a.n_rays);
      if (!SPLIT && a.tail_n > 0 && item >= n_items - (unsigned long long)a.tail_n) // the last items: longest ray first
        ir = a.n_rays - a.tail_n + a.tail_perm[item - (n_items - (unsigned long long)a.tail_n)];
    }
    if (SPLIT) {
      g0 = gblk * a.gases_per_block;
      g1 = min(ng, g0 + a.gases_per_block);
    }
    const int id_raw = grp * 32 + lane;
    const bool lane_on = id_raw < nd;
    const int id = lane_on ? id_raw : nd - 1;

    const double *__restrict__ rec_g = a.los_data + (size_t)ir * kNLOS * L.rec;
    const int np = a.ray_np[ir];
    const int win = a.window[id];
    for (int ig = g0; ig < g1; ig++) {
      tau_s[(ig - g0) * sstride] = 1.0;
      hint_s[(ig - g0) * sstride] = (Tb.np[ig * nd + id] >= 2) ? 0ull : ~0ull; // no table: factor 1 (src/jr_common.h:240)
    }
    double rad = 0.0, tau = 1.0;
    bool dead = false; // a gas went opaque: the remaining segments change nothing (src/jr_common.h:239,295)
    int n_done = 0;    // SPLIT: segments for which this lane stored a block product
    double *__restrict__ part = nullptr;
    if (SPLIT) part = a.partial + (((size_t)gblk * (size_t)a.n_rays + (size_t)ir) * kNLOS) * (size_t)nd + id;
    const int ntiles = (np + T - 1) / T;

    __syncwarp();
    if (lane == 0 && np > 0) {
      const unsigned bytes = (unsigned)(min(T, np) * L.rec) * 8u;
      fast::mbar_expect_tx(&bars[0], bytes);
      fast::tma_load_1d(recbuf, rec_g, bytes, &bars[0]);
    }
    for (int tile = 0; tile < ntiles; ++tile) {
      const int b = tile & 1, ip0 = tile * T, nseg = min(T, np - ip0);
      __syncwarp(); // every lane is done with the other buffer
      const bool all_dead = __all_sync(0xffffffffu, dead);
      if (lane == 0 && !all_dead && tile + 1 < ntiles) { // the next tile travels while this one is computed
        const unsigned bytes = (unsigned)(min(T, np - ip0 - T) * L.rec) * 8u;
        fast::mbar_expect_tx(&bars[b ^ 1], bytes);
        fast::tma_load_1d(recbuf + (size_t)(b ^ 1) * T * L.rec, rec_g + (size_t)(ip0 + T) * L.rec, bytes, &bars[b ^ 1]);
      }
      if (b == 0) { fast::mbar_wait(&bars[0], parity0); parity0 ^= 1; } else { fast::mbar_wait(&bars[1], parity1); parity1 ^= 1; }
      if (all_dead) break;
      if (dead) continue;
      const double *__restrict__ RT = recbuf + (size_t)b * T * L.rec;

#pragma unroll
      for (int s = 0; s < T; s++) prod_s[s * sstride] = 1.0;
      bool any_opaque = false;

#pragma unroll 1
      for (int ig = g0; ig < g1; ig++) {
        unsigned long long h = hint_s[(ig - g0) * sstride];
        if (h == ~0ull) continue; // this (gas, channel) pair has no table: factor 1 in every segment
        double tp = tau_s[(ig - g0) * sstride];
        unsigned ccell = (unsigned)(h >> 40);      // cell the bracket indices in h belong to
        bool loaded = false;                       // descriptors / brackets in registers are valid for ccell
        uint2 c00 = make_uint2(0, 0), c01 = c00, c10 = c00, c11 = c00;
        float4 b00 = make_float4(0, 0, 0, 0), b01 = b00, b10 = b00, b11 = b00;
        int k00 = 0, k01 = 0, k10 = 0, k11 = 0;
        bool usable = false;                       // all four columns of the cell have >= 2 entries
        unsigned unsorted = 0;
#pragma unroll 1
        for (int s = 0; s < nseg; s++) {
          const double *__restrict__ R = RT + (size_t)s * L.rec;
          double f;
          if (tp < 1e-9) {
            f = 0.0;
            any_opaque = true;
          } else {
            f = 1.0;
            const unsigned cell = fast::load_cell(R, L, ig);
            if (cell != kCellInvalid) {
              if (!loaded || cell != ccell) {
                // the ray entered another table cell (or the tile starts): descriptors, hints remapped by column identity,
                // the four hinted brackets requested back to back
                if (loaded) h = (unsigned long long)k00 | ((unsigned long long)k01 << 10) | ((unsigned long long)k10 << 20) | ((unsigned long long)k11 << 30);
                if (ccell != cell) h = fast::remap_hints(h, ccell, cell);
                ccell = cell; loaded = true;
                fast::load_coldesc(Tb, ig, cell, nd, id, c00, c01, c10, c11);
                unsorted = ROBUST ? ((c00.y | c01.y | c10.y | c11.y) & kColNonMonotone) : 0u;
                if (ROBUST) { c00.y &= ~kColNonMonotone; c01.y &= ~kColNonMonotone; c10.y &= ~kColNonMonotone; c11.y &= ~kColNonMonotone; }
                usable = c00.y >= 2 && c01.y >= 2 && c10.y >= 2 && c11.y >= 2;
                if (usable) {
                  k00 = min((int)(h & 0x3ffu), (int)c00.y - 2); k01 = min((int)((h >> 10) & 0x3ffu), (int)c01.y - 2);
                  k10 = min((int)((h >> 20) & 0x3ffu), (int)c10.y - 2); k11 = min((int)((h >> 30) & 0x3ffu), (int)c11.y - 2);
                  if (!ROBUST || !unsorted) {
                    const float4 *__restrict__ brk = Tb.brk;
                    b00 = brk[c00.x + (unsigned)k00]; b01 = brk[c01.x + (unsigned)k01];
                    b10 = brk[c10.x + (unsigned)k10]; b11 = brk[c11.x + (unsigned)k11];
                  }
                } else {
                  k00 = (int)(h & 0x3ffu); k01 = (int)((h >> 10) & 0x3ffu); k10 = (int)((h >> 20) & 0x3ffu); k11 = (int)((h >> 30) & 0x3ffu);
                }
              }
              if (usable) {
                const float4 *__restrict__ brk = Tb.brk;
                const double *__restrict__ cw = R + L.c0 + L.cstride * ig;
                const double eps = 1 - tp, useg = R[L.u0 + ig];
                double e00, e01, e10, e11;
                if (!ROBUST || !unsorted) {
                  fast::column_finish4(brk, c00.x, c01.x, c10.x, c11.x, (int)c00.y, (int)c01.y, (int)c10.y, (int)c11.y, eps, fast::round_down(eps),
                                            useg, k00, k01, k10, k11, b00, b01, b10, b11, e00, e01, e10, e11);
                } else {
                  e00 = fast::column_finish_bisect(brk + c00.x, (int)c00.y, eps, useg);
                  e01 = fast::column_finish_bisect(brk + c01.x, (int)c01.y, eps, useg);
                  e10 = fast::column_finish_bisect(brk + c10.x, (int)c10.y, eps, useg);
                  e11 = fast::column_finish_bisect(brk + c11.x, (int)c11.y, eps, useg);
                }
                const double ep0 = clamp01(fma(cw[1], e01 - e00, e00));
                const double ep1 = clamp01(fma(cw[2], e11 - e10, e10));
                const double ept = clamp01(fma(cw[0], ep1 - ep0, ep0));
                f = (1. - ept) * fast_rcp(tp);
              }
            }
            const double tn = tp * f;
            tp = tn;
            any_opaque |= tn < 1e-9;
          }
          prod_s[s * sstride] *= f;
        }
        tau_s[(ig - g0) * sstride] = tp;
        if (loaded)
          hint_s[(ig - g0) * sstride] = (unsigned long long)k00 | ((unsigned long long)k01 << 10) | ((unsigned long long)k10 << 20) |
                                 ((unsigned long long)k11 << 30) | ((unsigned long long)ccell << 40);
      }

      if constexpr (SPLIT) { // the block's products of this tile (exact zeros from the segment on where a gas went opaque)
        if (lane_on)
          for (int s = 0; s < nseg; s++) part[(size_t)(ip0 + s) * nd] = prod_s[s * sstride];
        n_done = ip0 + nseg;
      } else { // radiance update of the tile (continua_core_bbbb, src_planck_core, new_obs_core)
#pragma unroll 1
        for (int s = 0; s < nseg; s++) {
          const double *__restrict__ R = RT + (size_t)s * L.rec;
          const double p = R[0], t = R[1], ds = R[2];
          const double u_co2 = (MASK & 8) ? R[L.u0 + a.ig_co2] : 0.0;
          const double u_h2o = (MASK & 4) ? R[L.u0 + a.ig_h2o] : 0.0;
          const double beta_ds = continuum_beta_ds(MASK, a.chan, nd, id, p, t, ds, a.nw > 0 ? R[4 + win] : 0.0, u_co2, u_h2o, R[3]);
          const double tau_gas = prod_s[s * sstride];
          const double src = planck_source(Tb.sr, nd, id, t);
          accumulate(rad, tau, beta_ds, src, tau_gas);
        }
      }
      // an opaque gas keeps its factor 0: every later product is 0 and nothing is accumulated any more
      if (any_opaque && prod_s[(nseg - 1) * sstride] == 0.0) dead = true;
    }
    if constexpr (SPLIT) {
      if (lane_on) a.partial_len[((size_t)gblk * (size_t)a.n_rays + (size_t)ir) * nd + id] = n_done;
    } else {
      epilogue(rad, tau, a.ray_tsurf[ir], Tb.sr, nd, id, a.write_bbt, a.chan[CH_NU * nd + id]);
      if (lane_on) {
        a.rad[(size_t)ir * nd + id] = rad;
        a.tau[(size_t)ir * nd + id] = tau;
        if (a.rad_host) {
          a.rad_host[ir][id] = rad;
          a.tau_host[ir][id] = tau;
        }
      }
    }
  }
}

// counting sort of the last n rays of a launch by their number of segments, descending: perm[k] = index (within the tail)
// of the k-th longest ray.  One CTA; n is a few thousand.
static __global__ void __launch_bounds__(1024) tail_sort_kernel(const int *__restrict__ np_tail, const int n, int *__restrict__ perm) {
  __shared__ int bins[kNLOS + 2];
  for (int i = threadIdx.x; i < kNLOS + 2; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&bins[min(max(np_tail[i], 0), kNLOS)], 1);
  __syncthreads();
  if (threadIdx.x == 0) { // start offsets, longest first
    int acc = 0;
    for (int b = kNLOS; b >= 0; b--) { const int c = bins[b]; bins[b] = acc; acc += c; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[atomicAdd(&bins[min(max(np_tail[i], 0), kNLOS)], 1)] = i;
}

template <int MASK, bool ROBUST, bool SPLIT = false>
cudaError_t launch_ega_tiled_tm(const EgaArgs &a, cudaStream_t stream, int sm_count, int *n_launched = nullptr) {
  int nl = 0;
  const int ng_state = SPLIT ? a.gases_per_block : a.ng;
  const long long n_items = a.n_rays * ((a.nd + 31) / 32) * (SPLIT ? a.n_gas_blocks : 1);
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int block = kEgaBlock;
  if (const char *s = getenv("JRB_EGA_THREADS")) { const int v = atoi(s); if (v >= 32 && v <= kEgaBlock && v % 32 == 0) block = v; } // experiments
  else if (n_items < 16ll * sm_count * (kEgaBlock / 32)) block = kEgaSmallBlock; // small batches: finer tail (as in launch_ega_fast_tm)
  while (block > 32 && ega_tiled_smem_bytes(ng_state, a.los.rec, block) > (size_t)smem_max) block -= 32;
  const size_t smem = ega_tiled_smem_bytes(ng_state, a.los.rec, block);
  if (smem > (size_t)smem_max) return cudaErrorInvalidConfiguration;
  static std::atomic<unsigned long long> attr_done{0};
  cudaError_t e = cudaSuccess;
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    e = cudaFuncSetAttribute(ega_tiled_kernel<MASK, ROBUST, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  int blocks_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, ega_tiled_kernel<MASK, ROBUST, SPLIT>, block, smem);
  if (e != cudaSuccess) return e;
  if (blocks_per_sm < 1) return cudaErrorInvalidConfiguration;
  EgaArgs args = a;
  if (args.work_chunk <= 0) args.work_chunk = block / 32;
  if (args.phase_lock_mode < 0 && args.balance != nullptr && a.n_rays > 0) { // let the device decide: equal-length chunks -> lock step
    const long long n_chunks = (n_items + block / 32 - 1) / (block / 32);
    chunk_balance_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, stream>>>(a.ray_np, a.n_rays, 1, block / 32, args.balance);
    nl++;
  }
  long long grid = (long long)sm_count * blocks_per_sm;
  const long long need = (n_items + block / 32 - 1) / (block / 32);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  // the last two rounds of rays longest first (only worth it when the launch has many rounds)
  args.tail_n = 0;
  if (SPLIT && args.tail_perm && a.n_rays > 0 && a.n_rays <= args.tail_cap && !getenv("JRB_NO_TAIL_SORT")) {
    args.tail_n = (int)a.n_rays; // latency mode: the whole (small) launch in order of decreasing ray length
    tail_sort_kernel<<<1, 1024, 0, stream>>>(a.ray_np, (int)a.n_rays, args.tail_perm);
    nl++;
  }
  if (!SPLIT && args.tail_perm && !getenv("JRB_NO_TAIL_SORT")) {
    const long long tail = 2 * grid * (block / 32);
    if (tail <= args.tail_cap && a.n_rays >= 4 * tail) {
      args.tail_n = (int)tail;
      tail_sort_kernel<<<1, 1024, 0, stream>>>(a.ray_np + (a.n_rays - tail), (int)tail, args.tail_perm);
      nl++;
    }
  }
  ega_tiled_kernel<MASK, ROBUST, SPLIT><<<(unsigned)grid, block, smem, stream>>>(args);
  if (n_launched) *n_launched = nl + 1;
  return cudaGetLastError();
}

} // namespace jrb
