// jrb_ega_fast.cuh -- specialised EGA kernels, template <continuum MASK> (the 16 variants the reference stamps out
// with src/jr_multiversion4gases.h are template instantiations here).
//
// Mapping: one warp = one ray x 32 consecutive channels (lane = channel), or fewer channels of several consecutive rays
// (MULTI).  Persistent CTAs draw work in chunks of consecutive rays, one ray per warp, from a global counter -- rays of
// different length (130..393 segments) balance themselves, and the warps of an SM work on neighbouring rays, whose table
// brackets they share through L1; CTAs run in lock step when the rays of a chunk are equally long (next_item and the
// work loop of the kernel).
//
// Data movement per segment:
//   * the head of the ray's LOS record (p, T, ds, extinction, per-gas u, table cell + interpolation weights; 112 B for
//     5 gases sharing one (p,T) grid) is staged into shared memory by a TMA bulk copy (cp.async.bulk + mbarrier, double buffered per warp): segment ip+1 is in
//     flight while segment ip is computed, and every lane reads the fields as shared-memory broadcasts;
//   * per gas, the four column descriptors are read coalesced (channel innermost), then the four hinted brackets
//     (one aligned 16-byte load each) are requested back to back so their L2 latencies overlap.
// State: the along-ray recurrence (rad, tau) lives in registers and is stored once per ray; the per-gas path
// transmittances tau_path[ng] and the bracket hints live in shared memory ([gas][thread], conflict free) so that the
// gas loop stays rolled -- an unrolled body was > 100 KB of SASS and stalled on instruction fetch (profiles/).
//
// Search-free table access: every (gas, column slot) keeps the bracket index it ended on in the previous segment as a
// 10-bit hint (remapped by column identity when the ray moves to a neighbouring (p,T) cell).  The next lookup loads that
// bracket first and only steps / jumps / bisects when the hint is off.  For monotone
// columns the index found equals the reference's full bisection (locate_tbl_id, src/jr_common.h:116-125), so results
// do not depend on the hints.
#pragma once
#include "jrb_ega_common.cuh"
#include <atomic>
#include <cstdlib>

namespace jrb {

namespace fast {

// ---- mbarrier / TMA bulk copy (PTX) -----------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// the same with an L2 eviction-priority hint: line-of-sight records are read once per (ray, channel group) and should not
// push the table brackets -- which every warp keeps coming back to -- out of the L2
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar,
                                                 unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
               : "memory");
}

// ---- bracket search ------------------------------------------------------------------------------------------------
// Exact comparison of a double x with float table values: for any float v,  v > x  <=>  v > rd(x), where rd() rounds x
// to float toward -infinity.  The bracket tests therefore run in single precision -- no float->double conversion per
// tested value -- and still reproduce the reference's float-vs-double comparisons (src/jr_common.h:116-125) exactly.
__device__ __forceinline__ float round_down(const double x) { return __double2float_rd(x); }

// slow path, shared by all call sites: reference bisection restricted to [ilo, ihi]; xd = rd(x)
static __device__ __noinline__ int search_range(const float4 *__restrict__ col, int ilo, int ihi, const float xd, const int on_eps) {
  while (ihi > ilo + 1) {
    const int i = (ihi + ilo) >> 1;
    const float v = on_eps ? col[i].y : col[i].x;
    if (v > xd) ihi = i; else ilo = i;
  }
  return ilo;
}

template <bool ON_EPS>
__device__ __forceinline__ float lo_of(const float4 b) { return ON_EPS ? b.y : b.x; }
template <bool ON_EPS>
__device__ __forceinline__ float hi_of(const float4 b) { return ON_EPS ? b.w : b.z; }

// Move (k, b) so that  val[k] <= x < val[k+1]  with k clipped to [0, nu-2]  (== reference bisection result); xd = rd(x).
// The neighbouring bracket is tried first.  Larger upward moves of the column-density lookup (long segments near the
// tangent point add several grid steps at once) jump by the distance estimated from the local grid ratio -- the u axes
// of JURASSIC tables are geometric -- and whatever is still not bracketed goes to the out-of-line bisection.
template <bool ON_EPS>
__device__ __forceinline__ void relocate(const float4 *__restrict__ brk, const unsigned first, const int nu, const float xd, int &k,
                                         float4 &b) {
  // brackets are addressed as brk[first + k] with a 32-bit index: one IMAD.WIDE per load instead of a 64-bit pointer chain
  if (lo_of<ON_EPS>(b) > xd) { // x < val[k]
    if (k > 0) {
      --k; b = brk[first + (unsigned)k];
      if (k > 0 && lo_of<ON_EPS>(b) > xd) { k = search_range(brk + first, 0, k, xd, ON_EPS); b = brk[first + (unsigned)k]; }
    }
  } else if (hi_of<ON_EPS>(b) <= xd) { // x >= val[k+1]
    if (k < nu - 2) {
      ++k; b = brk[first + (unsigned)k];
      if (k < nu - 2 && hi_of<ON_EPS>(b) <= xd) {
        if (!ON_EPS) {
          const float l0 = __log2f(b.x), r = __log2f(b.z) - l0;
          const float d = (r > 0.f) ? fminf((__log2f(xd) - l0) * __frcp_rn(r), 65535.f) : 1.f;
          const int k0 = k;
          k = min(k0 + max((int)d, 1), nu - 2);
          b = brk[first + (unsigned)k];
          if (b.x > xd) { k = search_range(brk + first, k0, k, xd, 0); b = brk[first + (unsigned)k]; }
          else if (k < nu - 2 && b.z <= xd) { k = search_range(brk + first, k, nu - 1, xd, 0); b = brk[first + (unsigned)k]; }
        } else {
          k = search_range(brk + first, k, nu - 1, xd, 1); b = brk[first + (unsigned)k];
        }
      }
    }
  }
}

// The four table columns of an EGA step, starting from the prefetched brackets (k, b): per column u* = u(eps) (get_u, may
// extrapolate), then eps(u* + u_seg) clamped to [0,1] (get_eps + c01, src/jr_common.h:249-257).  Written stage by stage
// over the four columns so that their independent FP64 chains interleave (D -0.5 %, E -1.7 % against column after column).
// (k, b) are updated in place: on return b is the bracket at k, where the column-density lookup ended.
__device__ __forceinline__ void column_finish4(const float4 *__restrict__ brk, const unsigned f0, const unsigned f1, const unsigned f2,
                                               const unsigned f3, const int n0, const int n1, const int n2, const int n3, const double eps,
                                               const float epsd, const double useg, int &k0, int &k1, int &k2, int &k3, float4 &b0, float4 &b1,
                                               float4 &b2, float4 &b3, double &r0, double &r1, double &r2, double &r3) {
  relocate<true>(brk, f0, n0, epsd, k0, b0);
  relocate<true>(brk, f1, n1, epsd, k1, b1);
  relocate<true>(brk, f2, n2, epsd, k2, b2);
  relocate<true>(brk, f3, n3, epsd, k3, b3);
  const double x0 = lerp_fast((double)b0.y, (double)b0.x, (double)b0.w, (double)b0.z, eps) + useg;
  const double x1 = lerp_fast((double)b1.y, (double)b1.x, (double)b1.w, (double)b1.z, eps) + useg;
  const double x2 = lerp_fast((double)b2.y, (double)b2.x, (double)b2.w, (double)b2.z, eps) + useg;
  const double x3 = lerp_fast((double)b3.y, (double)b3.x, (double)b3.w, (double)b3.z, eps) + useg;
  const float d0 = round_down(x0), d1 = round_down(x1), d2 = round_down(x2), d3 = round_down(x3);
  if (b0.x > d0 || b0.z <= d0) relocate<false>(brk, f0, n0, d0, k0, b0);
  if (b1.x > d1 || b1.z <= d1) relocate<false>(brk, f1, n1, d1, k1, b1);
  if (b2.x > d2 || b2.z <= d2) relocate<false>(brk, f2, n2, d2, k2, b2);
  if (b3.x > d3 || b3.z <= d3) relocate<false>(brk, f3, n3, d3, k3, b3);
  r0 = clamp01(lerp_fast((double)b0.x, (double)b0.y, (double)b0.z, (double)b0.w, x0));
  r1 = clamp01(lerp_fast((double)b1.x, (double)b1.y, (double)b1.z, (double)b1.w, x1));
  r2 = clamp01(lerp_fast((double)b2.x, (double)b2.y, (double)b2.z, (double)b2.w, x2));
  r3 = clamp01(lerp_fast((double)b3.x, (double)b3.y, (double)b3.z, (double)b3.w, x3));
}

// The same for a cell that contains a column which is not sorted in u or eps (flagged at pack time): the reference's
// plain bisection over the whole column, out of line -- sorted columns give the same result either way.
static __device__ __noinline__ double column_finish_bisect(const float4 *__restrict__ col, const int nu, const double eps,
                                                           const double useg) {
  int k = search_range(col, 0, nu - 1, round_down(eps), 1);
  float4 b = col[k];
  const double ustar = lerp_fast((double)b.y, (double)b.x, (double)b.w, (double)b.z, eps);
  const double x = ustar + useg;
  k = search_range(col, 0, nu - 1, round_down(x), 0);
  b = col[k];
  return clamp01(lerp_fast((double)b.x, (double)b.y, (double)b.z, (double)b.w, x));
}

// table cell of gas ig in the staged LOS record: ipr | it0 << 8 | it1 << 16, or kCellInvalid
__device__ __forceinline__ unsigned load_cell(const double *__restrict__ R, const LosLayout &L, const int ig) {
  return (unsigned)__double_as_longlong(R[L.c0 + L.cstride * ig + 3]);
}

// the four column descriptors of a cell, one 8-byte load each (coalesced over the channels of a warp)
__device__ __forceinline__ void load_coldesc(const TblDev &T, const int ig, const unsigned cell, const int nd, const int id,
                                             uint2 &c00, uint2 &c01, uint2 &c10, uint2 &c11) {
  const unsigned c = (cell == kCellInvalid) ? 0u : cell; // any valid address; the result is ignored for an invalid cell
  const int ipr = c & 0xff, it0 = (c >> 8) & 0xff, it1 = (c >> 16) & 0xff;
  // 32-bit index arithmetic (the descriptor array has ng*npmax*ntmax*nd < 2^31 elements), one 64-bit address per level
  const unsigned row = (unsigned)ig * (unsigned)T.npmax + (unsigned)ipr;
  const uint2 *__restrict__ q0 = T.col + ((row * (unsigned)T.ntmax + (unsigned)it0) * (unsigned)nd + (unsigned)id);
  const uint2 *__restrict__ q1 = T.col + (((row + 1u) * (unsigned)T.ntmax + (unsigned)it1) * (unsigned)nd + (unsigned)id);
  c00 = __ldg(q0); c01 = __ldg(q0 + nd); c10 = __ldg(q1); c11 = __ldg(q1 + nd);
}

// Hints follow the COLUMN, not the slot: when the ray moves to a neighbouring (p,T) cell, every new slot inherits the
// hint of the old slot that addressed the same column (or, failing that, the nearest one on the nearest level).
// h = 4 x 10-bit bracket indices | cell << 40.  Returns the remapped hints (cell bits are rewritten by the caller).
__device__ __forceinline__ unsigned long long remap_hints(const unsigned long long h, const unsigned ocell, const unsigned cell) {
  const int oipr = ocell & 0xff, oit0 = (ocell >> 8) & 0xff, oit1 = (ocell >> 16) & 0xff;
  const int ipr = cell & 0xff, it0 = (cell >> 8) & 0xff, it1 = (cell >> 16) & 0xff;
  unsigned long long out = 0;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int ip = ipr + (s >> 1), it = ((s >> 1) ? it1 : it0) + (s & 1);
    const bool upper = (ip == oipr + 1) || (ip != oipr && ip > oipr); // old level to borrow from
    const int oit = upper ? oit1 : oit0;
    const int sel = (upper ? 2 : 0) + (it > oit ? 1 : 0);
    out |= ((h >> (10 * sel)) & 0x3ffull) << (10 * s);
  }
  return out;
}

// ---- channel-dependent (p,T) axes -----------------------------------------------------------------------------------------
// init_tbl stores the pressure and temperature axes per (gas, channel) (src/jurassic.c:383-384), and the reference locates
// them per channel (locate_id, src/jr_common.h:237-247).  When the axes of a table set really differ between channels the
// table cell cannot be resolved once per ray and segment by the tracer; each lane (= channel) then locates its own cell,
// starting from the cell it used for the previous segment (kept in the hint word), and computes its own interpolation
// weights.  The axis arrays are channel-innermost, so the lanes of a warp read them coalesced.
//
// reference bisection result for an ascending axis of n points: max{i <= n-2 : ax[i] <= x}, 0 if x < ax[0]
static __device__ __noinline__ int axis_bisect(const double *__restrict__ ax, const int stride, const int n, const double x) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) {
    const int i = (ihi + ilo) >> 1;
    if (ax[(size_t)i * stride] > x) ihi = i; else ilo = i;
  }
  return ilo;
}
// index as above, tried at `guess` and its neighbours first; returns the weight (x - ax[i]) / (ax[i+1] - ax[i])
__device__ __forceinline__ int axis_locate(const double *__restrict__ ax, const int stride, const int n, const double x, const int guess,
                                           double &w) {
  int i = min(guess, n - 2);
  double x0 = ax[(size_t)i * stride], x1 = ax[(size_t)(i + 1) * stride];
  if (!((x0 <= x || i == 0) && (x1 > x || i == n - 2))) {
    i = (x0 > x) ? max(i - 1, 0) : min(i + 1, n - 2);
    x0 = ax[(size_t)i * stride]; x1 = ax[(size_t)(i + 1) * stride];
    if (!((x0 <= x || i == 0) && (x1 > x || i == n - 2))) {
      i = axis_bisect(ax, stride, n, x);
      x0 = ax[(size_t)i * stride]; x1 = ax[(size_t)(i + 1) * stride];
    }
  }
  w = (x - x0) * fast_rcp(x1 - x0);
  return i;
}
// this lane's table cell of gas ig at (p, t) and its three interpolation weights; kCellInvalid if an axis is too short
__device__ __forceinline__ unsigned locate_cell_lane(const TblDev &T, const int ig, const int nd, const int id, const double p, const double t,
                                                     const unsigned guess, double &wp, double &wt0, double &wt1) {
  const int np = T.np[ig * nd + id];
  if (np < 2) return kCellInvalid;
  const size_t gbase = (size_t)ig * T.npmax;
  const int ipr = axis_locate(T.pax + gbase * nd + id, nd, np, p, (int)(guess & 0xff), wp);
  const int nt0 = T.nt[(gbase + ipr) * nd + id], nt1 = T.nt[(gbase + ipr + 1) * nd + id];
  if (nt0 < 2 || nt1 < 2) return kCellInvalid;
  const double *__restrict__ t0ax = T.tax + (gbase + ipr) * T.ntmax * nd + id;
  const int it0 = axis_locate(t0ax, nd, nt0, t, (int)((guess >> 8) & 0xff), wt0);
  const int it1 = axis_locate(t0ax + (size_t)T.ntmax * nd, nd, nt1, t, (int)((guess >> 16) & 0xff), wt1);
  return (unsigned)ipr | ((unsigned)it0 << 8) | ((unsigned)it1 << 16);
}

// Work distribution.  Items (rays, or ray x channel group) are handed out in CHUNKS of consecutive items per CTA: the
// warps of a CTA draw single items from the CTA's current chunk (shared-memory word: chunk base << 8 | items taken) and
// whoever finds it exhausted fetches the next chunk from the global counter.  Consecutive items are neighbouring rays of
// one scan / swath: they cross the same (p,T) cells with nearly the same column amounts, so the warps that run side by
// side on an SM gather the same or neighbouring brackets and share them through L1.  Still dynamic at chunk granularity,
// so rays of different length balance themselves.
constexpr unsigned long long kChunkEmpty = 0xffull;  // "items taken" = 255: no chunk yet / exhausted
constexpr unsigned long long kChunkLocked = 0xfeull; // a warp is fetching the next chunk
__device__ __forceinline__ unsigned long long next_item(unsigned long long *state, unsigned long long *global_counter, const unsigned chunk) {
  for (;;) {
    const unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(state);
    const unsigned taken = (unsigned)(old & 0xffull);
    if (taken == (unsigned)kChunkLocked) continue;                      // another warp is refilling
    if (taken < chunk) {
      if (atomicCAS(state, old, old + 1ull) == old) return (old >> 8) + taken;
      continue;
    }
    if (atomicCAS(state, old, kChunkLocked) != old) continue;           // somebody else was faster
    const unsigned long long base = atomicAdd(global_counter, (unsigned long long)chunk);
    atomicExch(state, (base << 8) | 1ull);
    return base;
  }
}

} // namespace fast

// 24 warps per SM at 80 registers is the measured optimum (profiles/README.md).  Large batches run them as ONE 768-thread
// CTA per SM, so that all 24 warps draw neighbouring rays from the same work chunk (L1 sharing, see next_item); small
// batches, and gas counts whose per-thread state does not fit one CTA's shared memory, use 256-thread CTAs (3 per SM),
// which spread a single package over all SMs.
#ifndef JRB_EGA_BLOCK
#define JRB_EGA_BLOCK 768
#endif
constexpr int kEgaBlock = JRB_EGA_BLOCK; // launch bound (register budget); the launched block is kEgaBlock or kEgaSmallBlock
constexpr int kEgaSmallBlock = 256;
#ifndef JRB_EGA_MINBLOCKS
#define JRB_EGA_MINBLOCKS 1
#endif


__host__ __device__ inline size_t ega_fast_smem_bytes(int ng, int rec, int threads, int rpw) {
  const int nwarps = threads / 32;
  return 16                                      // work-chunk state of the CTA
         + (size_t)nwarps * 2 * 8                // mbarriers
         + (size_t)nwarps * 2 * rpw * rec * 8    // LOS record double buffers (one record per ray of the warp)
         + (size_t)ng * threads * 16;            // tau_path + hints
}

// MULTI = false: one ray per warp, lane = channel of a 32-channel group.
// MULTI = true : a warp handles cpw < 32 channels of floor(32/cpw) consecutive rays (lane = ray_in_warp * cpw + channel); every
//                ray of the warp gets its own staged record.  Used (a) for few-channel instruments, cpw = nd <= 16 -- the
//                reference's own examples have 2 and 3 channels, which would leave 90 % of the lanes idle -- and (b) to
//                narrow the channel group when (32 channels x ng gases) of tables do not fit the L2: work is channel-group
//                major, so the hot table set is cpw channels x ng gases (EgaArgs::cpw, chosen by the runtime).
// ROBUST = true: the table set contains columns that are not sorted in u or eps (flagged kColNonMonotone at pack time);
//                cells touching such a column are evaluated with the reference's plain bisection.  The ROBUST = false
//                instantiation is used for fully sorted table sets and carries no trace of this.
// SPLIT = true : gas-block pass.  The gases are cut into a.n_gas_blocks blocks of a.gases_per_block and a work item is
//                (gas block, channel group, ray): the warp runs the EGA recurrence for ITS gases only and stores, per
//                segment, the product of their factors to a.partial[block][ray][segment][channel] (continuum, source and
//                accumulation are left to ega_combine_kernel).  The gases of a ray do not depend on each other, so this
//                (a) multiplies the number of independent warps by the number of blocks -- a single 1088-ray package fills
//                the GPU (latency mode) -- and (b) bounds the per-thread state (16 B per gas) and the hot table set
//                (channels per warp x gases per block) for many-gas set-ups such as the 30-gas refspec shape.  With one
//                gas per block the combine step multiplies the factors in the same order as the fused kernel: bit-identical.
//                MASK plays no role in this pass (instantiated for MASK = 0 only).
// PERCH = true : the (p,T) axes of the tables depend on the channel: every lane locates its own table cell and weights
//                (locate_cell_lane) instead of taking the ray's cell from the line-of-sight record.  Instantiated with
//                ROBUST = true only (such table sets are rare; one variant serves sorted and unsorted columns).
template <int MASK, bool MULTI, bool ROBUST, bool SPLIT = false, bool PERCH = false>
__global__ void __launch_bounds__(kEgaBlock, JRB_EGA_MINBLOCKS) ega_fast_kernel(const EgaArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const LosLayout L = a.los;
  const TblDev &T = a.tbl;
  const int nd = a.nd, ng = a.ng;
  const int ngs = SPLIT ? a.gases_per_block : ng;   // gases whose state a thread holds
  const unsigned rec_bytes = (unsigned)L.head * 8u; // only the head of a record is staged

  const int cpw = MULTI ? a.cpw : 32;           // channels of a ray handled by one warp
  const int rpw = MULTI ? 32 / cpw : 1;         // rays per warp
  const int bufstride = rpw * L.head;           // doubles per record buffer of a warp
  unsigned long long *chunk_state = reinterpret_cast<unsigned long long *>(smem_raw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + 16) + warp * 2;
  double *recbuf = reinterpret_cast<double *>(smem_raw + 16 + (size_t)nwarps * 16) + (size_t)warp * 2 * bufstride;
  double *tau_s = reinterpret_cast<double *>(smem_raw + 16 + (size_t)nwarps * 16 + (size_t)nwarps * 2 * bufstride * 8) + tid;
  unsigned long long *hint_s = reinterpret_cast<unsigned long long *>(tau_s - tid + (size_t)ngs * blockDim.x) + tid;
  const int sstride = blockDim.x;

  if (lane == 0) { fast::mbar_init(&bars[0], 1); fast::mbar_init(&bars[1], 1); }
  if (tid == 0) *chunk_state = fast::kChunkEmpty;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned parity0 = 0, parity1 = 0;
  const unsigned long long los_policy = a.los_evict_first ? fast::l2_policy_evict_first() : 0ull;
  // balance[0] = idle segment slots, balance[1] = all segment slots if every chunk of blockDim/32 consecutive items ran in
  // lock step (chunk_balance_kernel): lock step is chosen when less than 1/32 of the slots would idle
  const bool phase_lock = a.phase_lock_mode == 1 || (a.phase_lock_mode < 0 && a.balance != nullptr && a.balance[0] * 32ull < a.balance[1]);

  const int ngroups = (nd + cpw - 1) / cpw;
  const unsigned long long n_blocks = MULTI ? (unsigned long long)((a.n_rays + rpw - 1) / rpw) : (unsigned long long)a.n_rays;
  const unsigned long long n_items_blk = n_blocks * ngroups; // channel-group major: item = group * n_blocks + ray block
  const unsigned long long n_items = SPLIT ? n_items_blk * (unsigned long long)a.n_gas_blocks : n_items_blk; // gas-block major
  const int sub = MULTI ? lane / cpw : 0;       // ray of this lane within the warp

  for (;;) {
    unsigned long long item = 0;
    if (phase_lock) {
      // all warps of the CTA start their rays together and therefore stay in phase (same segment index = same altitude
      // range = same table cells): what one warp gathers, its neighbours find in L1.  Costs the idle time of the
      // shorter rays at the end of each chunk, hence only used when the rays of a chunk are of (nearly) equal length.
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
    int g0 = 0, g1 = ng, gblk = 0; // gases of this item
    if (SPLIT) {
      gblk = (int)(item / n_items_blk);
      item -= (unsigned long long)gblk * n_items_blk;
      g0 = gblk * a.gases_per_block;
      g1 = min(ng, g0 + a.gases_per_block);
    }
    long long ir;
    int id;
    bool lane_on;
    bool ray_on = true; // this lane's ray exists (MULTI: the lane may still be off if its channel is beyond nd)
    if (MULTI) {
      const int grp = (int)(item / n_blocks);
      ir = (long long)(item - (unsigned long long)grp * n_blocks) * rpw + sub;
      ray_on = sub < rpw && ir < a.n_rays;
      const int id_raw = grp * cpw + (lane - sub * cpw);
      lane_on = ray_on && id_raw < nd;
      if (!ray_on) ir = a.n_rays - 1; // idle lanes shadow a valid ray and never store
      id = lane_on ? id_raw : 0;
    } else {
      // channel-group major: all warps in flight work on the same 32 channels, so the part of the tables that is hot
      // at any time is (32 channels x ng gases), which is what has to fit into L2
      const int grp = (int)(item / (unsigned long long)a.n_rays);
      ir = (long long)(item - (unsigned long long)grp * (unsigned long long)a.n_rays);
      const int id_raw = grp * 32 + lane;
      lane_on = id_raw < nd;
      id = lane_on ? id_raw : nd - 1;
    }

    const double *__restrict__ rec_g = a.los_data + (size_t)ir * kNLOS * L.rec;
    const int np = ray_on ? a.ray_np[ir] : 0;                  // segments of this lane's ray
    const int np_max = MULTI ? __reduce_max_sync(0xffffffffu, np) : np;
    const bool head_lane = MULTI ? (ray_on && lane == sub * cpw) : (lane == 0); // issues the record copies of its ray
    const int win = a.window[id];

    for (int ig = g0; ig < g1; ig++) {
      // a (gas, channel) pair without table (np < 2) contributes the factor 1 (src/jr_common.h:240): hint = all ones
      tau_s[(ig - g0) * sstride] = 1.0;
      hint_s[(ig - g0) * sstride] = (T.np[ig * nd + id] >= 2) ? 0ull : ~0ull;
    }
    double rad = 0.0, tau = 1.0;
    bool dead = false; // a gas went opaque (tau_path < 1e-9): nothing changes any more (src/jr_common.h:239,295)
    int n_done = 0;    // SPLIT: segments for which this lane stored a block product
    double *__restrict__ part = nullptr;
    if (SPLIT) part = a.partial + (((size_t)gblk * (size_t)a.n_rays + (size_t)ir) * kNLOS) * (size_t)nd + id;

    __syncwarp();
    // one copy per ray of the warp and segment; the barrier of a buffer expects the bytes of all copies aimed at it
    {
      const unsigned cnt = MULTI ? __popc(__ballot_sync(0xffffffffu, head_lane && np > 0)) : (np > 0 ? 1u : 0u);
      if (cnt && lane == 0) fast::mbar_expect_tx(&bars[0], cnt * rec_bytes);
      __syncwarp();
      if (head_lane && np > 0) {
        if (los_policy) fast::tma_load_1d_hint(recbuf + (size_t)sub * L.head, rec_g, rec_bytes, &bars[0], los_policy);
        else fast::tma_load_1d(recbuf + (size_t)sub * L.head, rec_g, rec_bytes, &bars[0]);
      }
    }
    for (int ip = 0; ip < np_max; ++ip) {
      const int b = ip & 1;
      __syncwarp(); // every lane is done with the other buffer (segment ip-1)
      const bool all_dead = __all_sync(0xffffffffu, dead || ip >= np);
      {
        const bool want = !all_dead && head_lane && (ip + 1 < np); // segment ip+1 travels while segment ip is computed
        const unsigned cnt = MULTI ? __popc(__ballot_sync(0xffffffffu, want)) : (want ? 1u : 0u);
        if (MULTI) {
          if (cnt && lane == 0) fast::mbar_expect_tx(&bars[b ^ 1], cnt * rec_bytes);
          __syncwarp();
        } else if (want) {
          fast::mbar_expect_tx(&bars[b ^ 1], rec_bytes);
        }
        if (want) {
          if (los_policy)
            fast::tma_load_1d_hint(recbuf + (size_t)(b ^ 1) * bufstride + (size_t)sub * L.head, rec_g + (size_t)(ip + 1) * L.rec, rec_bytes,
                                   &bars[b ^ 1], los_policy);
          else
            fast::tma_load_1d(recbuf + (size_t)(b ^ 1) * bufstride + (size_t)sub * L.head, rec_g + (size_t)(ip + 1) * L.rec, rec_bytes,
                              &bars[b ^ 1]);
        }
      }
      // the copies of segment ip are always in flight here (issued above one iteration earlier, or before the loop)
      if (b == 0) { fast::mbar_wait(&bars[0], parity0); parity0 ^= 1; } else { fast::mbar_wait(&bars[1], parity1); parity1 ^= 1; }
      if (all_dead) break; // nothing further was requested
      if (dead || ip >= np) continue;

      const double *__restrict__ R = recbuf + (size_t)b * bufstride + (size_t)sub * L.head;
      double t = 0.0, beta_ds = 0.0;
      const double p_seg = PERCH ? R[0] : 0.0, t_seg = PERCH ? R[1] : 0.0; // per-lane cell location needs them in every pass
      if (!SPLIT) {
        const double p = R[0], ds = R[2];
        t = R[1];
        const double u_co2 = (MASK & 8) ? R[L.u0 + a.ig_co2] : 0.0;
        const double u_h2o = (MASK & 4) ? R[L.u0 + a.ig_h2o] : 0.0;
        beta_ds = continuum_beta_ds(MASK, a.chan, nd, id, p, t, ds, a.nw > 0 ? R[4 + win] : 0.0, u_co2, u_h2o, R[3]);
      }

      double tau_gas = 1.0;
      bool any_opaque = false;
      uint2 n00, n01, n10, n11;
      unsigned ncell;
#ifdef JRB_PREFETCH_NEXT_GAS
      // column descriptors of gas 0; inside the loop those of gas ig+1 are requested before gas ig is computed
      ncell = fast::load_cell(R, L, g0);
      fast::load_coldesc(T, g0, ncell, nd, id, n00, n01, n10, n11);
#endif
#pragma unroll 1
      for (int ig = g0; ig < g1; ig++) {
        double wl_p = 0.0, wl_t0 = 0.0, wl_t1 = 0.0; // PERCH: this lane's interpolation weights
#ifndef JRB_PREFETCH_NEXT_GAS
        if (PERCH) {
          const unsigned long long h0 = hint_s[(ig - g0) * sstride];
          ncell = (h0 == ~0ull) ? kCellInvalid : fast::locate_cell_lane(T, ig, nd, id, p_seg, t_seg, (unsigned)(h0 >> 40), wl_p, wl_t0, wl_t1);
        } else {
          ncell = fast::load_cell(R, L, ig);
        }
        fast::load_coldesc(T, ig, ncell, nd, id, n00, n01, n10, n11);
#endif
        const uint2 c00 = n00, c01 = n01, c10 = n10, c11 = n11;
        const unsigned cell = ncell;
#ifdef JRB_PREFETCH_NEXT_GAS // measured: no gain at 3 CTAs/SM (profiles/README.md), costs registers
        if (ig + 1 < g1) {
          ncell = fast::load_cell(R, L, ig + 1);
          fast::load_coldesc(T, ig + 1, ncell, nd, id, n00, n01, n10, n11);
        }
#endif
        const double tp = tau_s[(ig - g0) * sstride];
        double f;
        if (tp < 1e-9) {
          f = 0.0;
          any_opaque = true;
        } else {
          f = 1.0;
          unsigned long long h = hint_s[(ig - g0) * sstride];
          const unsigned unsorted = ROBUST ? ((c00.y | c01.y | c10.y | c11.y) & kColNonMonotone) : 0u;
          const unsigned nmask = ROBUST ? ~kColNonMonotone : ~0u;
          const unsigned n00u = c00.y & nmask, n01u = c01.y & nmask, n10u = c10.y & nmask, n11u = c11.y & nmask;
          if (h != ~0ull && cell != kCellInvalid && n00u >= 2 && n01u >= 2 && n10u >= 2 && n11u >= 2) {
            const unsigned ocell = (unsigned)(h >> 40);
            if (ocell != cell) h = fast::remap_hints(h, ocell, cell); // uniform per ray: the cell belongs to the ray
            const float4 *__restrict__ brk = T.brk;
            int k00 = min((int)(h & 0x3ffu), (int)n00u - 2), k01 = min((int)((h >> 10) & 0x3ffu), (int)n01u - 2),
                k10 = min((int)((h >> 20) & 0x3ffu), (int)n10u - 2), k11 = min((int)((h >> 30) & 0x3ffu), (int)n11u - 2);
            const double *__restrict__ cw = R + L.c0 + L.cstride * ig;
            const double eps = 1 - tp, useg = R[L.u0 + ig];
            double e00, e01, e10, e11;
            if (!ROBUST || !unsorted) {
              // the four hinted brackets are requested back to back: their latencies overlap
              float4 b00 = brk[c00.x + (unsigned)k00], b01 = brk[c01.x + (unsigned)k01], b10 = brk[c10.x + (unsigned)k10],
                     b11 = brk[c11.x + (unsigned)k11];
              const float epsd = fast::round_down(eps);
              fast::column_finish4(brk, c00.x, c01.x, c10.x, c11.x, (int)n00u, (int)n01u, (int)n10u, (int)n11u, eps, epsd, useg, k00, k01, k10, k11,
                                   b00, b01, b10, b11, e00, e01, e10, e11);
            } else {
              e00 = fast::column_finish_bisect(brk + c00.x, (int)n00u, eps, useg); // (hints keep their old values)
              e01 = fast::column_finish_bisect(brk + c01.x, (int)n01u, eps, useg);
              e10 = fast::column_finish_bisect(brk + c10.x, (int)n10u, eps, useg);
              e11 = fast::column_finish_bisect(brk + c11.x, (int)n11u, eps, useg);
            }
            hint_s[(ig - g0) * sstride] = (unsigned long long)k00 | ((unsigned long long)k01 << 10) | ((unsigned long long)k10 << 20) |
                                   ((unsigned long long)k11 << 30) | ((unsigned long long)cell << 40);
            const double ep0 = clamp01(fma(PERCH ? wl_t0 : cw[1], e01 - e00, e00));
            const double ep1 = clamp01(fma(PERCH ? wl_t1 : cw[2], e11 - e10, e10));
            const double ept = clamp01(fma(PERCH ? wl_p : cw[0], ep1 - ep0, ep0));
            f = (1. - ept) * fast_rcp(tp);
          }
          const double tn = tp * f;
          tau_s[(ig - g0) * sstride] = tn;
          any_opaque |= tn < 1e-9;
        }
        tau_gas *= f;
      }
      // an opaque gas keeps its factor 0 for the rest of the ray: tau_gas stays 0, accumulate() is skipped for good
      if (tau_gas == 0.0 && any_opaque) dead = true;
      if (SPLIT) {
        if (lane_on) part[(size_t)ip * nd] = tau_gas; // product of this block's factors (0 from here on once a gas is opaque)
        n_done = ip + 1;
        continue;
      }
      const double src = planck_source(T.sr, nd, id, t);
      accumulate(rad, tau, beta_ds, src, tau_gas);
    }
    if (SPLIT) {
      // segments [0, n_done) carry a product; beyond it the block's factor is 0 (a gas went opaque at n_done - 1)
      if (lane_on) a.partial_len[((size_t)gblk * (size_t)a.n_rays + (size_t)ir) * nd + id] = n_done;
      continue;
    }
    epilogue(rad, tau, a.ray_tsurf[ir], T.sr, nd, id, a.write_bbt, a.chan[CH_NU * nd + id]);
    if (lane_on) {
      a.rad[(size_t)ir * nd + id] = rad;
      a.tau[(size_t)ir * nd + id] = tau;
      if (a.rad_host) { // the ray's rows in host-mapped memory: the result lands there while the kernel keeps running
        a.rad_host[ir][id] = rad;
        a.tau_host[ir][id] = tau;
      }
    }
  }
}

// Load balance of lock-step execution: for every chunk of `chunk` consecutive items (an item = `rpw` consecutive rays
// handled by one warp, its length the longest of them) the segment slots that would idle while the chunk's longest
// item finishes, and the total.  One thread per chunk; rays are short rows of ints, the kernel is negligible.
static __global__ void chunk_balance_kernel(const int *__restrict__ ray_np, const long long n_rays, const int rpw, const int chunk,
                                            unsigned long long *__restrict__ balance) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_items = (n_rays + rpw - 1) / rpw;
  unsigned long long idle = 0, total = 0;
  if (c * chunk < n_items) {
    int lmax = 0, cnt = 0;
    long long sum = 0;
    for (int j = 0; j < chunk; j++) {
      const long long item = c * chunk + j;
      if (item >= n_items) break;
      int len = 0;
      for (int r = 0; r < rpw; r++) { const long long ir = item * rpw + r; if (ir < n_rays) len = max(len, ray_np[ir]); }
      lmax = max(lmax, len); sum += len; cnt++;
    }
    total = (unsigned long long)lmax * (unsigned long long)cnt;
    idle = total - (unsigned long long)sum;
  }
  // warp-level reduction, one atomic pair per warp
  for (int o = 16; o > 0; o >>= 1) { idle += __shfl_down_sync(0xffffffffu, idle, o); total += __shfl_down_sync(0xffffffffu, total, o); }
  if ((threadIdx.x & 31) == 0 && total) { atomicAdd(&balance[0], idle); atomicAdd(&balance[1], total); }
}

template <int MASK, bool MULTI, bool ROBUST, bool SPLIT = false, bool PERCH = false>
cudaError_t launch_ega_fast_tm(const EgaArgs &a, cudaStream_t stream, int sm_count) {
  const int cpw = MULTI ? a.cpw : 32, rpw = 32 / cpw;
  const int ngroups = (a.nd + cpw - 1) / cpw;
  const long long n_items = ((a.n_rays + rpw - 1) / rpw) * ngroups * (SPLIT ? a.n_gas_blocks : 1);
  const int ng_state = SPLIT ? a.gases_per_block : a.ng; // gases whose state a thread holds
  // Block size.  Large batches: ONE 768-thread CTA per SM, so that all 24 warps share one work chunk.  Small batches (fewer
  // than 16 rounds of work per SM: the coarser chunks would cost more in the tail than the L1 sharing gains, measured
  // cross-over between 35 k and 125 k items) and gas counts whose per-thread state (16 B per gas and thread) does not
  // fit one such CTA use 256-thread CTAs, as many per SM as fit (3 up to 16 gases).  For the 30-gas refspec shape that
  // is 8 warps per SM; a 448-thread CTA (14 warps) was measured SLOWER there (226 vs 145 ms): its 960 (gas, channel)
  // table pairs exceed the L2 and more rays in flight only widen the working set.
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int block = kEgaBlock;
  if (const char *s = getenv("JRB_EGA_THREADS")) { const int v = atoi(s); if (v >= 32 && v <= kEgaBlock && v % 32 == 0) block = v; } // experiments
  else if (n_items < 16ll * sm_count * (kEgaBlock / 32)) block = kEgaSmallBlock;
  // few channels x many gases (rpw records per warp + 16 B of state per gas and thread) can exceed even the small CTA's
  // shared memory: shrink the CTA until it fits (jrb_stage has checked that a one-warp CTA fits, else the generic kernel runs)
  while (block > 32 && ega_fast_smem_bytes(ng_state, a.los.head, block, rpw) > (size_t)smem_max) block = block > kEgaSmallBlock ? kEgaSmallBlock : block / 2;
  const size_t smem = ega_fast_smem_bytes(ng_state, a.los.head, block, rpw);
  if (smem > (size_t)smem_max) return cudaErrorInvalidConfiguration;
  // the dynamic shared-memory limit is a per-device property of the function: set once to the opt-in maximum, so that
  // contexts launching concurrently with different sizes (lanes, several host threads) cannot lower it under each other
  static std::atomic<unsigned long long> attr_done{0};
  cudaError_t e = cudaSuccess;
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    e = cudaFuncSetAttribute(ega_fast_kernel<MASK, MULTI, ROBUST, SPLIT, PERCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  if (const char *s = getenv("JRB_EGA_CARVEOUT")) { // experiments: shared-memory carve-out in percent (the rest of the 256 KB is L1)
    const int v = atoi(s);
    if (v >= 0 && v <= 100) cudaFuncSetAttribute(ega_fast_kernel<MASK, MULTI, ROBUST, SPLIT, PERCH>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
  }
  int blocks_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, ega_fast_kernel<MASK, MULTI, ROBUST, SPLIT, PERCH>, block, smem);
  if (e != cudaSuccess) return e;
  if (blocks_per_sm < 1) return cudaErrorInvalidConfiguration;
  if (const char *s = getenv("JRB_EGA_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v < blocks_per_sm) blocks_per_sm = v; } // occupancy experiments
  EgaArgs args = a;
  if (args.work_chunk <= 0) args.work_chunk = block / 32; // one item per warp of the CTA: the warps of a CTA stay on neighbouring rays
  if (args.phase_lock_mode < 0 && args.balance != nullptr && a.n_rays > 0) { // let the device decide: equal-length chunks -> lock step
    const long long n_chunks = (n_items + block / 32 - 1) / (block / 32);
    chunk_balance_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, stream>>>(a.ray_np, a.n_rays, rpw, block / 32, args.balance);
  }
  long long grid = (long long)sm_count * blocks_per_sm; // persistent: a whole number of CTAs per SM
  const long long per_cta = args.work_chunk > block / 32 ? args.work_chunk : block / 32;
  const long long need = (n_items + per_cta - 1) / per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  ega_fast_kernel<MASK, MULTI, ROBUST, SPLIT, PERCH><<<(unsigned)grid, block, smem, stream>>>(args);
  return cudaGetLastError();
}

template <int MASK>
cudaError_t launch_ega_fast_t(const EgaArgs &a, cudaStream_t stream, int sm_count) {
  const bool multi = a.cpw < 32; // several rays per warp
  if (a.per_channel_axes)
    return multi ? launch_ega_fast_tm<MASK, true, true, false, true>(a, stream, sm_count) : launch_ega_fast_tm<MASK, false, true, false, true>(a, stream, sm_count);
  if (a.unsorted_columns)
    return multi ? launch_ega_fast_tm<MASK, true, true>(a, stream, sm_count) : launch_ega_fast_tm<MASK, false, true>(a, stream, sm_count);
  return multi ? launch_ega_fast_tm<MASK, true, false>(a, stream, sm_count) : launch_ega_fast_tm<MASK, false, false>(a, stream, sm_count);
}

// gas-block pass (SPLIT): independent of the continuum mask, one set of instantiations (jrb_ega_split.cu)
cudaError_t launch_ega_split(const EgaArgs &a, cudaStream_t stream, int sm_count);

// one translation unit per MASK
template <int MASK>
cudaError_t launch_ega_fast_mask(const EgaArgs &a, cudaStream_t stream, int sm_count, int *ngb_out);
template <int MASK>
cudaError_t launch_ega_tiled_mask(const EgaArgs &a, cudaStream_t stream, int sm_count, int *n_launched); // segment-tiled form (jrb_ega_tiled.cuh)

} // namespace jrb
