// jrb_device.cuh -- device-side data layout and small math helpers shared by the sm_100a kernels.
//
// Layout decisions:
//  * Emissivity tables are re-laid out from tbl_t's [g][p][T][u][d] float arrays (channel innermost, stride ND*4 B
//    between consecutive u entries; src/jurassic.h:408-411) into ONE position-independent blob:
//      - per (gas,channel,p,T) column a contiguous run of "brackets": float4 {u_k, eps_k, u_k+1, eps_k+1}, so the
//        operand set of one linear interpolation (get_u / get_eps, src/jr_common.h:156-185) is a single aligned
//        16-byte load;
//      - column descriptors {first bracket, nu} and the (p,T) axes with the CHANNEL innermost, so the 32 lanes of a
//        warp (= 32 channels of one ray) read them coalesced.
//  * Line-of-sight data is an array of fixed-size records per ray (ray-major), written by the ray tracer and read
//    as warp-uniform broadcasts by the EGA kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jrb {

// ---- constants of the reference (src/jurassic.h:111-129) -------------------------------------------------
constexpr double kC1 = 1.19104259e-8;
constexpr double kC2 = 1.43877506;
constexpr double kP0 = 1013.25;
constexpr double kRE = 6367.421;
constexpr double kBoltzmann = 1.3806504e-23; // GSL_CONST_MKSA_BOLTZMANN (GSL 2.5)
constexpr double kAvogadro = 6.02214199e23;  // GSL_CONST_NUM_AVOGADRO  (GSL 2.5)
constexpr int kNLOS = 400;
constexpr int kTBLNS = 1201;

// ---- packed table blob -------------------------------------------------------------------------------------
struct TblHeader {
  uint64_t magic;   // 'JRBTBL01'
  uint64_t nbytes;  // total blob size
  int32_t ng, nd, npmax, ntmax;
  int32_t all_shared; // 1: every gas has channel-independent (p,T) axes -> fast kernel allowed
  int32_t monotone;   // 1: every column is non-decreasing in u and eps; otherwise the offending columns carry kColNonMonotone
  int32_t max_nu;     // longest column (the specialised kernel packs bracket indices into 10 bits)
  int32_t gas_axes_same; // 1: all gases that have tables share one (p,T) grid -> one table cell per LOS segment
  // byte offsets from blob start
  uint64_t off_np;      // int32  [ng][nd]
  uint64_t off_nt;      // int32  [ng][npmax][nd]
  uint64_t off_pax;     // double [ng][npmax][nd]
  uint64_t off_tax;     // double [ng][npmax][ntmax][nd]
  uint64_t off_col;     // uint2  [ng][npmax][ntmax][nd]  {first bracket index, nu}
  uint64_t off_brk;     // float4 [n_entries]
  uint64_t off_sr;      // double [TBLNS][nd]
  uint64_t off_gnp;     // int32  [ng]                shared axes (valid if all_shared)
  uint64_t off_gnt;     // int32  [ng][npmax]
  uint64_t off_gp;      // double [ng][npmax]
  uint64_t off_gt;      // double [ng][npmax][ntmax]
  uint64_t n_entries;
};
constexpr uint64_t kTblMagic = 0x31304c4254424a52ull; // "RJBTBL01" little endian tag

struct TblDev { // resolved device pointers (built on the host from the header + blob base)
  int ng, nd, npmax, ntmax;
  const int32_t *np;
  const int32_t *nt;
  const double *pax;
  const double *tax;
  const uint2 *col;
  const float4 *brk;
  const double *sr;
  const int32_t *gnp;
  const int32_t *gnt;
  const double *gp;
  const double *gt;
};

// ---- per-channel continuum / channel constants, SoA [field][nd] ---------------------------------------------
enum ChanField {
  CH_NU = 0,
  CH_CO2_296, CH_CO2_260, CH_CO2_230, // co2 continuum coefficients interpolated to nu (src/jr_common.h:319-325)
  CH_H2O_S296,                         // sfac*cw296                                  (:341-352)
  CH_H2O_RATIO,                        // cw260/cw296
  CH_H2O_LNRATIO,                      // log(cw260/cw296)
  CH_H2O_FRN,                          // cwfrn*fscal                                 (:353-357)
  CH_N2_B, CH_N2_BETA,                 // (:368-372)
  CH_O2_B, CH_O2_BETA,                 // (:382-386)
  CH_NFIELDS
};

// ---- LOS record ------------------------------------------------------------------------------------------------
// doubles: [0]p [1]t [2]ds [3]q_h2o [4..4+nw)k  [U0..U0+ng)u  (fast: {wp,wt0,wt1,cell} once, or per gas if the gases'
//          (p,T) grids differ); `rec` = `head` doubles, a multiple of 16 bytes (records are moved by TMA bulk copies).
// The stepping kernels do not write records: a thread that walks a ray would touch a few doubles of one 100-200 byte record
// per step, 64 KB away from its neighbour's -- partial-sector writes scattered over DRAM pages.  They write one compact raw
// point per step instead (kRaw doubles = 64 bytes = two full sectors, consecutive steps adjacent):
//   raw: [0]p [1]t, then the tail [2] altitude z, [3] raw step length, [4] atmosphere level index, [5..8) Cartesian x,y,z
// and los_finalize_kernel, fully parallel and coalesced, turns raw points into records.
struct LosLayout {
  int nw, ng, fast;
  int u0, c0, cstride, head, rec; // offsets in doubles; cell block of gas ig at c0 + cstride*ig (cstride 0: shared)
};
__host__ __device__ inline LosLayout make_los_layout(int ng, int nw, int fast, int shared_cell) {
  LosLayout L;
  L.nw = nw; L.ng = ng; L.fast = fast;
  L.u0 = 4 + nw;
  L.c0 = L.u0 + ng;
  L.cstride = (fast && !shared_cell) ? 4 : 0;
  L.head = L.c0 + (fast ? (shared_cell ? 4 : 4 * ng) : 0);
  L.head = (L.head + 1) & ~1;       // 16-byte multiples
  L.rec = L.head;
  return L;
}
constexpr int kRaw = 8, kRawTail = 2; // doubles per raw point; offset of the tail in it
enum LosTail { LT_Z = 0, LT_DSRAW = 1, LT_LEVEL = 2, LT_X = 3 };
// column descriptor {first bracket, nu}: top bit of nu marks a column that is not monotone in u or eps; its lookups use the
// reference's plain bisection over the whole column instead of the hinted search (whose equivalence needs sorted data)
constexpr unsigned kColNonMonotone = 0x80000000u;
constexpr unsigned kCellInvalid = 0xffffffffu; // "no usable table cell -> gas factor 1"

// ---- math helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double clamp01(double x) { return x > 1.0 ? 1.0 : (x < 0.0 ? 0.0 : x); }

// y0 + (x-x0)*(y1-y0)/(x1-x0)   (lip, src/jr_common.h:48-50) with IEEE division
__device__ __forceinline__ double lerp_div(double x0, double y0, double x1, double y1, double x) {
  return y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}

// Fast reciprocal: MUFU.RCP64H seed + two Newton steps (relative error ~1e-16; no special-case handling --
// callers guarantee a finite, non-zero, normal argument).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// Fast reciprocal square root: MUFU.RSQ64H seed (~2^-20) + two Newton steps (relative error ~1e-16); same caveats as fast_rcp.
// sqrt(x) = x * fast_rsqrt(x) and 1/sqrt(x) come out of ONE dependent chain instead of an IEEE sqrt followed by a reciprocal.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ double lerp_fast(double x0, double y0, double x1, double y1, double x) {
  return fma((x - x0) * (y1 - y0), fast_rcp(x1 - x0), y0);
}

// Reference bisection for ascending data: max{ i <= n-2 : xx[i] <= x }, 0 if x < xx[0]
// (locate_id / locate_tbl_id, src/jr_common.h:106-125)
template <typename F>
__device__ __forceinline__ int bisect_asc(F get, int n, double x, int ilo = 0) {
  int ihi = n - 1;
  while (ihi > ilo + 1) {
    int i = (ihi + ilo) >> 1;
    if (get(i) > x) ihi = i; else ilo = i;
  }
  return ilo;
}

} // namespace jrb
