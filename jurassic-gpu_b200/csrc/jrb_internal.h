// jrb_internal.h -- kernel argument blocks and launch prototypes shared between the host runtime and the kernels.
#pragma once
#include "jrb_device.cuh"

namespace jrb {

struct TraceArgs {
  long long n_rays;
  // ray geometry SoA: rows 0..6 = obsz, obslon, obslat, vpz, vplon, vplat, time ; row stride geo_stride
  const double *geo;
  long long geo_stride;
  const int *ray_pkg; // [n_rays] package (= atmosphere set) of each ray
  // atmosphere sets, concatenated point arrays
  const long long *pkg_atm_off; // [npk]
  const int *pkg_atm_np;        // [npk]
  const double *atm_time, *atm_z, *atm_lon, *atm_lat, *atm_p, *atm_t;
  const double *atm_q, *atm_k; // [ng][atm_stride], [nw][atm_stride]
  long long atm_stride;
  double *atm_lnp_slope;       // [2][atm_stride] scratch: log(p[i+1]/p[i])/(z[i+1]-z[i]) (NaN where not both positive), dT/dz
  long long n_atm;
  int prepare_atm;             // 1: (re)compute atm_lnp_slope before tracing
  int small_blocks;            // 1: 32/64-thread CTAs that fit beside the resident EGA CTAs (pipelined chunks)
  // 2-D / 3-D atmospheres (ip != 1): geo2cart(0, lon, lat) of every point [3][atm_stride] and the end of its column [atm_stride]
  double *atm_cart;
  int *atm_next;
  // control
  int refrac, ig_h2o;
  double rayds, raydz;
  int ip;                      // atmosphere interpolation (src/jurassic.c:685-691): 1 = 1-D, 2 = 2-D, 3 = 3-D
  double cz, cx;               // influence radii of the 3-D form
  // outputs
  LosLayout los;
  double *los_data; // [n_rays][NLOS][rec] records (written by los_finalize_kernel)
  double *raw;      // [n_rays][NLOS][kRaw] raw points (written by the stepping kernels, see jrb_device.cuh)
  int *ray_np;      // [n_rays]
  int *ray_level0;  // [n_rays] first atmosphere level of the ray's profile (relative to its package)
  double *ray_tsurf;
  double *tp;       // rows tpz, tplon, tplat ; row stride geo_stride
  double *const *tp_host; // optional [3][geo_stride] per-ray addresses in host-mapped memory (the caller's obs_t or the
                          // pinned result buffer): the tangent point is also stored there -- no device-to-host copy phase
  TblDev tbl;       // used when los.fast
  int *error_flag;  // host-mapped words, each one a fatal condition of the reference: [0] a ray needs NLOS or more points ("Too many
                    // LOS points!"), [1] / [2] profile list of the 2-D interpolation ("Cannot identify profiles", "Distance ... too large")
};

struct EgaArgs {
  long long n_rays;
  int ng, nd, nw;
  int ctm_mask;       // CO2*8 + H2O*4 + N2*2 + O2 (fourbit, src/CPUdrivers.c:130-134)
  int ig_co2, ig_h2o;
  int write_bbt;
  int unsorted_columns; // the table set has columns flagged kColNonMonotone -> ROBUST kernel instantiation
  int use_tiled;        // segment-tiled form of the specialised kernel (jrb_ega_tiled.cuh) where it applies
  int los_evict_first;  // line-of-sight record copies carry the L2 evict_first hint (they are streamed, the tables are reused)
  int per_channel_axes; // the (p,T) axes depend on the channel -> PERCH instantiation (lanes locate their own table cells)
  LosLayout los;
  const double *los_data;
  const int *ray_np;
  const double *ray_tsurf;
  const double *chan;   // [CH_NFIELDS][nd]
  const int *window;    // [nd]
  TblDev tbl;
  double *rad, *tau;    // [n_rays][nd]
  // optional per-ray row addresses in host-mapped memory (the caller's registered obs_t rows, or the library's pinned
  // result buffer): every ray's rad/tau row is ALSO stored there by the kernel when the ray finishes, so results land in
  // host memory while the kernel is still running (zero-copy over PCIe) and there is no device-to-host copy phase
  double *const *rad_host, *const *tau_host;
  unsigned long long *work_counter; // dynamic work distribution (zeroed before launch)
  int work_chunk;                   // consecutive items a CTA draws at a time (1..200); 0 = one per warp of the CTA
  // Tail of the work list: the last tail_n items of a launch are handed out longest ray first (tail_perm, built on the device
  // from ray_np by tail_sort_kernel), so that the kernel ends with its shortest rays -- what a warp still has to do after the
  // list ran empty is at most one SHORT ray.  tail_cap = capacity of tail_perm; tail_n is set by the launcher (0 = off).
  int *tail_perm;
  int tail_cap, tail_n;
  unsigned long long *balance;      // [2] scratch (zeroed before launch): idle / total segment slots of lock-step execution
  int phase_lock_mode;              // -1 decide on the device from `balance`, 0 never, 1 always
  int cpw;                          // channels of a ray per warp: 32, or less (= several rays per warp, see jrb_ega_fast.cuh)
  // gas-block passes (split mode, see ega_fast_kernel<SPLIT>): block b holds gases [b*gases_per_block, ...)
  int n_gas_blocks, gases_per_block;
  int blocks_per_group;             // combine: the products of this many consecutive blocks are multiplied first (canonical order)
  double *partial;                  // [n_gas_blocks][n_rays][NLOS][nd] per-segment product of the block's gas factors
  int *partial_len;                 // [n_gas_blocks][n_rays][nd] segments that carry a product (a gas went opaque at len-1 if < np)
  double2 *seg_pre;                 // [n_rays][NLOS][nd] {src * eps, 1 - eps} per segment and channel (ega_segment_kernel)
};

struct FovArgs { // optional epilogue: field-of-view convolution (formod_fov, src/jurassic.c:214-258)
  long long n_rays;
  int nd, n_shape;
  const double *dz, *w;        // [n_shape] FOV shape: altitude offsets [km] and weights
  const int *ray_pkg;          // [n_rays]
  const double *time, *vpz;    // [n_rays]
  const double *rad_in, *tau_in; // [n_rays][nd] pencil-beam results
  double *rad_out, *tau_out;
  int *error;                  // bit 0: a ray with fewer than 2 rays of its own time in its +-NFOV window
};
cudaError_t launch_fov(const FovArgs &a, cudaStream_t stream);

// ---- package I/O (jrb_io.cu) --------------------------------------------------------------------------------------------
// Per-package output addresses in host-mapped memory
struct OutTab { double *rad, *tau, *tpz, *tplon, *tplat; long long stride; };
struct StageArgs {
  int npk;
  int n_geo, n_atm_fields;          // 7 ; 6 + ng + nw
  const long long *ray_off, *atm_off; // [npk+1]
  // direct mode: src[pk*(n_geo+n_atm_fields) + f] = device-visible address of field f of package pk in the caller's
  // registered (page-locked, mapped) atm_t / obs_t; NULL = inputs were packed on the host and copied (staged mode)
  const double *const *src;
  double *geo; long long R;         // [7][R]
  double *atm; long long A;         // [n_atm_fields][A]
  int *ray_pkg;                     // [R]
  int *pkg_atm_np;                  // [npk]
  const OutTab *out;                // [npk]
  double **ray_out;                 // [5][R]: per-ray addresses of the rad row, tau row, tpz, tplon, tplat
};
// fills ray_pkg, pkg_atm_np, the per-ray output addresses and -- in direct mode -- gathers the populated prefixes of the
// packages' arrays straight from the caller's registered host memory into the device SoA (no host-side packing)
cudaError_t launch_stage(const StageArgs &a, cudaStream_t stream);
// copy device-resident results to their host rows (used after the FOV epilogue and by the multi-rank gather)
cudaError_t launch_publish(const double *rad, const double *tau, double *const *rad_host, double *const *tau_host, long long n_rays,
                           int nd, cudaStream_t stream);
cudaError_t launch_nan_mask(double *rad, const long long *flat, long long n, cudaStream_t stream);

// three launches: level slopes, ray stepping (thread per ray), LOS finalisation (thread per ray x segment)
cudaError_t launch_raytrace(const TraceArgs &a, cudaStream_t stream, int *launches);
cudaError_t launch_ega_generic(const EgaArgs &a, cudaStream_t stream);
// fast path: returns cudaErrorInvalidValue if (ng, ctm_mask) has no instantiation
cudaError_t launch_ega_fast(const EgaArgs &a, cudaStream_t stream, int *ngb_out);
// split mode: gas-block passes of the specialised kernel (a.partial / a.partial_len filled), then the combine kernel that
// multiplies the block products per segment and does continuum, Planck source, accumulation and the epilogues
cudaError_t launch_ega_split_passes(const EgaArgs &a, cudaStream_t stream);
cudaError_t launch_ega_segments(const EgaArgs &a, cudaStream_t stream); // everything but the recurrence, per (segment, channel), fully parallel
cudaError_t launch_ega_combine(const EgaArgs &a, cudaStream_t stream);
cudaError_t launch_ega_tiled(const EgaArgs &a, cudaStream_t stream, int *n_launched); // segment-tiled form (jrb_ega_tiled.cuh); kernels it launched
bool ega_tiled_fits(int ng, int los_rec, size_t smem_max);
bool ega_fast_available(int ng, int ctm_mask);
bool ega_fast_fits(int ng, int los_head, int cpw, size_t smem_max);

} // namespace jrb
