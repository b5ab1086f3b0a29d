// jrb_runtime.cu -- host runtime behind the C ABI of include/jurassic_b200.h: context, table residency,
// staging of packages into compact device arrays, kernel sequencing, result landing.
//
// Replaces the reference's lane machinery (gpuLane_t / formod_one_package / formod_GPU,
// src/GPUdrivers.cu:176-360).  The reference copies whole atm_t/obs_t structs (2.8 MB + 1.8 MB per 1088 rays, mostly
// unused NP/NR/ND padding) in and out per call.  Here any number of packages is one device batch and there are two I/O
// modes, chosen per batch:
//   staged : the populated prefixes of the packages are packed into one pinned buffer and moved with a single H2D copy;
//            results are stored by the kernels into a pinned, host-mapped result buffer as each ray finishes and are
//            scattered from there into the caller's obs_t rows;
//   direct : when the caller's atm_t / obs_t blocks are page-locked (jrb_host_register), a staging kernel gathers the
//            inputs straight from them over PCIe and the compute kernels store every ray's results straight into the
//            caller's obs_t rows -- no host-side packing, no copy phase, no scatter.
#include "jrb_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <chrono>
#include <thread>
#include <unistd.h>

using namespace jrb;

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf { // page-locked and mapped into the device address space (unified addressing: device address == host address)
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// packed tables of one device; shared by the contexts (lanes) of that device
struct TableSet {
  int device = 0;
  DevBuf blob;
  size_t blob_bytes = 0;
  bool adopted = false;
  TblHeader th;
  TblDev td;
  ~TableSet() { cudaSetDevice(device); blob.release(); }
};

// ---- registry of caller memory page-locked with jrb_host_register (process-wide) -------------------------------------
struct HostRange { uintptr_t lo, hi; char *dev; }; // [lo, hi) page aligned; dev = device address of lo
std::mutex g_reg_mtx;
std::vector<HostRange> g_reg; // sorted by lo, disjoint

// device address of host block [p, p+n) if it lies inside registered memory (seamlessly adjacent ranges are walked), else NULL
void *registered_dev_ptr(const void *p, size_t n) {
  if (g_reg.empty() || !p) return nullptr;
  const uintptr_t a = (uintptr_t)p, b = a + (n ? n : 1);
  size_t lo = 0, hi = g_reg.size();
  while (lo < hi) { const size_t m = (lo + hi) / 2; if (g_reg[m].hi <= a) lo = m + 1; else hi = m; }
  if (lo >= g_reg.size() || g_reg[lo].lo > a) return nullptr;
  char *dev = g_reg[lo].dev + (a - g_reg[lo].lo);
  uintptr_t covered = g_reg[lo].hi;
  for (size_t i = lo; covered < b;) {
    const HostRange &cur = g_reg[i];
    if (++i >= g_reg.size() || g_reg[i].lo != covered || g_reg[i].dev != cur.dev + (cur.hi - cur.lo)) return nullptr;
    covered = g_reg[i].hi;
  }
  return dev;
}

std::string g_create_error;

constexpr int kTailCap = 16384; // rays of a launch that may be handed out in sorted order (two rounds of warps)
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct jrb_context {
  int device = 0;
  int sm_count = 148;
  int l2_bytes = 0, smem_optin = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  std::mutex mtx;

  // control
  bool have_ctl = false;
  int ng = 0, nd = 0, nw = 0, ctm_mask = 0, ig_co2 = -1, ig_h2o = -1, refrac = 1, write_bbt = 0;
  double rayds = 10, raydz = 0.5, hydz = -999;
  int ip = 1;            // atmosphere interpolation: 1-D profiles, 2-D track, 3-D weighted average (src/jurassic.c:685-691)
  double cz = 0, cx = 0; // influence radii of the 3-D form
  std::vector<double> nu;
  std::vector<int> window;
  DevBuf d_chan, d_window;

  // tables (shared between the contexts of a device)
  std::shared_ptr<TableSet> tbl;
  bool have_tbl() const { return tbl && tbl->adopted; }
  int variant_req = -1;
  double los_limit_gb = 0; // 0: JRB_LOS_GB or the default

  // optional field-of-view epilogue (jrb_set_fov)
  int fov_n = 0;
  bool fov_applied = false; // of the last run: results are convolved and the NaN mask is already in them
  DevBuf d_fov, d_fov_out, d_mask, d_flag;

  // staged batch
  bool staged = false, ran = false;
  bool direct = false;          // I/O mode of the staged batch
  int npk = 0;
  long long n_rays = 0, n_atm = 0;
  std::vector<int> pk_nr;
  std::vector<long long> pk_ray_off;
  std::vector<jrb_obs_view> st_obs;                // the obs views of the staged batch (direct mode: where results land)
  std::vector<std::pair<long long, int>> nan_mask; // (global ray, channel) whose input radiance was non-finite
  PinBuf h_in, h_out, h_tab;
  DevBuf d_in, d_tab, d_out, d_rayout, d_los, d_np, d_tsurf, d_counter, d_slope, d_level0, d_raypkg, d_pkgnp;
  // pointers into d_in / d_tab / d_out
  double *geo = nullptr, *atm = nullptr;
  int *ray_pkg = nullptr, *pkg_atm_np = nullptr;
  long long *pkg_atm_off = nullptr;
  double *o_rad = nullptr, *o_tau = nullptr, *o_tp = nullptr;
  double **ray_out = nullptr; // [5][R] per-ray host-mapped addresses: rad row, tau row, tpz, tplon, tplat
  long long chunk_rays = 0;
  int nbuf = 1;                 // LOS buffers: 1 = chunks run back to back, 3 = tracer(c+1) overlaps EGA(c)
  cudaStream_t s_trace = nullptr, s_ega[2] = {nullptr, nullptr};
  LosLayout los;
  int use_fast = 0;
  int cpw = 32;
  int n_gas_blocks = 1, gases_per_block = 1, blocks_per_group = 1; // split mode (jrb_ega_split.cu) when n_gas_blocks > 1
  bool scan_pending = false; // direct mode: the NaN-mask scan of the staged batch is still to be done (before the EGA kernel starts)
  int npk_staging = 0;
  int *err_flag = nullptr;   // word in the pinned table buffer the tracer reports "too many LOS points" through
  int los_evict_first = 0;   // L2 policy (see apply_l2_policy)
  bool l2_window_set = false;
  DevBuf d_partial, d_tail;
  std::vector<cudaEvent_t> events;
  jrb_stats stats;
  bool np_fetched = false;

  int fail(int code, const std::string &m) { err = m; return code; }
  int cuda_fail(cudaError_t e, const char *what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return JRB_ERR_CUDA;
  }
  void invalidate() { staged = false; ran = false; }
};

#define CU(call)                                                \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if (e_ != cudaSuccess) return ctx->cuda_fail(e_, #call);    \
  } while (0)

extern "C" {

const char *jrb_version(void) { return "jurassic-b200 0.2 (sm_100a)"; }

int jrb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char *jrb_last_error(const jrb_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int jrb_create(jrb_context **out, int device) {
  if (!out) return JRB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    g_create_error = std::string("no usable CUDA device (this library has no CPU fallback): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return JRB_ERR_CUDA;
  }
  if (device < 0 || device >= n) { g_create_error = "device ordinal out of range"; return JRB_ERR_ARG; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return JRB_ERR_CUDA; }
  jrb_context *ctx = new jrb_context();
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->l2_bytes, cudaDevAttrL2CacheSize, device);
  cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete ctx; return JRB_ERR_CUDA; }
  int lo_prio = 0, hi_prio = 0;
  cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
  if (cudaStreamCreateWithPriority(&ctx->s_trace, cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->s_ega[0], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->s_ega[1], cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cudaStreamCreate failed"; delete ctx; return JRB_ERR_CUDA;
  }
  std::memset(&ctx->stats, 0, sizeof(ctx->stats));
  *out = ctx;
  return JRB_OK;
}

void jrb_destroy(jrb_context *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto ev : ctx->events) cudaEventDestroy(ev);
  ctx->d_chan.release(); ctx->d_window.release(); ctx->tbl.reset();
  ctx->d_in.release(); ctx->d_tab.release(); ctx->d_out.release(); ctx->d_rayout.release(); ctx->d_los.release();
  ctx->d_np.release(); ctx->d_tsurf.release(); ctx->d_counter.release(); ctx->d_slope.release(); ctx->d_level0.release();
  ctx->d_raypkg.release(); ctx->d_pkgnp.release(); ctx->d_partial.release(); ctx->d_tail.release();
  ctx->h_in.release(); ctx->h_out.release(); ctx->h_tab.release();
  ctx->d_fov.release(); ctx->d_fov_out.release(); ctx->d_mask.release(); ctx->d_flag.release();
  cudaStreamDestroy(ctx->stream);
  if (ctx->s_trace) cudaStreamDestroy(ctx->s_trace);
  for (auto st : ctx->s_ega) if (st) cudaStreamDestroy(st);
  delete ctx;
}

int jrb_context_device(const jrb_context *ctx) { return ctx ? ctx->device : -1; }

int jrb_set_control(jrb_context *ctx, const jrb_ctl_view *c) {
  if (!ctx || !c) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (c->ng < 0 || c->ng > JRB_MAX_NG) return ctx->fail(JRB_ERR_LIMIT, "ng out of range (max 64)");
  if (c->nd < 1) return ctx->fail(JRB_ERR_ARG, "nd must be >= 1");
  if (c->nw < 0 || c->nw > JRB_MAX_NW) return ctx->fail(JRB_ERR_LIMIT, "nw out of range (max 8)");
  if (c->formod != 2) return ctx->fail(JRB_ERR_ARG, "only FORMOD=2 (EGA) is supported (reference asserts the same, src/jr_common.h:707)");
  // 1-D profiles are the reference's formod() path; 2 and 3 are the dispatch of src/jurassic.c:685-691 (see jrb_raytrace.cu)
  if (c->ip < 1 || c->ip > 3) return ctx->fail(JRB_ERR_ARG, "Unknown interpolation method, check IP!"); // src/jurassic.c:690
  if (c->ip == 3 && !(c->cz > 0 && c->cx > 0)) return ctx->fail(JRB_ERR_ARG, "IP=3 needs influence radii CZ > 0 and CX > 0");
  for (int id = 0; id < c->nd; id++)
    if (c->window[id] < 0 || c->window[id] >= (c->nw > 0 ? c->nw : 1)) return ctx->fail(JRB_ERR_ARG, "window index out of range");
  const int mask = ((1 == c->ctm_co2) && (c->ig_co2 >= 0)) * 8 + ((1 == c->ctm_h2o) && (c->ig_h2o >= 0)) * 4 +
                   (1 == c->ctm_n2) * 2 + (1 == c->ctm_o2) * 1; // fourbit of the reference (src/CPUdrivers.c:130-134)
  // unchanged control (the drop-in layer pushes it on every call, like the reference re-uploads ctl_t): nothing to do
  if (ctx->have_ctl && ctx->ng == c->ng && ctx->nd == c->nd && ctx->nw == c->nw && ctx->ig_co2 == c->ig_co2 && ctx->ig_h2o == c->ig_h2o &&
      ctx->ctm_mask == mask && ctx->refrac == c->refrac && ctx->rayds == c->rayds && ctx->raydz == c->raydz && ctx->hydz == c->hydz &&
      ctx->write_bbt == c->write_bbt && ctx->ip == c->ip && ctx->cz == c->cz && ctx->cx == c->cx && std::equal(ctx->nu.begin(), ctx->nu.end(), c->nu) &&
      std::equal(ctx->window.begin(), ctx->window.end(), c->window))
    return JRB_OK;
  CU(cudaSetDevice(ctx->device));
  ctx->invalidate();
  if (ctx->have_ctl && (ctx->ng != c->ng || ctx->nd != c->nd)) ctx->tbl.reset(); // tables must be re-packed
  ctx->ng = c->ng; ctx->nd = c->nd; ctx->nw = c->nw;
  ctx->ig_co2 = c->ig_co2; ctx->ig_h2o = c->ig_h2o;
  ctx->ctm_mask = mask;
  ctx->ip = c->ip; ctx->cz = c->cz; ctx->cx = c->cx;
  ctx->refrac = c->refrac; ctx->rayds = c->rayds; ctx->raydz = c->raydz; ctx->hydz = c->hydz;
  ctx->write_bbt = c->write_bbt;
  ctx->nu.assign(c->nu, c->nu + c->nd);
  ctx->window.assign(c->window, c->window + c->nd);
  std::vector<double> chan;
  channel_constants(ctx->nd, ctx->nu.data(), ctx->ctm_mask, chan);
  CU(ctx->d_chan.ensure(chan.size() * sizeof(double)));
  CU(ctx->d_window.ensure(ctx->nd * sizeof(int)));
  CU(cudaMemcpyAsync(ctx->d_chan.p, chan.data(), chan.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_window.p, ctx->window.data(), ctx->nd * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->have_ctl = true;
  return JRB_OK;
}

static int adopt_blob_locked(jrb_context *ctx) {
  TableSet &ts = *ctx->tbl;
  ts.adopted = false;
  ctx->invalidate();
  unsigned char hdr[sizeof(TblHeader)];
  CU(cudaMemcpy(hdr, ts.blob.p, sizeof(TblHeader), cudaMemcpyDeviceToHost));
  int rc = resolve_tables(hdr, (const unsigned char *)ts.blob.p, ts.th, ts.td, ctx->err);
  if (rc != JRB_OK) return rc;
  if (ts.th.nbytes > ts.blob_bytes) return ctx->fail(JRB_ERR_ARG, "table blob truncated");
  if (ts.th.ng != ctx->ng || ts.th.nd != ctx->nd) return ctx->fail(JRB_ERR_ARG, "table blob was packed for another ng/nd");
  ts.adopted = true;
  ctx->stats.table_blob_bytes = (long long)ts.th.nbytes;
  return JRB_OK;
}

// a fresh table set for this context (other contexts sharing the previous one keep it)
static int new_blob_locked(jrb_context *ctx, size_t nbytes) {
  ctx->invalidate();
  ctx->tbl = std::make_shared<TableSet>();
  ctx->tbl->device = ctx->device;
  CU(ctx->tbl->blob.ensure(nbytes));
  ctx->tbl->blob_bytes = nbytes;
  return JRB_OK;
}

int jrb_set_tables(jrb_context *ctx, const jrb_tbl_view *tbl) {
  if (!ctx || !tbl) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->have_ctl) return ctx->fail(JRB_ERR_STATE, "jrb_set_control must be called before jrb_set_tables");
  CU(cudaSetDevice(ctx->device));
  std::vector<unsigned char> blob;
  int rc = pack_tables(*tbl, ctx->ng, ctx->nd, blob, ctx->err);
  if (rc != JRB_OK) return rc;
  rc = new_blob_locked(ctx, blob.size());
  if (rc != JRB_OK) return rc;
  CU(cudaMemcpy(ctx->tbl->blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  return adopt_blob_locked(ctx);
}

// host-only: pack and report the properties of the packed form (no GPU needed; used by tests of the packer)
int jrb_tables_pack_info(const jrb_tbl_view *tbl, int ng, int nd, size_t *nbytes, int *all_shared, int *monotone,
                         unsigned long long *n_entries, int *gas_axes_same) {
  if (!tbl) return JRB_ERR_ARG;
  std::vector<unsigned char> blob;
  std::string err;
  int rc = pack_tables(*tbl, ng, nd, blob, err);
  if (rc != JRB_OK) { g_create_error = err; return rc; }
  TblHeader h;
  std::memcpy(&h, blob.data(), sizeof(h));
  if (nbytes) *nbytes = blob.size();
  if (all_shared) *all_shared = h.all_shared;
  if (monotone) *monotone = h.monotone;
  if (n_entries) *n_entries = h.n_entries;
  if (gas_axes_same) *gas_axes_same = h.gas_axes_same;
  return JRB_OK;
}

// host-only: the packed blob itself (call with out == NULL to query the size)
int jrb_tables_pack_host(const jrb_tbl_view *tbl, int ng, int nd, void *out, size_t capacity, size_t *nbytes) {
  if (!tbl || !nbytes) return JRB_ERR_ARG;
  std::vector<unsigned char> blob;
  std::string err;
  int rc = pack_tables(*tbl, ng, nd, blob, err);
  if (rc != JRB_OK) { g_create_error = err; return rc; }
  *nbytes = blob.size();
  if (out) {
    if (capacity < blob.size()) { g_create_error = "jrb_tables_pack_host: buffer too small"; return JRB_ERR_ARG; }
    std::memcpy(out, blob.data(), blob.size());
  }
  return JRB_OK;
}

// upload a blob produced by jrb_tables_pack_host (possibly on another rank) and adopt it
int jrb_tables_upload_blob(jrb_context *ctx, const void *host_blob, size_t nbytes) {
  if (!ctx || !host_blob || nbytes < sizeof(TblHeader)) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->have_ctl) return ctx->fail(JRB_ERR_STATE, "jrb_set_control must be called first");
  CU(cudaSetDevice(ctx->device));
  int rc = new_blob_locked(ctx, nbytes);
  if (rc != JRB_OK) return rc;
  CU(cudaMemcpy(ctx->tbl->blob.p, host_blob, nbytes, cudaMemcpyHostToDevice));
  return adopt_blob_locked(ctx);
}

int jrb_tables_blob(jrb_context *ctx, void **dev_ptr, size_t *nbytes) {
  if (!ctx || !dev_ptr || !nbytes) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->have_tbl()) return ctx->fail(JRB_ERR_STATE, "no tables set");
  *dev_ptr = ctx->tbl->blob.p; *nbytes = ctx->tbl->blob_bytes;
  return JRB_OK;
}

int jrb_tables_alloc_blob(jrb_context *ctx, size_t nbytes, void **dev_ptr) {
  if (!ctx || !dev_ptr || nbytes < sizeof(TblHeader)) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  CU(cudaSetDevice(ctx->device));
  int rc = new_blob_locked(ctx, nbytes);
  if (rc != JRB_OK) return rc;
  *dev_ptr = ctx->tbl->blob.p;
  return JRB_OK;
}

int jrb_tables_adopt_blob(jrb_context *ctx) {
  if (!ctx) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->have_ctl) return ctx->fail(JRB_ERR_STATE, "jrb_set_control must be called first");
  if (!ctx->tbl || !ctx->tbl->blob.p) return ctx->fail(JRB_ERR_STATE, "no blob allocated");
  CU(cudaSetDevice(ctx->device));
  return adopt_blob_locked(ctx);
}

// lanes: a second context of the same device uses the tables of the first (no copy, shared ownership)
int jrb_tables_share(jrb_context *ctx, jrb_context *from) {
  if (!ctx || !from || ctx == from) return JRB_ERR_ARG;
  std::shared_ptr<TableSet> t;
  {
    std::lock_guard<std::mutex> lk(from->mtx);
    if (!from->have_tbl()) return from->fail(JRB_ERR_STATE, "no tables set");
    t = from->tbl;
  }
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (ctx->device != t->device) return ctx->fail(JRB_ERR_ARG, "jrb_tables_share: contexts live on different devices");
  if (!ctx->have_ctl) return ctx->fail(JRB_ERR_STATE, "jrb_set_control must be called first");
  if (t->th.ng != ctx->ng || t->th.nd != ctx->nd) return ctx->fail(JRB_ERR_ARG, "tables were packed for another ng/nd");
  ctx->invalidate();
  ctx->tbl = t;
  ctx->stats.table_blob_bytes = (long long)t->th.nbytes;
  return JRB_OK;
}

int jrb_set_kernel_variant(jrb_context *ctx, int variant) {
  if (!ctx || variant < -1 || variant > 1) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  ctx->variant_req = variant;
  ctx->invalidate();
  return JRB_OK;
}

int jrb_set_los_limit_gb(jrb_context *ctx, double gb) {
  if (!ctx || gb < 0) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  ctx->los_limit_gb = gb;
  return JRB_OK;
}

// FOV shape (what read_shape returns for ctl->fov, src/jurassic.c:222): n = 0 switches the epilogue off
int jrb_set_fov(jrb_context *ctx, int n, const double *dz, const double *w) {
  if (!ctx || n < 0 || (n > 0 && (!dz || !w))) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (n > 2048) return ctx->fail(JRB_ERR_LIMIT, "Too many data points!"); // NSHAPE (src/jurassic.h:172, src/jurassic.c:1143)
  CU(cudaSetDevice(ctx->device));
  if (n > 0) {
    CU(ctx->d_fov.ensure((size_t)2 * n * 8));
    CU(cudaMemcpyAsync(ctx->d_fov.p, dz, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync((double *)ctx->d_fov.p + n, w, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  ctx->fov_n = n;
  return JRB_OK;
}

// ---- page-locking of caller memory ---------------------------------------------------------------------------------------
// Registers the pages of [ptr, ptr+bytes) that are not registered yet (blocks may share pages or overlap earlier calls).
int jrb_host_register(void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return JRB_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); g_create_error = "jrb_host_register: no CUDA device"; return JRB_ERR_CUDA; }
  const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
  uintptr_t lo = (uintptr_t)ptr / page * page, hi = ((uintptr_t)ptr + bytes + page - 1) / page * page;
  std::lock_guard<std::mutex> lk(g_reg_mtx);
  // subtract what is registered already; register the remaining pieces
  std::vector<std::pair<uintptr_t, uintptr_t>> todo;
  uintptr_t cur = lo;
  for (const HostRange &r : g_reg) {
    if (r.hi <= cur) continue;
    if (r.lo >= hi) break;
    if (r.lo > cur) todo.push_back({cur, r.lo});
    cur = std::max(cur, r.hi);
    if (cur >= hi) break;
  }
  if (cur < hi) todo.push_back({cur, hi});
  for (auto &t : todo) {
    cudaError_t e = cudaHostRegister((void *)t.first, t.second - t.first, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
      cudaGetLastError();
      g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e);
      return JRB_ERR_CUDA;
    }
    void *dev = nullptr;
    e = cudaHostGetDevicePointer(&dev, (void *)t.first, 0);
    if (e != cudaSuccess) { cudaGetLastError(); cudaHostUnregister((void *)t.first); g_create_error = std::string("cudaHostGetDevicePointer: ") + cudaGetErrorString(e); return JRB_ERR_CUDA; }
    HostRange r{t.first, t.second, (char *)dev};
    g_reg.insert(std::upper_bound(g_reg.begin(), g_reg.end(), r, [](const HostRange &a, const HostRange &b) { return a.lo < b.lo; }), r);
  }
  return JRB_OK;
}

// releases the registered ranges that lie completely inside [ptr, ptr+bytes) (page granular)
int jrb_host_unregister(void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return JRB_ERR_ARG;
  const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
  const uintptr_t lo = (uintptr_t)ptr / page * page, hi = ((uintptr_t)ptr + bytes + page - 1) / page * page;
  std::lock_guard<std::mutex> lk(g_reg_mtx);
  for (size_t i = 0; i < g_reg.size();) {
    if (g_reg[i].lo >= lo && g_reg[i].hi <= hi) {
      if (cudaHostUnregister((void *)g_reg[i].lo) != cudaSuccess) cudaGetLastError();
      g_reg.erase(g_reg.begin() + (long)i);
    } else ++i;
  }
  return JRB_OK;
}

int jrb_host_unregister_all(void) {
  std::lock_guard<std::mutex> lk(g_reg_mtx);
  for (const HostRange &r : g_reg) if (cudaHostUnregister((void *)r.lo) != cudaSuccess) cudaGetLastError();
  g_reg.clear();
  return JRB_OK;
}

// 1 if [ptr, ptr+bytes) lies in memory registered with jrb_host_register
int jrb_host_is_registered(const void *ptr, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_reg_mtx);
  return registered_dev_ptr(ptr, bytes) != nullptr;
}

// ---- hydrostatic adjustment on the host (hydrostatic_1d_h2o, src/jr_common.h:714-761; default off) -----------
static void hydrostatic_host(const jrb_atm_view &a, double hydz, int ig_h2o) {
  const int npts = 20, ip0 = 0, ip1 = a.np;
  double dzmin = 1e99;
  int ipref = 0;
  for (int ip = ip0; ip < ip1; ip++) {
    const double dz = std::fabs(a.z[ip] - hydz);
    if (dz < dzmin) { dzmin = dz; ipref = ip; }
  }
  const double lat = a.lat[ipref];
  const double mmair = 28.96456e-3, mmh2o = 18.0153e-3, rgas = 8.314472; // GSL_CONST_MKSA_MOLAR_GAS (GSL 2.5)
  auto gravity = [](double z, double la) {
    const double deg2rad = M_PI / 180., x = std::sin(la * deg2rad), y = std::sin(2 * la * deg2rad);
    return 9.780318 * (1. + 0.0053024 * x * x - 5.8e-6 * y * y) - 3.086e-3 * z;
  };
  auto lin = [](double x0, double y0, double x1, double y1, double x) { return y0 + (x - x0) * (y1 - y0) / (x1 - x0); };
  const double *qh = (ig_h2o >= 0) ? (a.q_rows ? a.q_rows[ig_h2o] : a.q + (size_t)ig_h2o * a.q_stride) : nullptr;
  double e = 0.;
  for (int dir = 0; dir < 2; dir++) {
    const int step = dir == 0 ? 1 : -1;
    for (int ip = ipref + step; dir == 0 ? ip < ip1 : ip >= ip0; ip += step) {
      const int prev = ip - step;
      double mean = 0.;
      for (int i = 0; i < npts; i++) {
        const double z = lin(0.0, a.z[prev], npts - 1.0, a.z[ip], (double)i);
        const double grav = gravity(z, lat);
        if (qh) e = lin(0.0, qh[prev], npts - 1.0, qh[ip], (double)i);
        const double temp = lin(0.0, a.t[prev], npts - 1.0, a.t[ip], (double)i);
        mean += (e * mmh2o + (1 - e) * mmair) * grav / (rgas * temp * npts);
      }
      a.p[ip] = a.p[prev] * std::exp(-1000 * mean * (a.z[ip] - a.z[prev]));
    }
  }
}

// channels per warp of the specialised kernel: all of them for few-channel instruments; otherwise 32 unless the tables of
// (32 channels x ng gases) cannot stay in L2 -- then narrower channel groups with several rays per warp (channel-group-major
// order keeps the hot set at cpw x ng pairs).  Hot bytes per (gas, channel) pair: ~30 % of its brackets (measured on the
// synthetic sets: a package touches 357 KB of a 1.26 MB pair); budget 0.7 x L2, which reproduces the measured optima
// (segment-by-segment kernel -- Config D, 57 MB at 32 channels: 32 = 16 > 8; Config E, 94 MB: 16 best, -4 %; 30 gases, 343 MB: 8 best, -14 %).
static int choose_cpw(const jrb_context *ctx, int gases_per_pass) {
  const int nd = ctx->nd, ng = ctx->ng;
  int cpw = nd <= 16 ? nd : 32;
  if (nd > 16 && ng > 0) {
    const double pair_hot = 0.3 * 16.0 * (double)ctx->tbl->th.n_entries / ((double)ng * nd);
    // the segment-tiled kernel (32-channel groups only) re-reads a bracket half as often, so it tolerates a larger hot set:
    // Config E (94 MB at 32 channels) runs 7 % faster tiled at 32 than segment by segment at 16; at 121 MB (a 10-gas pass of
    // the 30-gas refspec shape) the narrower groups win again (82.7 vs 98.7 ms)
    const bool tiled_ok = ctx->tbl->th.all_shared && !(getenv("JRB_EGA_TILED") && atoi(getenv("JRB_EGA_TILED")) == 0) &&
                          ega_tiled_fits(gases_per_pass, make_los_layout(ng, ctx->nw, 1, ctx->tbl->th.gas_axes_same).rec, (size_t)ctx->smem_optin);
    if (!(tiled_ok && pair_hot * 32 * gases_per_pass <= 0.8 * (double)ctx->l2_bytes)) // (E: 94 MB yes; 10 of 30 refspec gases: 121 MB no)
      while (cpw > 4 && pair_hot * cpw * gases_per_pass > 0.7 * (double)ctx->l2_bytes) cpw >>= 1;
  }
  if (const char *s = getenv("JRB_EGA_CPW")) { const int v = atoi(s); if (v >= 1 && v <= 32 && (v == nd || (32 % v == 0 && v <= nd))) cpw = v; } // experiments
  return cpw;
}

// save_mask (src/jr_common.h:193-200) for one package: the (ray, channel) pairs whose input radiance is not finite; the
// columns [nd, nd_reset) are reset here in direct mode (the kernels only write the first nd)
static void scan_package(const jrb_obs_view &o, long long r0, int nd, bool reset_tail, std::vector<std::pair<long long, int>> &mask) {
  for (int ir = 0; ir < o.nr; ir++) {
    const double *row = o.rad + (size_t)ir * o.row_stride;
    // a row is finite iff its sum of (x - x) is 0; only rows that fail this cheap test are inspected element-wise
    double probe = 0.0;
    for (int id = 0; id < nd; id++) probe += row[id] - row[id];
    if (probe != 0.0 || probe != probe)
      for (int id = 0; id < nd; id++)
        if (!std::isfinite(row[id])) mask.push_back({r0 + ir, id});
    if (reset_tail) {
      double *rr = o.rad + (size_t)ir * o.row_stride, *tt = o.tau + (size_t)ir * o.row_stride;
      for (int id = nd; id < o.nd_reset; id++) { rr[id] = 0.0; tt[id] = 1.0; } // all ND columns are reset (src/CPUdrivers.c:58-60)
    }
  }
}

// direct mode: NaN-mask scan and reset of the columns beyond nd over the staged obs views (touches only the caller's rad/tau
// rows, which the device does not read and the EGA kernel has not started to write)
static void scan_staged_obs(jrb_context *ctx) {
  const int npk = ctx->npk_staging, nd = ctx->nd;
  const double t0 = now_ms();
  std::vector<std::vector<std::pair<long long, int>>> masks(npk);
#pragma omp parallel for schedule(dynamic, 4) num_threads(host_threads())
  for (int k = 0; k < npk; k++) scan_package(ctx->st_obs[k], ctx->pk_ray_off[k], nd, ctx->st_obs[k].nd_reset > nd, masks[k]);
  ctx->nan_mask.clear();
  for (int k = 0; k < npk; k++) ctx->nan_mask.insert(ctx->nan_mask.end(), masks[k].begin(), masks[k].end());
  ctx->scan_pending = false;
  ctx->stats.host_ms_pack = (float)(now_ms() - t0);
}

static int stage_locked(jrb_context *ctx, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs, bool defer_scan = false) {
  ctx->invalidate(); // whatever was staged before is gone, whether or not this call succeeds
  if (!ctx->have_ctl || !ctx->have_tbl()) return ctx->fail(JRB_ERR_STATE, "control and tables must be set before staging");
  CU(cudaSetDevice(ctx->device));
  const int ng = ctx->ng, nd = ctx->nd, nw = ctx->nw;
  const TblHeader &th = ctx->tbl->th;
  const double t_begin = now_ms();

  long long R = 0, A = 0;
  for (int k = 0; k < npk; k++) {
    if (obs[k].nr < 0 || atm[k].np < 1) return ctx->fail(JRB_ERR_ARG, "package with nr < 0 or empty atmosphere");
    R += obs[k].nr; A += atm[k].np;
  }

  // ---- kernel choice (before anything is allocated) ----
  // Split mode (gas-block passes + combine kernel, jrb_ega_split.cu).  The ORDER in which the gas factors of a segment are
  // multiplied is a function of ng alone, so that results do not depend on how a batch is cut (packages per call, devices):
  // up to 12 gases one running product in gas order (what the fused kernel and the reference do); above, groups of at most
  // 10 gases whose products are multiplied in group order.  What varies with the batch is only how the work is cut:
  //  (a) many gases: one pass per group, so that 24 warps per SM keep their per-gas state (16 B per gas and thread) in
  //      shared memory and the tables of one pass fit the L2;
  //  (b) small batches -- above all the single 1088-ray package of an unmodified formod() caller: one gas per pass item, which
  //      gives the GPU ng times more independent warps; the combine kernel multiplies the factors in the canonical order.
  const int gcan = ng > 12 ? (ng + (ng + 9) / 10 - 1) / ((ng + 9) / 10) : (ng > 0 ? ng : 1); // gases per product group
  int gpb = gcan;                                                                             // gases per pass item
  const char *pipe_env = getenv("JRB_PIPELINE");
  const bool want_pipe = pipe_env && atoi(pipe_env) != 0 && R >= 32768; // (experimental chunk pipeline: fused kernel only)
  int bpg = 1; // pass blocks per product group
  if (ng > 1 && !getenv("JRB_NO_SPLIT") && !want_pipe) {
    const int cpw0 = nd <= 16 ? nd : 32, rpw0 = 32 / cpw0;
    const long long items = ((R + rpw0 - 1) / rpw0) * ((nd + cpw0 - 1) / cpw0);
    const long long slots = (long long)ctx->sm_count * 24;
    if (items > 0 && items * ((ng + gcan - 1) / gcan) < 2 * slots) { gpb = 1; bpg = gcan; }
  } else {
    gpb = ng > 0 ? ng : 1; // fused kernel (for more than 12 gases its product order differs from the canonical one)
  }
  if (const char *e = getenv("JRB_EGA_GAS_BLOCK")) { const int v = atoi(e); if (v >= 1 && v <= ng) { gpb = v; bpg = 1; } } // experiments
  const int nblk = ng > 0 ? (ng + gpb - 1) / gpb : 1;
  const int cpw = choose_cpw(ctx, nblk > 1 ? gpb : ng);
  // channel-dependent (p,T) axes: the specialised kernel locates the table cell per lane (PERCH) and the records carry no cell
  const LosLayout los_fast = make_los_layout(ng, nw, th.all_shared ? 1 : 0, th.gas_axes_same);
  const bool fits = ega_fast_fits(nblk > 1 ? gpb : ng, los_fast.head, cpw, (size_t)ctx->smem_optin);
  const bool fast_ok = th.max_nu <= 1023 && ega_fast_available(ng, ctx->ctm_mask) && fits;
  if (ctx->variant_req == 1 && !fast_ok)
    return ctx->fail(JRB_ERR_STATE, std::string("specialised kernel not applicable: shared_axes=") + std::to_string(th.all_shared) +
                     " monotone=" + std::to_string(th.monotone) + " max_nu=" + std::to_string(th.max_nu) + " ng=" + std::to_string(ng) + " mask=" + std::to_string(ctx->ctm_mask) +
                     " built=" + std::to_string((int)ega_fast_available(ng, ctx->ctm_mask)) + " fits_shared_memory=" + std::to_string((int)fits));
  const int use_fast = (ctx->variant_req == 0) ? 0 : (fast_ok ? 1 : 0);

  // ---- package tables (offsets, output addresses, input addresses) in one small pinned buffer ----
  const int n_geo = 7, n_af = 6 + ng + nw, n_src = n_geo + n_af;
  const size_t off_roff = 0, off_aoff = off_roff + (size_t)(npk + 1) * 8, off_out = off_aoff + (size_t)(npk + 1) * 8;
  const size_t off_src = off_out + (size_t)npk * sizeof(OutTab);
  const size_t off_flag = align_up(off_src + (size_t)npk * n_src * 8, 256); // error word written by the tracer (host-mapped)
  const size_t tab_bytes = off_flag + 256;
  CU(ctx->h_tab.ensure(tab_bytes));
  CU(ctx->d_tab.ensure(tab_bytes));
  unsigned char *HT = (unsigned char *)ctx->h_tab.p;
  long long *h_roff = (long long *)(HT + off_roff), *h_aoff = (long long *)(HT + off_aoff);
  OutTab *h_outtab = (OutTab *)(HT + off_out);
  const double **h_src = (const double **)(HT + off_src);

  ctx->err_flag = (int *)(HT + off_flag);
  ctx->err_flag[0] = ctx->err_flag[1] = ctx->err_flag[2] = 0;
  ctx->pk_nr.resize(npk);
  ctx->pk_ray_off.resize(npk + 1);
  {
    long long r = 0, a = 0;
    for (int k = 0; k < npk; k++) {
      ctx->pk_nr[k] = obs[k].nr;
      ctx->pk_ray_off[k] = r; h_roff[k] = r; h_aoff[k] = a;
      r += obs[k].nr; a += atm[k].np;
    }
    ctx->pk_ray_off[npk] = r; h_roff[npk] = r; h_aoff[npk] = a;
  }

  // ---- I/O mode: direct if every array of every package lies in registered (page-locked, mapped) memory ----
  bool direct = npk > 0 && !getenv("JRB_NO_DIRECT_IO");
  if (direct) {
    std::lock_guard<std::mutex> rl(g_reg_mtx);
    if (g_reg.empty()) direct = false;
    for (int k = 0; k < npk && direct; k++) {
      const jrb_obs_view &o = obs[k];
      const jrb_atm_view &a = atm[k];
      const size_t nrb = (size_t)o.nr * 8, npb = (size_t)a.np * 8;
      const double **s = h_src + (size_t)k * n_src;
      const double *geo_in[7] = {o.obsz, o.obslon, o.obslat, o.vpz, o.vplon, o.vplat, o.time};
      for (int f = 0; f < 7; f++) s[f] = (const double *)registered_dev_ptr(geo_in[f], nrb);
      const double *atm_in[6] = {a.time, a.z, a.lon, a.lat, a.p, a.t};
      for (int f = 0; f < 6; f++) s[7 + f] = (const double *)registered_dev_ptr(atm_in[f], npb);
      for (int ig = 0; ig < ng; ig++)
        s[13 + ig] = (const double *)registered_dev_ptr(a.q_rows ? a.q_rows[ig] : a.q + (size_t)ig * a.q_stride, npb);
      for (int iw = 0; iw < nw; iw++)
        s[13 + ng + iw] = (const double *)registered_dev_ptr(a.k_rows ? a.k_rows[iw] : a.k + (size_t)iw * a.k_stride, npb);
      for (int f = 0; f < n_src; f++) if (!s[f]) direct = false;
      OutTab &t = h_outtab[k];
      const size_t rows = o.nr > 0 ? ((size_t)(o.nr - 1) * o.row_stride + nd) * 8 : 8;
      t.rad = (double *)registered_dev_ptr(o.rad, rows); t.tau = (double *)registered_dev_ptr(o.tau, rows);
      t.tpz = (double *)registered_dev_ptr(o.tpz, nrb); t.tplon = (double *)registered_dev_ptr(o.tplon, nrb);
      t.tplat = (double *)registered_dev_ptr(o.tplat, nrb);
      t.stride = o.row_stride;
      if (!t.rad || !t.tau || !t.tpz || !t.tplon || !t.tplat) direct = false;
    }
  }

  if (ctx->hydz >= 0) // modifies the caller's atm->p like the reference's CPU path (src/CPUdrivers.c:97-103)
    for (int k = 0; k < npk; k++) hydrostatic_host(atm[k], ctx->hydz, ctx->ig_h2o);

  // ---- device arrays ----
  const size_t n_geo_d = 7 * (size_t)R, n_atm_d = (size_t)n_af * A;
  const size_t off_geo = 0, off_atm = off_geo + n_geo_d * 8;
  const size_t in_bytes = align_up(off_atm + n_atm_d * 8, 256) + 256;
  CU(ctx->d_in.ensure(in_bytes));
  unsigned char *Dv = (unsigned char *)ctx->d_in.p;
  ctx->geo = (double *)(Dv + off_geo); ctx->atm = (double *)(Dv + off_atm);
  CU(ctx->d_raypkg.ensure((size_t)(R ? R : 1) * 4));
  CU(ctx->d_pkgnp.ensure((size_t)(npk ? npk : 1) * 4));
  ctx->ray_pkg = (int *)ctx->d_raypkg.p; ctx->pkg_atm_np = (int *)ctx->d_pkgnp.p;
  unsigned char *DT = (unsigned char *)ctx->d_tab.p;
  ctx->pkg_atm_off = (long long *)(DT + off_aoff);
  const size_t out_bytes = ((size_t)2 * R * nd + 3 * (size_t)R) * 8 + 256;
  CU(ctx->d_out.ensure(out_bytes));
  ctx->o_rad = (double *)ctx->d_out.p; ctx->o_tau = ctx->o_rad + (size_t)R * nd; ctx->o_tp = ctx->o_tau + (size_t)R * nd;
  CU(ctx->d_rayout.ensure((size_t)5 * (R ? R : 1) * 8));
  ctx->ray_out = (double **)ctx->d_rayout.p;
  CU(ctx->d_np.ensure((size_t)(R ? R : 1) * 4));
  CU(ctx->d_tsurf.ensure((size_t)(R ? R : 1) * 8));
  // level slopes [2][A]; 2-D / 3-D atmospheres: + Cartesian positions [3][A] and column ends [A]
  CU(ctx->d_slope.ensure((size_t)(A ? A : 1) * (ctx->ip != 1 ? 48 : 16)));
  CU(ctx->d_level0.ensure((size_t)(R ? R : 1) * 4));

  double t_pack0 = now_ms(), t_pack1 = t_pack0;
  ctx->nan_mask.clear();
  std::vector<std::vector<std::pair<long long, int>>> masks(npk);
  size_t h2d_bytes = tab_bytes;
  if (!direct) {
    // staged mode: pack the populated prefixes into the pinned buffer; results go to the pinned result buffer
    CU(ctx->h_in.ensure(in_bytes));
    CU(ctx->h_out.ensure(out_bytes));
    unsigned char *H = (unsigned char *)ctx->h_in.p;
    double *hgeo = (double *)(H + off_geo), *hatm = (double *)(H + off_atm);
    double *hrad = (double *)ctx->h_out.p, *htau = hrad + (size_t)R * nd, *htp = htau + (size_t)R * nd;
#pragma omp parallel for schedule(dynamic, 4) num_threads(host_threads())
    for (int k = 0; k < npk; k++) {
      const jrb_obs_view &o = obs[k];
      const jrb_atm_view &a = atm[k];
      const long long r0 = ctx->pk_ray_off[k], a0 = h_aoff[k];
      const size_t nrb = (size_t)o.nr * 8, npb = (size_t)a.np * 8;
      std::memcpy(hgeo + 0 * R + r0, o.obsz, nrb);
      std::memcpy(hgeo + 1 * R + r0, o.obslon, nrb);
      std::memcpy(hgeo + 2 * R + r0, o.obslat, nrb);
      std::memcpy(hgeo + 3 * R + r0, o.vpz, nrb);
      std::memcpy(hgeo + 4 * R + r0, o.vplon, nrb);
      std::memcpy(hgeo + 5 * R + r0, o.vplat, nrb);
      std::memcpy(hgeo + 6 * R + r0, o.time, nrb);
      std::memcpy(hatm + 0 * A + a0, a.time, npb);
      std::memcpy(hatm + 1 * A + a0, a.z, npb);
      std::memcpy(hatm + 2 * A + a0, a.lon, npb);
      std::memcpy(hatm + 3 * A + a0, a.lat, npb);
      std::memcpy(hatm + 4 * A + a0, a.p, npb);
      std::memcpy(hatm + 5 * A + a0, a.t, npb);
      for (int ig = 0; ig < ng; ig++)
        std::memcpy(hatm + (size_t)(6 + ig) * A + a0, a.q_rows ? a.q_rows[ig] : a.q + (size_t)ig * a.q_stride, npb);
      for (int iw = 0; iw < nw; iw++)
        std::memcpy(hatm + (size_t)(6 + ng + iw) * A + a0, a.k_rows ? a.k_rows[iw] : a.k + (size_t)iw * a.k_stride, npb);
      scan_package(o, r0, nd, false, masks[k]);
      OutTab &t = h_outtab[k]; // device address == host address for cudaHostAlloc'ed mapped memory (unified addressing)
      t.rad = hrad + (size_t)r0 * nd; t.tau = htau + (size_t)r0 * nd;
      t.tpz = htp + 0 * R + r0; t.tplon = htp + 1 * R + r0; t.tplat = htp + 2 * R + r0;
      t.stride = nd;
    }
    t_pack1 = now_ms();
    CU(cudaMemcpyAsync(ctx->d_in.p, H, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    h2d_bytes += in_bytes;
  }
  CU(cudaMemcpyAsync(ctx->d_tab.p, HT, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
  {
    StageArgs sa;
    sa.npk = npk; sa.n_geo = n_geo; sa.n_atm_fields = n_af;
    sa.ray_off = (const long long *)(DT + off_roff); sa.atm_off = (const long long *)(DT + off_aoff);
    sa.src = direct ? (const double *const *)(DT + off_src) : nullptr;
    sa.geo = ctx->geo; sa.R = R; sa.atm = ctx->atm; sa.A = A;
    sa.ray_pkg = ctx->ray_pkg; sa.pkg_atm_np = ctx->pkg_atm_np;
    sa.out = (const OutTab *)(DT + off_out); sa.ray_out = ctx->ray_out;
    CU(launch_stage(sa, ctx->stream));
  }
  ctx->st_obs.assign(obs, obs + npk);
  ctx->npk_staging = npk;
  ctx->scan_pending = false;
  for (int k = 0; k < npk; k++) ctx->nan_mask.insert(ctx->nan_mask.end(), masks[k].begin(), masks[k].end());
  if (direct) {
    // the staging kernel is pulling the inputs over PCIe meanwhile.  The NaN-mask scan (and the reset of the columns beyond
    // nd) only has to be done before the EGA kernel starts to store results: a complete call (jrb_formod_batch) does it
    // while the ray tracer runs, the split API does it here
    if (defer_scan) ctx->scan_pending = true;
    else scan_staged_obs(ctx);
    t_pack1 = now_ms();
    long long in_b = 0;
    for (int k = 0; k < npk; k++) in_b += (long long)obs[k].nr * 8 * 7 + (long long)atm[k].np * 8 * n_af;
    h2d_bytes += (size_t)in_b; // read by the staging kernel from the caller's memory
  }

  // ---- LOS buffer ----
  ctx->use_fast = use_fast;
  ctx->cpw = cpw;
  ctx->los = make_los_layout(ng, nw, use_fast && th.all_shared, th.gas_axes_same);
  ctx->n_gas_blocks = use_fast ? nblk : 1;
  ctx->gases_per_block = gpb;
  ctx->blocks_per_group = bpg;
  // scratch per ray: the line-of-sight records, plus the per-segment block products in split mode
  const size_t part_per_ray = ctx->n_gas_blocks > 1 ? (size_t)ctx->n_gas_blocks * ((size_t)kNLOS * nd * 8 + (size_t)nd * 4) + (size_t)kNLOS * nd * 16 : 0;
  // (records + the raw points the stepping kernels write, jrb_device.cuh)
  const size_t per_ray = (size_t)kNLOS * (ctx->los.rec + kRaw) * 8 + part_per_ray;
  // LOS scratch: JRB_LOS_GB / jrb_set_los_limit_gb (default 72 GB: the 1 000 960 rays of BASELINE's config D need 64 GB),
  // but never more than half of what is free on the device.  The driver is only asked (cudaMemGetInfo takes a
  // device-wide lock and was seen to stall for tens of ms) when the buffer has to grow.
  double los_gb = 72.0;
  if (const char *s = getenv("JRB_LOS_GB")) { double v = atof(s); if (v > 0.01) los_gb = v; }
  if (ctx->los_limit_gb > 0) los_gb = ctx->los_limit_gb;
  if ((double)R * (double)(per_ray - part_per_ray) > (double)ctx->d_los.cap || (double)R * (double)part_per_ray > (double)ctx->d_partial.cap) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const double avail = 0.5 * ((double)free_b + (double)ctx->d_los.cap + (double)ctx->d_partial.cap) / 1e9;
      if (avail < los_gb) los_gb = avail;
    }
  } else if ((double)(ctx->d_los.cap + ctx->d_partial.cap) / 1e9 < los_gb) {
    los_gb = (double)R * (double)per_ray / 1e9 + 1e-9; // the existing buffers are enough for the whole batch
  }
  // Chunking.  By default a batch is one chunk (or as many as the LOS scratch limit requires), run back to back.
  // JRB_PIPELINE=1 cuts large batches into ~8 chunks held in 3 rotating LOS buffers so that the (latency-bound, low
  // occupancy) tracer kernels of chunk c+1 run beside the EGA kernel of chunk c.  Measured on B200 (profiles/README.md):
  // the extra EGA kernel tails cost more than the hidden tracer time (143.4 vs 138.3 ms), hence opt-in.
  long long chunk = (long long)(los_gb * 1e9 / (double)per_ray);
  ctx->nbuf = 1;
  if (want_pipe) {
    long long pc = (R + 7) / 8;
    if (pc < 8192) pc = 8192;
    if (pc * 3 > chunk) pc = chunk / 3;
    if (pc >= 4096) { chunk = pc; ctx->nbuf = 3; }
  }
  if (chunk < 1024) chunk = 1024;
  if (chunk > R) chunk = R;
  // equal chunks: the last one is not a short straggler
  if (chunk > 0 && ctx->nbuf == 1) { const long long nch = (R + chunk - 1) / chunk; chunk = (R + nch - 1) / nch; }
  ctx->chunk_rays = chunk;
  CU(ctx->d_los.ensure((size_t)(chunk ? chunk : 1) * (per_ray - part_per_ray) * ctx->nbuf));
  if (part_per_ray) {
    if (ctx->nbuf > 1) return ctx->fail(JRB_ERR_STATE, "JRB_PIPELINE cannot be combined with split mode");
    CU(ctx->d_partial.ensure((size_t)(chunk ? chunk : 1) * part_per_ray + 256));
  }
  {
    const long long nch = R > 0 ? (R + chunk - 1) / chunk : 1;
    CU(ctx->d_counter.ensure((size_t)nch * 32 + 256)); // per LOS chunk: work counter, lock-step balance (idle, total), pad
    CU(ctx->d_tail.ensure((size_t)nch * kTailCap * 4));  // per LOS chunk: order of the last rays of the launch (longest first)
  }

  CU(cudaStreamSynchronize(ctx->stream));
  ctx->npk = npk; ctx->n_rays = R; ctx->n_atm = A;
  ctx->direct = direct;
  if (!direct) ctx->stats.host_ms_pack = (float)(t_pack1 - t_pack0);
  ctx->stats.host_ms_h2d = (float)(now_ms() - t_pack1);
  ctx->stats.host_ms_stage = (float)(now_ms() - t_begin);
  ctx->stats.h2d_bytes = (long long)h2d_bytes;
  ctx->stats.io_direct = direct ? 1 : 0;
  ctx->stats.n_packages = npk; ctx->stats.n_rays = R; ctx->stats.n_ray_channels = R * nd;
  ctx->stats.cum_runs = ctx->stats.cum_launches = ctx->stats.cum_ega_launches = 0;
  ctx->stats.cum_ms_ega = ctx->stats.cum_ms_raytrace = ctx->stats.cum_ms_device = 0.0;
  ctx->staged = true; ctx->np_fetched = false;
  return JRB_OK;
}

int jrb_stage(jrb_context *ctx, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs) {
  if (!ctx || npk < 0 || (npk > 0 && (!atm || !obs))) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  return stage_locked(ctx, npk, atm, obs);
}

// L2 residency (north star: "pinned in L2 via access-policy windows").  The table brackets are what every warp keeps coming
// back to (98 % of the gathers hit the L2), the line-of-sight records are streamed through once per channel group:
//   JRB_L2_PERSIST=1     : persisting-L2 carve-out + access-policy window over the bracket array on the compute stream
//   JRB_LOS_EVICT_FIRST=1: the TMA copies of the records carry the L2::evict_first hint
// Both are measured in profiles/ and default to what measured best.
static void apply_l2_policy(jrb_context *ctx) {
  const char *ef = getenv("JRB_LOS_EVICT_FIRST");
  ctx->los_evict_first = ef ? (atoi(ef) != 0) : 0;
  const char *ps = getenv("JRB_L2_PERSIST");
  const bool want = ps ? (atoi(ps) != 0) : false;
  if (want == ctx->l2_window_set && !want) return;
  cudaStreamAttrValue attr;
  std::memset(&attr, 0, sizeof(attr));
  if (want) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    if (max_persist <= 0 || max_window <= 0) return;
    double frac = 0.75;
    if (const char *f = getenv("JRB_L2_PERSIST_FRACTION")) { const double v = atof(f); if (v > 0 && v <= 1) frac = v; }
    const size_t carve = (size_t)((double)max_persist * frac);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    const size_t brk_bytes = (size_t)ctx->tbl->th.n_entries * 16;
    const size_t win = std::min(brk_bytes, (size_t)max_window);
    attr.accessPolicyWindow.base_ptr = (void *)ctx->tbl->td.brk;
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)win);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;
  }
  if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
  ctx->l2_window_set = want;
}

static int run_locked(jrb_context *ctx) {
  if (!ctx->staged) return ctx->fail(JRB_ERR_STATE, "nothing staged");
  ctx->ran = false;
  CU(cudaSetDevice(ctx->device));
  apply_l2_policy(ctx);
  const long long R = ctx->n_rays, A = ctx->n_atm;
  const int ng = ctx->ng, nd = ctx->nd, nw = ctx->nw;
  const TblHeader &th = ctx->tbl->th;
  const long long nchunks = R > 0 ? (R + ctx->chunk_rays - 1) / ctx->chunk_rays : 0;
  const bool pipe = ctx->nbuf > 1 && nchunks > 1;
  const bool fov = ctx->fov_n > 0 && R > 0;
  // events: [0] start, [1] end, per chunk: tracer start / tracer done / EGA start / EGA done
  const size_t need_ev = 2 + 4 * (size_t)nchunks;
  while (ctx->events.size() < need_ev) { cudaEvent_t ev; CU(cudaEventCreate(&ev)); ctx->events.push_back(ev); }
  auto EV = [&](long long c, int k) { return ctx->events[2 + 4 * (size_t)c + k]; };
  const size_t per_ray = (size_t)kNLOS * ctx->los.rec, per_ray_raw = (size_t)kNLOS * kRaw;
  long long launches = 0;
  int ngb = ng;
  cudaStream_t st_tr = pipe ? ctx->s_trace : ctx->stream;
  CU(cudaEventRecord(ctx->events[0], ctx->stream));
  if (pipe) CU(cudaStreamWaitEvent(st_tr, ctx->events[0], 0));
  if (ctx->use_fast && nchunks > 0) CU(cudaMemsetAsync(ctx->d_counter.p, 0, (size_t)nchunks * 32, st_tr));
  for (long long c = 0; c < nchunks; c++) {
    const long long r0 = c * ctx->chunk_rays, r1 = std::min(R, r0 + ctx->chunk_rays);
    // buffer of the chunk: [chunk_rays] records, then [chunk_rays] raw points
    double *los_buf = (double *)ctx->d_los.p + (size_t)(c % ctx->nbuf) * (size_t)ctx->chunk_rays * (per_ray + per_ray_raw);
    double *raw_buf = los_buf + (size_t)ctx->chunk_rays * per_ray;
    cudaStream_t st_e = pipe ? ctx->s_ega[c & 1] : ctx->stream;
    TraceArgs t;
    t.n_rays = r1 - r0;
    t.geo = ctx->geo + r0; t.geo_stride = R;
    t.ray_pkg = ctx->ray_pkg + r0;
    t.pkg_atm_off = ctx->pkg_atm_off; t.pkg_atm_np = ctx->pkg_atm_np;
    t.atm_time = ctx->atm + 0 * A; t.atm_z = ctx->atm + 1 * A; t.atm_lon = ctx->atm + 2 * A; t.atm_lat = ctx->atm + 3 * A;
    t.atm_p = ctx->atm + 4 * A; t.atm_t = ctx->atm + 5 * A; t.atm_q = ctx->atm + 6 * A; t.atm_k = ctx->atm + (size_t)(6 + ng) * A;
    t.atm_stride = A;
    t.atm_lnp_slope = (double *)ctx->d_slope.p; t.n_atm = A; t.prepare_atm = (c == 0);
    t.small_blocks = pipe && c > 0;
    t.atm_cart = (double *)ctx->d_slope.p + 2 * A; t.atm_next = (int *)((double *)ctx->d_slope.p + 5 * A);
    t.ip = ctx->ip; t.cz = ctx->cz; t.cx = ctx->cx;
    t.refrac = ctx->refrac; t.ig_h2o = (ctx->ctm_mask & 4) ? ctx->ig_h2o : -1;
    t.rayds = ctx->rayds; t.raydz = ctx->raydz;
    t.los = ctx->los; t.los_data = los_buf; t.raw = raw_buf;
    t.ray_np = (int *)ctx->d_np.p + r0; t.ray_tsurf = (double *)ctx->d_tsurf.p + r0;
    t.ray_level0 = (int *)ctx->d_level0.p + r0;
    t.tp = ctx->o_tp + r0;
    t.tp_host = ctx->ray_out + 2 * R + r0;
    t.error_flag = ctx->err_flag;
    t.tbl = ctx->tbl->td;
    if (pipe && c >= ctx->nbuf) CU(cudaStreamWaitEvent(st_tr, EV(c - ctx->nbuf, 3), 0)); // LOS buffer free again
    // segment-tiled form (jrb_ega_tiled.cuh): a warp handles one ray x 32 channels, shared (p,T) axes
    int use_tiled = 1;
    if (const char *s = getenv("JRB_EGA_TILED")) use_tiled = atoi(s) != 0;
    use_tiled = use_tiled && ctx->use_fast && ctx->cpw == 32 && th.all_shared &&
                ega_tiled_fits(ctx->n_gas_blocks > 1 ? ctx->gases_per_block : ng, ctx->los.rec, (size_t)ctx->smem_optin);
    CU(cudaEventRecord(EV(c, 0), st_tr));
    int nl = 0;
    CU(launch_raytrace(t, st_tr, &nl));
    launches += nl;
    CU(cudaEventRecord(EV(c, 1), st_tr));
    if (ctx->scan_pending) scan_staged_obs(ctx); // on the host, while the device traces rays

    EgaArgs e;
    e.n_rays = r1 - r0; e.ng = ng; e.nd = nd; e.nw = nw;
    e.ctm_mask = ctx->ctm_mask; e.ig_co2 = ctx->ig_co2 >= 0 ? ctx->ig_co2 : 0; e.ig_h2o = ctx->ig_h2o >= 0 ? ctx->ig_h2o : 0;
    e.write_bbt = ctx->write_bbt;
    e.unsorted_columns = th.monotone ? 0 : 1;
    e.per_channel_axes = th.all_shared ? 0 : 1;
    e.los = ctx->los; e.los_data = los_buf;
    e.ray_np = (const int *)ctx->d_np.p + r0; e.ray_tsurf = (const double *)ctx->d_tsurf.p + r0;
    e.chan = (const double *)ctx->d_chan.p; e.window = (const int *)ctx->d_window.p;
    e.tbl = ctx->tbl->td;
    e.rad = ctx->o_rad + (size_t)r0 * nd; e.tau = ctx->o_tau + (size_t)r0 * nd;
    // with the FOV epilogue the pencil-beam values stay on the device; the convolved ones are published afterwards
    e.rad_host = fov ? nullptr : ctx->ray_out + 0 * R + r0;
    e.tau_host = fov ? nullptr : ctx->ray_out + 1 * R + r0;
    e.los_evict_first = ctx->los_evict_first;
    if (getenv("JRB_NO_HOST_ROWS")) { e.rad_host = nullptr; e.tau_host = nullptr; } // measurement only: results stay on the device
    e.work_counter = (unsigned long long *)ctx->d_counter.p + 4 * c;
    e.balance = e.work_counter + 1;
    e.phase_lock_mode = -1;
    if (const char *s = getenv("JRB_EGA_LOCKSTEP")) { const int v = atoi(s); if (v == 0 || v == 1) e.phase_lock_mode = v; } // experiments
    e.cpw = ctx->cpw;
    ctx->stats.ega_channels_per_warp = ctx->use_fast ? e.cpw : 0;
    e.work_chunk = 0; // 0: the launcher picks one item per warp of the CTA
    if (const char *s = getenv("JRB_EGA_CHUNK")) { const int v = atoi(s); if (v >= 1 && v <= 200) e.work_chunk = v; } // experiments
    if (pipe) CU(cudaStreamWaitEvent(st_e, EV(c, 1), 0));
    CU(cudaEventRecord(EV(c, 2), st_e));
    e.use_tiled = use_tiled;
    e.tail_perm = (int *)ctx->d_tail.p + (size_t)c * kTailCap; e.tail_cap = kTailCap; e.tail_n = 0;
    e.n_gas_blocks = ctx->n_gas_blocks; e.gases_per_block = ctx->gases_per_block; e.blocks_per_group = ctx->blocks_per_group;
    e.partial = nullptr; e.partial_len = nullptr; e.seg_pre = nullptr;
    if (ctx->use_fast && ctx->n_gas_blocks > 1) { // split mode: gas-block passes, then the combine kernel
      char *pb = (char *)ctx->d_partial.p;
      e.seg_pre = (double2 *)pb; pb += (size_t)e.n_rays * kNLOS * nd * 16;
      e.partial = (double *)pb; pb += (size_t)ctx->n_gas_blocks * (size_t)e.n_rays * kNLOS * nd * 8;
      e.partial_len = (int *)pb;
      CU(launch_ega_split_passes(e, st_e));
      ctx->stats.ega_tiled = e.use_tiled;
      CU(launch_ega_segments(e, st_e));
      CU(launch_ega_combine(e, st_e));
      launches += 3 + ((e.phase_lock_mode < 0 && e.n_rays > 0) ? 1 : 0); // pass kernel (+ balance), segment kernel, combine kernel
    } else if (ctx->use_fast) {
      if (e.use_tiled) { int nk = 1; CU(launch_ega_tiled(e, st_e, &nk)); ctx->stats.ega_tiled = 1; launches += nk; } // + balance, tail sort
      else { CU(launch_ega_fast(e, st_e, &ngb)); ctx->stats.ega_tiled = 0; launches += (e.phase_lock_mode < 0 && e.n_rays > 0) ? 2 : 1; }
    } else { CU(launch_ega_generic(e, st_e)); launches++; }
    CU(cudaEventRecord(EV(c, 3), st_e));
  }
  if (ctx->scan_pending) scan_staged_obs(ctx); // (a batch without rays)
  if (pipe) {
    CU(cudaStreamWaitEvent(ctx->stream, EV(nchunks - 1, 3), 0));
    if (nchunks > 1) CU(cudaStreamWaitEvent(ctx->stream, EV(nchunks - 2, 3), 0));
  }
  ctx->fov_applied = false;
  if (fov) { // formod(); formod_fov(); of the reference: mask first, then convolve
    const size_t n_out = (size_t)R * nd;
    CU(ctx->d_fov_out.ensure(2 * n_out * 8));
    CU(ctx->d_flag.ensure(256));
    CU(cudaMemsetAsync(ctx->d_flag.p, 0, 4, ctx->stream));
    if (!ctx->nan_mask.empty()) {
      std::vector<long long> flat(ctx->nan_mask.size());
      for (size_t i = 0; i < flat.size(); i++) flat[i] = ctx->nan_mask[i].first * nd + ctx->nan_mask[i].second;
      CU(ctx->d_mask.ensure(flat.size() * 8));
      CU(cudaMemcpyAsync(ctx->d_mask.p, flat.data(), flat.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream)); // flat is a local
      CU(launch_nan_mask(ctx->o_rad, (const long long *)ctx->d_mask.p, (long long)flat.size(), ctx->stream));
      launches++;
    }
    FovArgs f;
    f.n_rays = R; f.nd = nd; f.n_shape = ctx->fov_n;
    f.dz = (const double *)ctx->d_fov.p; f.w = f.dz + ctx->fov_n;
    f.ray_pkg = ctx->ray_pkg; f.time = ctx->geo + 6 * R; f.vpz = ctx->geo + 3 * R;
    f.rad_in = ctx->o_rad; f.tau_in = ctx->o_tau;
    f.rad_out = (double *)ctx->d_fov_out.p; f.tau_out = f.rad_out + n_out;
    f.error = (int *)ctx->d_flag.p;
    CU(launch_fov(f, ctx->stream));
    launches++;
    CU(cudaMemcpyAsync(ctx->o_rad, f.rad_out, 2 * n_out * 8, cudaMemcpyDeviceToDevice, ctx->stream)); // o_tau follows o_rad
    CU(launch_publish(ctx->o_rad, ctx->o_tau, ctx->ray_out + 0 * R, ctx->ray_out + 1 * R, R, nd, ctx->stream));
    launches++;
    ctx->fov_applied = true;
  }
  CU(cudaEventRecord(ctx->events[1], ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->err_flag && (ctx->err_flag[0] | ctx->err_flag[1] | ctx->err_flag[2])) {
    const int e0 = ctx->err_flag[0], e1 = ctx->err_flag[1];
    ctx->err_flag[0] = ctx->err_flag[1] = ctx->err_flag[2] = 0;
    if (e1) return ctx->fail(JRB_ERR_ARG, "Cannot identify profiles. Check ordering of data points!"); // src/jurassic.c:727
    if (!e0) return ctx->fail(JRB_ERR_ARG, "Distance of profiles is too large!");                      // src/jurassic.c:728
    return ctx->fail(JRB_ERR_LIMIT, "Too many LOS points!"); // like the reference's CPU path (src/jr_common.h:693-695)
  }
  if (ctx->fov_applied) {
    int flag = 0;
    CU(cudaMemcpy(&flag, ctx->d_flag.p, 4, cudaMemcpyDeviceToHost));
    if (flag) return ctx->fail(JRB_ERR_ARG, "Cannot apply FOV convolution!"); // src/jurassic.c:236
  }
  // apply_mask (src/jr_common.h:203-210) where the results have landed: the caller's rows (direct) or the pinned buffer
  // (with the FOV epilogue the mask went in on the device, before the convolution)
  if (!ctx->fov_applied && !ctx->nan_mask.empty()) {
    const double nan = std::nan("");
    if (ctx->direct) {
      size_t k = 0;
      for (auto &m : ctx->nan_mask) {
        while (m.first >= ctx->pk_ray_off[k + 1]) ++k;
        const jrb_obs_view &o = ctx->st_obs[k];
        o.rad[(size_t)(m.first - ctx->pk_ray_off[k]) * o.row_stride + m.second] = nan;
      }
    } else {
      double *hrad = (double *)ctx->h_out.p;
      for (auto &m : ctx->nan_mask) hrad[(size_t)m.first * nd + m.second] = nan;
    }
  }
  float ms_rt = 0, ms_ega = 0, ms_tot = 0, ms;
  for (long long c = 0; c < nchunks; c++) {
    CU(cudaEventElapsedTime(&ms, EV(c, 0), EV(c, 1))); ms_rt += ms;
    CU(cudaEventElapsedTime(&ms, EV(c, 2), EV(c, 3))); ms_ega += ms;
  }
  CU(cudaEventElapsedTime(&ms_tot, ctx->events[0], ctx->events[1]));
  // with pipelined chunks the per-kernel spans overlap each other: ms_raytrace + ms_ega may exceed ms_total_device
  ctx->stats.ms_raytrace = ms_rt; ctx->stats.ms_ega = ms_ega; ctx->stats.ms_total_device = ms_tot;
  ctx->stats.n_kernel_launches = launches;
  ctx->stats.cum_runs++; ctx->stats.cum_launches += launches; ctx->stats.cum_ega_launches += nchunks;
  ctx->stats.cum_ms_ega += ms_ega; ctx->stats.cum_ms_raytrace += ms_rt; ctx->stats.cum_ms_device += ms_tot;
  ctx->stats.n_chunks = (int)nchunks; ctx->stats.pipelined = pipe ? 1 : 0;
  ctx->stats.ega_kernel_variant = ctx->use_fast; ctx->stats.ega_ngb = ctx->use_fast ? ngb : 0;
  ctx->stats.ega_ctm_mask = ctx->ctm_mask;
  ctx->stats.ega_gas_blocks = ctx->use_fast ? ctx->n_gas_blocks : 0;
  ctx->stats.ega_per_channel_axes = (ctx->use_fast && !th.all_shared) ? 1 : 0;
  ctx->ran = true; ctx->np_fetched = false;
  return JRB_OK;
}

int jrb_run_staged(jrb_context *ctx) {
  if (!ctx) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  return run_locked(ctx);
}

// staged mode: scatter the pinned result buffer into the caller's obs_t rows; direct mode: the results are there already
static int fetch_locked(jrb_context *ctx, int npk, const jrb_obs_view *obs) {
  if (!ctx->ran) return ctx->fail(JRB_ERR_STATE, "jrb_run_staged has not completed");
  if (npk != ctx->npk) return ctx->fail(JRB_ERR_ARG, "package count differs from the staged batch");
  for (int k = 0; k < npk; k++)
    if (obs[k].nr != ctx->pk_nr[k]) return ctx->fail(JRB_ERR_ARG, "obs[k].nr changed between stage and fetch");
  const long long R = ctx->n_rays;
  const int nd = ctx->nd;
  const double t0 = now_ms();
  ctx->stats.d2h_bytes = (long long)(((size_t)2 * R * nd + 3 * (size_t)R) * 8); // stored by the kernels into host memory
  ctx->stats.host_ms_d2h = 0.f;
  bool same = ctx->direct;
  for (int k = 0; k < npk && same; k++)
    same = obs[k].rad == ctx->st_obs[k].rad && obs[k].tau == ctx->st_obs[k].tau && obs[k].tpz == ctx->st_obs[k].tpz;
  if (same) { ctx->stats.host_ms_scatter = 0.f; return JRB_OK; }
  if (ctx->direct) { // results were landed in other (registered) blocks than the ones asked for now: copy from the device
    CU(cudaSetDevice(ctx->device));
    const size_t out_bytes = ((size_t)2 * R * nd + 3 * (size_t)R) * 8;
    CU(ctx->h_out.ensure(out_bytes + 256));
    CU(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.host_ms_d2h = (float)(now_ms() - t0);
    if (!ctx->fov_applied) { const double nan = std::nan(""); for (auto &m : ctx->nan_mask) ((double *)ctx->h_out.p)[(size_t)m.first * nd + m.second] = nan; }
  }
  const double *hrad = (const double *)ctx->h_out.p, *htau = hrad + (size_t)R * nd, *htp = htau + (size_t)R * nd;
  const double t1 = now_ms();
#pragma omp parallel for schedule(dynamic, 4) num_threads(host_threads())
  for (int k = 0; k < npk; k++) {
    const jrb_obs_view &o = obs[k];
    const long long r0 = ctx->pk_ray_off[k];
    const int nr = ctx->pk_nr[k];
    std::memcpy(o.tpz, htp + 0 * R + r0, (size_t)nr * 8);
    std::memcpy(o.tplon, htp + 1 * R + r0, (size_t)nr * 8);
    std::memcpy(o.tplat, htp + 2 * R + r0, (size_t)nr * 8);
    if (o.row_stride == nd) { // nd == ND: the package's rad / tau blocks are contiguous, one copy each
      std::memcpy(o.rad, hrad + (size_t)r0 * nd, (size_t)nr * nd * 8);
      std::memcpy(o.tau, htau + (size_t)r0 * nd, (size_t)nr * nd * 8);
    } else {
      for (int ir = 0; ir < nr; ir++) {
        double *rr = o.rad + (size_t)ir * o.row_stride, *tt = o.tau + (size_t)ir * o.row_stride;
        std::memcpy(rr, hrad + (size_t)(r0 + ir) * nd, (size_t)nd * 8);
        std::memcpy(tt, htau + (size_t)(r0 + ir) * nd, (size_t)nd * 8);
        for (int id = nd; id < o.nd_reset; id++) { rr[id] = 0.0; tt[id] = 1.0; } // all ND columns are reset (src/CPUdrivers.c:58-60)
      }
    }
  }
  ctx->stats.host_ms_scatter = (float)(now_ms() - t1);
  return JRB_OK;
}

int jrb_fetch_staged(jrb_context *ctx, int npk, const jrb_obs_view *obs) {
  if (!ctx || (npk > 0 && !obs)) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  return fetch_locked(ctx, npk, obs);
}

int jrb_formod_batch(jrb_context *ctx, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs) {
  if (!ctx || npk < 0 || (npk > 0 && (!atm || !obs))) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx); // one batch at a time per context; concurrency = several contexts (lanes)
  const double t0 = now_ms();
  int rc = stage_locked(ctx, npk, atm, obs, /*defer_scan=*/true);
  if (rc != JRB_OK) return rc;
  const double t1 = now_ms();
  rc = run_locked(ctx);
  if (rc != JRB_OK) return rc;
  const double t2 = now_ms();
  rc = fetch_locked(ctx, npk, obs);
  if (getenv("JRB_DEBUG_TIMING"))
    fprintf(stderr, "[jrb] %s stage %.2f ms  run %.2f ms  fetch %.2f ms (pack %.2f h2d %.2f dev %.2f scatter %.2f)\n",
            ctx->direct ? "direct" : "staged", t1 - t0, t2 - t1, now_ms() - t2, ctx->stats.host_ms_pack, ctx->stats.host_ms_h2d,
            ctx->stats.ms_total_device, ctx->stats.host_ms_scatter);
  return rc;
}

int jrb_staged_results(jrb_context *ctx, double **rad_dev, double **tau_dev, long long *n_rays, int *nd) {
  if (!ctx) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->staged) return ctx->fail(JRB_ERR_STATE, "nothing staged");
  if (rad_dev) *rad_dev = ctx->o_rad;
  if (tau_dev) *tau_dev = ctx->o_tau;
  if (n_rays) *n_rays = ctx->n_rays;
  if (nd) *nd = ctx->nd;
  return JRB_OK;
}

// the compact device results of the last run as one block: rad[R][nd], tau[R][nd], tpz[R], tplon[R], tplat[R]
int jrb_staged_results_blob(jrb_context *ctx, void **dev, size_t *bytes, long long *n_rays, int *nd) {
  if (!ctx) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->ran) return ctx->fail(JRB_ERR_STATE, "no completed run");
  if (dev) *dev = ctx->d_out.p;
  if (bytes) *bytes = ((size_t)2 * ctx->n_rays * ctx->nd + 3 * (size_t)ctx->n_rays) * 8;
  if (n_rays) *n_rays = ctx->n_rays;
  if (nd) *nd = ctx->nd;
  return JRB_OK;
}

int jrb_debug_los(jrb_context *ctx, long long ray, double *out, int max_doubles, int *np_out, int *rec_doubles,
                  double *tsurf_out) {
  if (!ctx) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (!ctx->ran) return ctx->fail(JRB_ERR_STATE, "jrb_run_staged has not completed");
  if (ray < 0 || ray >= ctx->n_rays) return ctx->fail(JRB_ERR_ARG, "ray out of range");
  const long long nchunks = (ctx->n_rays + ctx->chunk_rays - 1) / ctx->chunk_rays;
  const long long c = ray / ctx->chunk_rays;
  if (c < nchunks - ctx->nbuf) return ctx->fail(JRB_ERR_STATE, "LOS of this ray was overwritten by a later chunk");
  const size_t per_ray = (size_t)kNLOS * ctx->los.rec, per_ray_raw = (size_t)kNLOS * kRaw;
  const double *buf = (const double *)ctx->d_los.p + (size_t)(c % ctx->nbuf) * (size_t)ctx->chunk_rays * (per_ray + per_ray_raw);
  const double *src = buf + (size_t)(ray - c * ctx->chunk_rays) * per_ray;
  const double *src_raw = buf + (size_t)ctx->chunk_rays * per_ray + (size_t)(ray - c * ctx->chunk_rays) * per_ray_raw;
  CU(cudaSetDevice(ctx->device));
  int np = 0;
  CU(cudaMemcpy(&np, (int *)ctx->d_np.p + ray, 4, cudaMemcpyDeviceToHost));
  if (tsurf_out) CU(cudaMemcpy(tsurf_out, (double *)ctx->d_tsurf.p + ray, 8, cudaMemcpyDeviceToHost));
  if (np_out) *np_out = np;
  // a point as reported: the record, then the tail of its raw point (altitude, raw step length, level, Cartesian position)
  const int rec = ctx->los.rec, out_rec = rec + (kRaw - kRawTail);
  if (rec_doubles) *rec_doubles = out_rec;
  const size_t n = (size_t)np * out_rec;
  if (out) {
    if ((size_t)max_doubles < n) return ctx->fail(JRB_ERR_ARG, "output buffer too small");
    std::vector<double> hr((size_t)np * rec), hw((size_t)np * kRaw);
    if (np > 0) {
      CU(cudaMemcpy(hr.data(), src, hr.size() * 8, cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(hw.data(), src_raw, hw.size() * 8, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < np; i++) {
      std::copy(hr.begin() + (size_t)i * rec, hr.begin() + (size_t)(i + 1) * rec, out + (size_t)i * out_rec);
      std::copy(hw.begin() + (size_t)i * kRaw + kRawTail, hw.begin() + (size_t)(i + 1) * kRaw, out + (size_t)i * out_rec + rec);
    }
  }
  return JRB_OK;
}

int jrb_get_stats(const jrb_context *cctx, jrb_stats *out) {
  if (!cctx || !out) return JRB_ERR_ARG;
  jrb_context *ctx = const_cast<jrb_context *>(cctx);
  std::lock_guard<std::mutex> lk(ctx->mtx);
  if (ctx->ran && !ctx->np_fetched) {
    CU(cudaSetDevice(ctx->device));
    std::vector<int> np((size_t)ctx->n_rays);
    if (ctx->n_rays > 0) CU(cudaMemcpy(np.data(), ctx->d_np.p, (size_t)ctx->n_rays * 4, cudaMemcpyDeviceToHost));
    long long s = 0;
    for (int v : np) s += v;
    ctx->stats.n_los_points = s;
    // lock-step decision of the specialised kernel (made on the device from the balance words of LOS chunk 0)
    ctx->stats.ega_phase_lock = 0;
    if (ctx->use_fast && ctx->n_rays > 0) {
      unsigned long long w[4] = {0, 0, 0, 0};
      CU(cudaMemcpy(w, ctx->d_counter.p, sizeof(w), cudaMemcpyDeviceToHost));
      int mode = -1;
      if (const char *e = getenv("JRB_EGA_LOCKSTEP")) { const int v = atoi(e); if (v == 0 || v == 1) mode = v; }
      ctx->stats.ega_phase_lock = mode >= 0 ? mode : (w[1] * 32ull < w[2] ? 1 : 0);
    }
    ctx->np_fetched = true;
  }
  *out = ctx->stats;
  return JRB_OK;
}

} // extern "C"
