// jrb_group.cu -- devices and lanes behind one handle (C ABI: jrb_group_* in include/jurassic_b200.h).
//
// What the reference has here (src/GPUdrivers.cu:275-358): up to 4 "lanes" (device buffers + stream) handed out round
// robin to the calling host threads, and a vestigial per-device loop (the device ordinal comes from ctl->MPIlocalrank,
// which nothing sets; the loop passes pointers of one device to all of them).  No collective call site exists.
//
// Here:
//   * a group owns `ndev` devices x `nlanes` contexts.  Concurrent callers (OpenMP host threads of a retrieval calling
//     formod_GPU with one package each) get a lane each and their kernels overlap on the device; a large batch is cut into
//     contiguous package slices, one per device, each processed by its own host thread, and every device stores its results
//     straight into the caller's obs rows (see jrb_runtime.cu) -- nothing is gathered through device 0;
//   * the packed tables are built once and broadcast with NCCL over NVLink: ncclCommInitAll + ncclBroadcast inside one
//     process, or rank style (one process per GPU: jrb_group_dist_init with an ncclUniqueId from the launcher);
//   * rank style also offers the gather of the north star: every rank's results to the root rank's host obs rows
//     (grouped ncclSend/ncclRecv of the compact device results, then page-locked D2H + scatter on the root).
// NCCL is loaded at run time (dlopen "libnccl.so.2"), so single-GPU users need no NCCL and a host process that already
// carries an NCCL (e.g. PyTorch's) shares that copy.
#include "jrb_host.h"

#include <nccl.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fcntl.h>
#include <memory>
#include <mutex>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

extern "C" int jrb_staged_results_blob(jrb_context *ctx, void **dev, size_t *bytes, long long *n_rays, int *nd);

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi *nccl_api(std::string &err) {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("NCCL is required for multi-GPU operation but could not be loaded: ") + dlerror(); return; }
#define JRB_SYM(field, name)                                                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));                      \
  if (!api.field) { api.error = std::string("NCCL symbol missing: ") + name; return; }
    JRB_SYM(GetVersion, "ncclGetVersion")
    JRB_SYM(GetUniqueId, "ncclGetUniqueId")
    JRB_SYM(CommInitRank, "ncclCommInitRank")
    JRB_SYM(CommInitAll, "ncclCommInitAll")
    JRB_SYM(CommDestroy, "ncclCommDestroy")
    JRB_SYM(Broadcast, "ncclBroadcast")
    JRB_SYM(Send, "ncclSend")
    JRB_SYM(Recv, "ncclRecv")
    JRB_SYM(GroupStart, "ncclGroupStart")
    JRB_SYM(GroupEnd, "ncclGroupEnd")
    JRB_SYM(GetErrorString, "ncclGetErrorString")
#undef JRB_SYM
  });
  if (!api.error.empty()) { err = api.error; return nullptr; }
  return &api;
}

struct LanePool {
  std::mutex m;
  std::condition_variable cv;
  std::vector<char> busy;
  int n_busy = 0;
};

inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct jrb_group {
  int ndev = 0, nlanes = 1;
  std::vector<int> devices;
  std::vector<std::vector<jrb_context *>> ctx; // [device][lane]
  std::vector<std::unique_ptr<LanePool>> pool;
  std::vector<std::vector<int>> lane_fov;      // FOV version applied to the lane (0 = off)
  std::vector<int> last_lane;                  // lane of the device's most recent batch (whose results a gather sends)
  std::mutex err_m;
  std::string err;
  // control / FOV shape shared by all lanes (pushed to a lane when it is handed out)
  std::mutex cfg_m;
  bool have_ctl = false;
  jrb_ctl_view ctl{};
  std::vector<double> ctl_nu;
  std::vector<int> ctl_window;
  int fov_version = 0;
  std::vector<double> fov_dz, fov_w;
  bool have_tables = false;
  // NCCL
  std::vector<ncclComm_t> comms;       // one per device (single process), or one (rank style)
  std::vector<cudaStream_t> cstreams;  // streams of the collectives
  int dist_rank = -1, dist_n = 0;
  std::atomic<unsigned> rr{0};
  // root side of the rank-style gather
  void *d_gather = nullptr; size_t d_gather_cap = 0;
  void *h_gather = nullptr; size_t h_gather_cap = 0;
  // last call
  jrb_group_stats stats{};

  int fail(int code, const std::string &m) { std::lock_guard<std::mutex> lk(err_m); err = m; return code; }
};

namespace {

int set_ctl_copy(jrb_group *g, const jrb_ctl_view *c) {
  std::lock_guard<std::mutex> lk(g->cfg_m);
  g->ctl = *c;
  g->ctl_nu.assign(c->nu, c->nu + c->nd);
  g->ctl_window.assign(c->window, c->window + c->nd);
  g->ctl.nu = g->ctl_nu.data();
  g->ctl.window = g->ctl_window.data();
  g->have_ctl = true;
  return JRB_OK;
}

// hand out a lane of device d; big batches want lane 0 (its line-of-sight scratch may grow, the others are capped)
int acquire_lane(jrb_group *g, int d, bool want_primary) {
  LanePool &p = *g->pool[d];
  std::unique_lock<std::mutex> lk(p.m);
  int lane = -1;
  p.cv.wait(lk, [&] {
    if (want_primary) { if (!p.busy[0]) { lane = 0; return true; } return false; }
    for (int l = 0; l < (int)p.busy.size(); l++) if (!p.busy[l]) { lane = l; return true; }
    return false;
  });
  p.busy[lane] = 1; p.n_busy++;
  return lane;
}
int try_acquire_lane(jrb_group *g, int d) {
  LanePool &p = *g->pool[d];
  std::lock_guard<std::mutex> lk(p.m);
  for (int l = 0; l < (int)p.busy.size(); l++) if (!p.busy[l]) { p.busy[l] = 1; p.n_busy++; return l; }
  return -1;
}
void release_lane(jrb_group *g, int d, int lane) {
  LanePool &p = *g->pool[d];
  { std::lock_guard<std::mutex> lk(p.m); p.busy[lane] = 0; p.n_busy--; }
  p.cv.notify_all();
}

// control, FOV shape and package slice -> one context; returns the context's code (message copied into the group)
int run_slice(jrb_group *g, int d, int lane, const jrb_ctl_view *ctl, int use_fov, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs) {
  jrb_context *c = g->ctx[d][lane];
  int rc = JRB_OK;
  {
    std::lock_guard<std::mutex> lk(g->cfg_m);
    const jrb_ctl_view *cv = ctl ? ctl : (g->have_ctl ? &g->ctl : nullptr);
    if (!cv) return g->fail(JRB_ERR_STATE, "no control set");
    rc = jrb_set_control(c, cv);
    if (rc == JRB_OK) {
      const int want = use_fov ? g->fov_version : 0;
      if (g->lane_fov[d][lane] != want) {
        rc = want ? jrb_set_fov(c, (int)g->fov_dz.size(), g->fov_dz.data(), g->fov_w.data()) : jrb_set_fov(c, 0, nullptr, nullptr);
        if (rc == JRB_OK) g->lane_fov[d][lane] = want;
      }
    }
  }
  if (rc == JRB_OK) rc = jrb_formod_batch(c, npk, atm, obs);
  g->last_lane[d] = lane;
  if (rc != JRB_OK) g->fail(rc, std::string("device ") + std::to_string(g->devices[d]) + ": " + jrb_last_error(c));
  return rc;
}

} // namespace

extern "C" {

const char *jrb_group_last_error(const jrb_group *g) { return g ? g->err.c_str() : jrb_last_error(nullptr); }

int jrb_group_create(jrb_group **out, int ndev, const int *devices, int nlanes) {
  if (!out) return JRB_ERR_ARG;
  *out = nullptr;
  const int avail = jrb_device_count();
  if (avail < 1) return JRB_ERR_CUDA; // jrb_create would say why; there is no CPU fallback
  if (ndev <= 0) ndev = avail;
  if (nlanes < 1) nlanes = 1;
  if (nlanes > 8) nlanes = 8;
  jrb_group *g = new jrb_group();
  g->ndev = ndev; g->nlanes = nlanes;
  for (int d = 0; d < ndev; d++) g->devices.push_back(devices ? devices[d] : d);
  g->ctx.resize(ndev); g->lane_fov.resize(ndev); g->last_lane.assign(ndev, 0);
  for (int d = 0; d < ndev; d++) {
    g->pool.emplace_back(new LanePool());
    g->pool[d]->busy.assign(nlanes, 0);
    g->lane_fov[d].assign(nlanes, 0);
    for (int l = 0; l < nlanes; l++) {
      jrb_context *c = nullptr;
      const int rc = jrb_create(&c, g->devices[d]);
      if (rc != JRB_OK) { jrb_group_destroy(g); return rc; }
      if (l > 0) jrb_set_los_limit_gb(c, 6.0); // side lanes serve small concurrent calls
      g->ctx[d].push_back(c);
    }
  }
  *out = g;
  return JRB_OK;
}

void jrb_group_destroy(jrb_group *g) {
  if (!g) return;
  std::string e;
  NcclApi *api = g->comms.empty() ? nullptr : nccl_api(e);
  for (size_t i = 0; i < g->comms.size(); i++) if (api && g->comms[i]) api->CommDestroy(g->comms[i]);
  for (size_t d = 0; d < g->cstreams.size(); d++) if (g->cstreams[d]) { cudaSetDevice(g->devices[d < g->devices.size() ? d : 0]); cudaStreamDestroy(g->cstreams[d]); }
  if (g->d_gather) { cudaSetDevice(g->devices[0]); cudaFree(g->d_gather); }
  if (g->h_gather) cudaFreeHost(g->h_gather);
  for (auto &dv : g->ctx) for (jrb_context *c : dv) jrb_destroy(c);
  delete g;
}

int jrb_group_size(const jrb_group *g, int *ndev, int *nlanes) {
  if (!g) return JRB_ERR_ARG;
  if (ndev) *ndev = g->ndev;
  if (nlanes) *nlanes = g->nlanes;
  return JRB_OK;
}

jrb_context *jrb_group_context(jrb_group *g, int dev, int lane) {
  if (!g || dev < 0 || dev >= g->ndev || lane < 0 || lane >= g->nlanes) return nullptr;
  return g->ctx[dev][lane];
}

int jrb_group_set_control(jrb_group *g, const jrb_ctl_view *ctl) {
  if (!g || !ctl || ctl->nd < 1 || !ctl->nu || !ctl->window) return JRB_ERR_ARG;
  return set_ctl_copy(g, ctl);
}

int jrb_group_set_fov(jrb_group *g, int n, const double *dz, const double *w) {
  if (!g || n < 0 || (n > 0 && (!dz || !w))) return JRB_ERR_ARG;
  std::lock_guard<std::mutex> lk(g->cfg_m);
  if (n == (int)g->fov_dz.size() && (n == 0 || (std::equal(dz, dz + n, g->fov_dz.begin()) && std::equal(w, w + n, g->fov_w.begin())))) return JRB_OK;
  g->fov_dz.assign(dz, dz + n); g->fov_w.assign(w, w + n);
  g->fov_version++;
  return JRB_OK;
}

// lanes 1.. of every device share the tables of lane 0
static int share_to_lanes(jrb_group *g) {
  std::lock_guard<std::mutex> lk(g->cfg_m);
  for (int d = 0; d < g->ndev; d++)
    for (int l = 1; l < g->nlanes; l++) {
      int rc = jrb_set_control(g->ctx[d][l], &g->ctl);
      if (rc == JRB_OK) rc = jrb_tables_share(g->ctx[d][l], g->ctx[d][0]);
      if (rc != JRB_OK) return g->fail(rc, jrb_last_error(g->ctx[d][l]));
    }
  g->have_tables = true;
  return JRB_OK;
}

static int ensure_cstreams(jrb_group *g) {
  while ((int)g->cstreams.size() < g->ndev) {
    const int d = (int)g->cstreams.size();
    cudaStream_t s = nullptr;
    if (cudaSetDevice(g->devices[d]) != cudaSuccess || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess)
      return g->fail(JRB_ERR_CUDA, "cudaStreamCreate for collectives failed");
    g->cstreams.push_back(s);
  }
  return JRB_OK;
}

#define NC(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess) return g->fail(JRB_ERR_CUDA, std::string(#call) + ": " + api->GetErrorString(r_)); \
  } while (0)
#define CG(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return g->fail(JRB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

// Pack the tables once, upload them to the first device and broadcast the packed blob to the others with NCCL.
int jrb_group_set_tables(jrb_group *g, const jrb_ctl_view *ctl, const jrb_tbl_view *tbl) {
  if (!g || !tbl) return JRB_ERR_ARG;
  if (ctl) { int rc = jrb_group_set_control(g, ctl); if (rc != JRB_OK) return rc; }
  if (!g->have_ctl) return g->fail(JRB_ERR_STATE, "no control set");
  const double t0 = now_ms();
  for (int d = 0; d < g->ndev; d++) {
    int rc = jrb_set_control(g->ctx[d][0], &g->ctl);
    if (rc != JRB_OK) return g->fail(rc, jrb_last_error(g->ctx[d][0]));
  }
  int rc = jrb_set_tables(g->ctx[0][0], tbl);
  if (rc != JRB_OK) return g->fail(rc, jrb_last_error(g->ctx[0][0]));
  g->stats.nccl_nranks = 0;
  if (g->ndev > 1) {
    std::string e;
    NcclApi *api = nccl_api(e);
    if (!api) return g->fail(JRB_ERR_CUDA, e);
    if (g->comms.empty()) {
      g->comms.assign(g->ndev, nullptr);
      NC(api->CommInitAll(g->comms.data(), g->ndev, g->devices.data()));
    }
    rc = ensure_cstreams(g);
    if (rc != JRB_OK) return rc;
    void *src = nullptr; size_t nbytes = 0;
    rc = jrb_tables_blob(g->ctx[0][0], &src, &nbytes);
    if (rc != JRB_OK) return g->fail(rc, jrb_last_error(g->ctx[0][0]));
    std::vector<void *> dst(g->ndev, nullptr);
    dst[0] = src;
    for (int d = 1; d < g->ndev; d++) {
      rc = jrb_tables_alloc_blob(g->ctx[d][0], nbytes, &dst[d]);
      if (rc != JRB_OK) return g->fail(rc, jrb_last_error(g->ctx[d][0]));
    }
    NC(api->GroupStart());
    for (int d = 0; d < g->ndev; d++) {
      CG(cudaSetDevice(g->devices[d]));
      NC(api->Broadcast(src, dst[d], nbytes, ncclUint8, 0, g->comms[d], g->cstreams[d]));
    }
    NC(api->GroupEnd());
    for (int d = 0; d < g->ndev; d++) { CG(cudaSetDevice(g->devices[d])); CG(cudaStreamSynchronize(g->cstreams[d])); }
    for (int d = 1; d < g->ndev; d++) {
      rc = jrb_tables_adopt_blob(g->ctx[d][0]);
      if (rc != JRB_OK) return g->fail(rc, jrb_last_error(g->ctx[d][0]));
    }
    g->stats.nccl_nranks = g->ndev;
    g->stats.table_bytes = (long long)nbytes;
  }
  rc = share_to_lanes(g);
  g->stats.ms_tables = (float)(now_ms() - t0);
  return rc;
}

int jrb_group_formod_batch(jrb_group *g, const jrb_ctl_view *ctl, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs, int use_fov) {
  if (!g || npk < 0 || (npk > 0 && (!atm || !obs))) return JRB_ERR_ARG;
  if (!g->have_tables) return g->fail(JRB_ERR_STATE, "tables must be set before the forward model runs");
  if (npk == 0) return JRB_OK;
  long long R = 0;
  for (int k = 0; k < npk; k++) R += obs[k].nr > 0 ? obs[k].nr : 0;
  const bool big = R > 32768;
  const double t0 = now_ms();
  // slices: contiguous package ranges with about equal numbers of rays; small batches stay on one device
  int nsl = 1;
  if (g->ndev > 1 && npk >= g->ndev && R >= 4352ll * g->ndev) nsl = g->ndev;
  if (nsl == 1) {
    int d = (int)(g->rr.fetch_add(1) % (unsigned)g->ndev), lane = -1;
    if (!big)
      for (int i = 0; i < g->ndev && lane < 0; i++) { const int dd = (d + i) % g->ndev; lane = try_acquire_lane(g, dd); if (lane >= 0) d = dd; }
    if (lane < 0) lane = acquire_lane(g, d, big);
    const int rc = run_slice(g, d, lane, ctl, use_fov, npk, atm, obs);
    release_lane(g, d, lane);
    g->stats.n_slices = 1; g->stats.ms_last_call = (float)(now_ms() - t0);
    return rc;
  }
  std::vector<int> first(nsl + 1, npk);
  {
    long long acc = 0; int s = 0;
    first[0] = 0;
    for (int k = 0; k < npk && s + 1 < nsl; k++) {
      acc += obs[k].nr > 0 ? obs[k].nr : 0;
      if (acc * nsl >= R * (s + 1)) first[++s] = k + 1;
    }
  }
  std::vector<int> lanes(nsl, -1), rcs(nsl, JRB_OK);
  for (int s = 0; s < nsl; s++) lanes[s] = acquire_lane(g, s, big); // ascending device order: concurrent big calls cannot deadlock
  std::vector<std::thread> th;
  auto work = [&](int s) {
    const int k0 = first[s], n = first[s + 1] - first[s];
    rcs[s] = n > 0 ? run_slice(g, s, lanes[s], ctl, use_fov, n, atm + k0, obs + k0) : JRB_OK;
  };
  for (int s = 1; s < nsl; s++) th.emplace_back(work, s);
  work(0);
  for (auto &t : th) t.join();
  for (int s = 0; s < nsl; s++) release_lane(g, s, lanes[s]);
  g->stats.n_slices = nsl; g->stats.ms_last_call = (float)(now_ms() - t0);
  for (int s = 0; s < nsl; s++) if (rcs[s] != JRB_OK) return rcs[s];
  return JRB_OK;
}

int jrb_group_get_stats(jrb_group *g, jrb_group_stats *out) {
  if (!g || !out) return JRB_ERR_ARG;
  g->stats.ndev = g->ndev; g->stats.nlanes = g->nlanes;
  g->stats.dist_rank = g->dist_rank; g->stats.dist_nranks = g->dist_n;
  *out = g->stats;
  return JRB_OK;
}

// ---- rank style: one process per GPU ------------------------------------------------------------------------------------
int jrb_dist_unique_id(void *id, size_t cap) {
  if (!id || cap < sizeof(ncclUniqueId)) return JRB_ERR_ARG;
  std::string e;
  NcclApi *api = nccl_api(e);
  if (!api) return JRB_ERR_CUDA;
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return JRB_ERR_CUDA;
  std::memset(id, 0, cap);
  std::memcpy(id, &u, sizeof(u));
  return JRB_OK;
}

int jrb_group_dist_init(jrb_group *g, int rank, int nranks, const void *id, size_t id_bytes) {
  if (!g || !id || rank < 0 || rank >= nranks || id_bytes < sizeof(ncclUniqueId)) return JRB_ERR_ARG;
  if (g->ndev != 1) return g->fail(JRB_ERR_ARG, "rank-style operation needs a group of exactly one device per process");
  std::string e;
  NcclApi *api = nccl_api(e);
  if (!api) return g->fail(JRB_ERR_CUDA, e);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  CG(cudaSetDevice(g->devices[0]));
  g->comms.assign(1, nullptr);
  NC(api->CommInitRank(&g->comms[0], nranks, u, rank));
  int rc = ensure_cstreams(g);
  if (rc != JRB_OK) return rc;
  g->dist_rank = rank; g->dist_n = nranks;
  g->stats.nccl_nranks = nranks;
  return JRB_OK;
}

// root: tbl != NULL, packs and uploads; everybody: receives the packed blob by ncclBroadcast and adopts it
int jrb_group_dist_set_tables(jrb_group *g, const jrb_ctl_view *ctl, const jrb_tbl_view *tbl, int root) {
  if (!g) return JRB_ERR_ARG;
  if (g->dist_n < 1) return g->fail(JRB_ERR_STATE, "jrb_group_dist_init has not been called");
  if (ctl) { int rc = jrb_group_set_control(g, ctl); if (rc != JRB_OK) return rc; }
  if (!g->have_ctl) return g->fail(JRB_ERR_STATE, "no control set");
  if ((g->dist_rank == root) != (tbl != nullptr)) return g->fail(JRB_ERR_ARG, "exactly the root rank passes the tables");
  std::string e;
  NcclApi *api = nccl_api(e);
  if (!api) return g->fail(JRB_ERR_CUDA, e);
  const double t0 = now_ms();
  jrb_context *c = g->ctx[0][0];
  int rc = jrb_set_control(c, &g->ctl);
  if (rc != JRB_OK) return g->fail(rc, jrb_last_error(c));
  CG(cudaSetDevice(g->devices[0]));
  unsigned long long *d_n = nullptr;
  CG(cudaMalloc(&d_n, 8));
  void *blob = nullptr; size_t nbytes = 0;
  if (tbl) {
    rc = jrb_set_tables(c, tbl);
    if (rc != JRB_OK) { cudaFree(d_n); return g->fail(rc, jrb_last_error(c)); }
    rc = jrb_tables_blob(c, &blob, &nbytes);
    if (rc != JRB_OK) { cudaFree(d_n); return g->fail(rc, jrb_last_error(c)); }
    unsigned long long n64 = nbytes;
    CG(cudaMemcpy(d_n, &n64, 8, cudaMemcpyHostToDevice));
  }
  NC(api->Broadcast(d_n, d_n, 8, ncclUint8, root, g->comms[0], g->cstreams[0]));
  CG(cudaStreamSynchronize(g->cstreams[0]));
  unsigned long long n64 = 0;
  CG(cudaMemcpy(&n64, d_n, 8, cudaMemcpyDeviceToHost));
  cudaFree(d_n);
  nbytes = (size_t)n64;
  if (!tbl) {
    rc = jrb_tables_alloc_blob(c, nbytes, &blob);
    if (rc != JRB_OK) return g->fail(rc, jrb_last_error(c));
  }
  NC(api->Broadcast(blob, blob, nbytes, ncclUint8, root, g->comms[0], g->cstreams[0]));
  CG(cudaStreamSynchronize(g->cstreams[0]));
  if (!tbl) {
    rc = jrb_tables_adopt_blob(c);
    if (rc != JRB_OK) return g->fail(rc, jrb_last_error(c));
  }
  g->stats.table_bytes = (long long)nbytes;
  rc = share_to_lanes(g);
  g->stats.ms_tables = (float)(now_ms() - t0);
  return rc;
}

// The gather of the north star ("each GPU owns a contiguous obs slice, radiances are gathered at the end"): every rank
// sends the compact device results of its last batch (lane 0: rad, tau, tangent points) to the root, which lands them in
// the obs views of ALL packages (rank order; counts[r] = packages of rank r).  The root's own slice is already in place.
int jrb_group_dist_gather(jrb_group *g, int root, const int *counts, int npk_all, const jrb_obs_view *obs_all) {
  if (!g || !counts) return JRB_ERR_ARG;
  if (g->dist_n < 1) return g->fail(JRB_ERR_STATE, "jrb_group_dist_init has not been called");
  std::string e;
  NcclApi *api = nccl_api(e);
  if (!api) return g->fail(JRB_ERR_CUDA, e);
  const double t0 = now_ms();
  jrb_context *c = g->ctx[0][g->last_lane[0]];
  CG(cudaSetDevice(g->devices[0]));
  void *mine = nullptr; size_t my_bytes = 0; long long my_rays = 0; int nd = 0;
  int rc = jrb_staged_results_blob(c, &mine, &my_bytes, &my_rays, &nd);
  if (rc != JRB_OK) return g->fail(rc, jrb_last_error(c));
  if (g->dist_rank != root) {
    if (my_bytes) NC(api->Send(mine, my_bytes, ncclUint8, root, g->comms[0], g->cstreams[0]));
    CG(cudaStreamSynchronize(g->cstreams[0]));
    g->stats.ms_gather = (float)(now_ms() - t0);
    return JRB_OK;
  }
  if (!obs_all) return g->fail(JRB_ERR_ARG, "the root rank passes the obs views of all packages");
  int total = 0;
  for (int r = 0; r < g->dist_n; r++) total += counts[r];
  if (total != npk_all) return g->fail(JRB_ERR_ARG, "counts do not add up to the number of obs views");
  // bytes per rank: (2 nd + 3) doubles per ray, rays from the obs views
  std::vector<long long> rays(g->dist_n, 0);
  std::vector<size_t> off(g->dist_n + 1, 0);
  {
    int k = 0;
    for (int r = 0; r < g->dist_n; r++) {
      for (int i = 0; i < counts[r]; i++, k++) rays[r] += obs_all[k].nr;
      off[r + 1] = off[r] + (r == root ? 0 : (size_t)rays[r] * (2 * (size_t)nd + 3) * 8);
    }
  }
  const size_t need = off[g->dist_n] + 256;
  if (need > g->d_gather_cap) {
    if (g->d_gather) cudaFree(g->d_gather);
    if (g->h_gather) cudaFreeHost(g->h_gather);
    g->d_gather = g->h_gather = nullptr; g->d_gather_cap = g->h_gather_cap = 0;
    CG(cudaMalloc(&g->d_gather, need));
    CG(cudaHostAlloc(&g->h_gather, need, cudaHostAllocPortable));
    g->d_gather_cap = g->h_gather_cap = need;
  }
  NC(api->GroupStart());
  for (int r = 0; r < g->dist_n; r++)
    if (r != root && off[r + 1] > off[r])
      NC(api->Recv((char *)g->d_gather + off[r], off[r + 1] - off[r], ncclUint8, r, g->comms[0], g->cstreams[0]));
  NC(api->GroupEnd());
  CG(cudaMemcpyAsync(g->h_gather, g->d_gather, off[g->dist_n], cudaMemcpyDeviceToHost, g->cstreams[0]));
  CG(cudaStreamSynchronize(g->cstreams[0]));
  const double t1 = now_ms();
  // scatter into the obs rows (what jrb_fetch_staged does for the own slice)
  std::vector<int> k_first(g->dist_n + 1, 0);
  for (int r = 0; r < g->dist_n; r++) k_first[r + 1] = k_first[r] + counts[r];
  for (int r = 0; r < g->dist_n; r++) {
    if (r == root) continue;
    const long long Rr = rays[r];
    const double *hrad = (const double *)((const char *)g->h_gather + off[r]), *htau = hrad + (size_t)Rr * nd, *htp = htau + (size_t)Rr * nd;
    std::vector<long long> r0(counts[r] + 1, 0);
    for (int i = 0; i < counts[r]; i++) r0[i + 1] = r0[i] + obs_all[k_first[r] + i].nr;
#pragma omp parallel for schedule(dynamic, 4) num_threads(jrb::host_threads())
    for (int i = 0; i < counts[r]; i++) {
      const jrb_obs_view &o = obs_all[k_first[r] + i];
      const long long b = r0[i];
      std::memcpy(o.tpz, htp + 0 * Rr + b, (size_t)o.nr * 8);
      std::memcpy(o.tplon, htp + 1 * Rr + b, (size_t)o.nr * 8);
      std::memcpy(o.tplat, htp + 2 * Rr + b, (size_t)o.nr * 8);
      for (int ir = 0; ir < o.nr; ir++) {
        double *rr = o.rad + (size_t)ir * o.row_stride, *tt = o.tau + (size_t)ir * o.row_stride;
        // save_mask / apply_mask (src/jr_common.h:193-210) on the root's copy of the package: non-finite input -> NaN
        double probe = 0.0;
        for (int id = 0; id < nd; id++) probe += rr[id] - rr[id];
        const bool masked = (probe != 0.0 || probe != probe);
        std::vector<int> bad;
        if (masked) for (int id = 0; id < nd; id++) if (!std::isfinite(rr[id])) bad.push_back(id);
        std::memcpy(rr, hrad + (size_t)(b + ir) * nd, (size_t)nd * 8);
        std::memcpy(tt, htau + (size_t)(b + ir) * nd, (size_t)nd * 8);
        for (int id : bad) rr[id] = std::nan("");
        for (int id = nd; id < o.nd_reset; id++) { rr[id] = 0.0; tt[id] = 1.0; }
      }
    }
  }
  g->stats.ms_gather = (float)(now_ms() - t0);
  g->stats.ms_gather_scatter = (float)(now_ms() - t1);
  g->stats.gather_bytes = (long long)off[g->dist_n];
  return JRB_OK;
}

// ---- node-shared page-locked memory ----------------------------------------------------------------------------------------
// One process per GPU on one node: obs_t blocks placed in a POSIX shared-memory segment that every rank maps and
// page-locks are written by each rank's GPU over its own PCIe link while its kernels run -- the "gather" of the results on
// the root rank then costs nothing and does not funnel through one device.  create != 0: make the segment (root), else
// attach.  The segment is unlinked by jrb_shared_free(..., unlink = 1).
int jrb_shared_alloc(const char *name, size_t bytes, int create, void **out) {
  if (!name || !out || bytes == 0) return JRB_ERR_ARG;
  *out = nullptr;
  int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) return JRB_ERR_ARG;
  if (create && ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); return JRB_ERR_ARG; }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { if (create) shm_unlink(name); return JRB_ERR_ARG; }
  const int rc = jrb_host_register(p, bytes);
  if (rc != JRB_OK) { munmap(p, bytes); if (create) shm_unlink(name); return rc; }
  *out = p;
  return JRB_OK;
}

int jrb_shared_free(const char *name, void *ptr, size_t bytes, int unlink_it) {
  if (ptr && bytes) { jrb_host_unregister(ptr, bytes); munmap(ptr, bytes); }
  if (name && unlink_it) shm_unlink(name);
  return JRB_OK;
}

} // extern "C"
