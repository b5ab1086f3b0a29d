/* jr_dropin.c -- the reference-facing C layer: formod_GPU() and its batched extension on top of the
 * dimension-agnostic core (see include/jurassic_b200_dropin.h for the interface each function replaces).
 * Plain C, compiled once per (ND,NG) like the reference's own objects. */
#ifdef JRB_USE_REFERENCE_HEADER
#include "jurassic.h" /* a maintainer building inside the reference tree uses the real header */
#else
#include "jr_structs.h"
#endif
#include <jurassic_b200.h>
#include <jurassic_b200_dropin.h>

#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

/* provided by the reference's CPUdrivers.o when linked into a JURASSIC executable (src/jr_common.h:60-78) */
extern tbl_t *get_tbl(ctl_t const *ctl) __attribute__((weak));
typedef tbl_t *(*get_tbl_fn)(ctl_t const *);

/* link-time weak reference first, then a run-time lookup (the reference objects may be loaded after this library) */
static get_tbl_fn find_get_tbl(void) {
  if (get_tbl) return get_tbl;
  return (get_tbl_fn)dlsym(RTLD_DEFAULT, "get_tbl");
}

#define JR_FATAL(msg)                                                                          \
  do {                                                                                         \
    printf("\nError (%s, %s, l%d): %s\n\n", __FILE__, __func__, __LINE__, msg);                \
    exit(EXIT_FAILURE);                                                                        \
  } while (0)

static jrb_group *g_grp = NULL; /* devices x lanes; created on first use */
static int g_have_tables = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER; /* initialisation only: forward-model calls run concurrently */

/* find_emitter (src/jurassic.c:198-207): case-insensitive, -1 if absent */
static int emitter_index(ctl_t const *ctl, char const *name) {
  for (int ig = 0; ig < ctl->ng; ig++)
    if (0 == strcasecmp(ctl->emitter[ig], name)) return ig;
  return -1;
}

static void fill_ctl_view(ctl_t const *ctl, jrb_ctl_view *v) {
  memset(v, 0, sizeof(*v));
  v->ng = ctl->ng; v->nd = ctl->nd; v->nw = ctl->nw;
  v->nu = ctl->nu; v->window = ctl->window;
  v->ctm_co2 = ctl->ctm_co2; v->ctm_h2o = ctl->ctm_h2o; v->ctm_n2 = ctl->ctm_n2; v->ctm_o2 = ctl->ctm_o2;
  /* looked up only when the continuum is requested, as in the reference (src/CPUdrivers.c:126-128) */
  v->ig_h2o = ctl->ctm_h2o ? emitter_index(ctl, "H2O") : -999;
  v->ig_co2 = ctl->ctm_co2 ? emitter_index(ctl, "CO2") : -999;
  v->refrac = ctl->refrac; v->rayds = ctl->rayds; v->raydz = ctl->raydz; v->hydz = ctl->hydz;
  v->write_bbt = ctl->write_bbt; v->formod = ctl->formod; v->ip = ctl->ip;
  v->cz = ctl->cz; v->cx = ctl->cx;
}

static void fill_tbl_view(tbl_t const *t, jrb_tbl_view *v) {
  v->dim_g = NG; v->dim_p = TBLNP; v->dim_t = TBLNT; v->dim_u = TBLNU; v->dim_d = ND; v->dim_s = TBLNS;
  v->np = &t->np[0][0]; v->nt = &t->nt[0][0][0]; v->nu = &t->nu[0][0][0][0];
  v->p = &t->p[0][0][0]; v->t = &t->t[0][0][0][0];
  v->u = &t->u[0][0][0][0][0]; v->eps = &t->eps[0][0][0][0][0];
  v->sr = &t->sr[0][0]; v->st = &t->st[0];
}

static void fill_atm_view(atm_t *a, jrb_atm_view *v) {
  v->np = a->np;
  v->time = a->time; v->z = a->z; v->lon = a->lon; v->lat = a->lat; v->p = a->p; v->t = a->t;
  v->q = &a->q[0][0]; v->q_stride = NP;
  v->k = &a->k[0][0]; v->k_stride = NP;
  v->q_rows = NULL; v->k_rows = NULL;
}

static void fill_obs_view(obs_t *o, jrb_obs_view *v) {
  v->nr = o->nr;
  v->time = o->time; v->obsz = o->obsz; v->obslon = o->obslon; v->obslat = o->obslat;
  v->vpz = o->vpz; v->vplon = o->vplon; v->vplat = o->vplat;
  v->tpz = o->tpz; v->tplon = o->tplon; v->tplat = o->tplat;
  v->rad = &o->rad[0][0]; v->tau = &o->tau[0][0];
  v->row_stride = ND; v->nd_reset = ND;
}

/* lanes per device: the reference hands out up to 4 (src/GPUdrivers.cu:292-296) */
static int lanes_wanted(void) {
  char const *s = getenv("JRB_LANES");
  int n = s ? atoi(s) : 4;
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}

/* ndev == 1: the device `device` (< 0: ctl->MPIlocalrank like the reference, src/GPUdrivers.cu:288); ndev > 1: devices
 * 0..ndev-1; ndev <= 0: all visible devices */
static void make_group_locked(ctl_t const *ctl, int device, int ndev) {
  if (g_grp) return;
  int const avail = jrb_device_count();
  if (avail < 1) JR_FATAL("no CUDA device available (there is no CPU fallback in this library)");
  if (ndev <= 0) ndev = avail;
  if (ndev > avail) JR_FATAL("More devices requested than visible. Abort.");
  if (ndev == 1) {
    if (device < 0) device = ctl->MPIlocalrank;
    if (device >= avail) JR_FATAL("More MPI-Ranks on Node than GPUs. Abort."); /* src/GPUdrivers.cu:284-287 */
    if (jrb_group_create(&g_grp, 1, &device, lanes_wanted()) != JRB_OK) JR_FATAL(jrb_last_error(NULL));
  } else {
    if (jrb_group_create(&g_grp, ndev, NULL, lanes_wanted()) != JRB_OK) JR_FATAL(jrb_last_error(NULL));
  }
}

static void init_locked(ctl_t const *ctl, tbl_t const *tbl, int device, int ndev) {
  make_group_locked(ctl, device, ndev);
  jrb_ctl_view cv;
  fill_ctl_view(ctl, &cv);
  jrb_tbl_view tv;
  fill_tbl_view(tbl, &tv);
  if (jrb_group_set_tables(g_grp, &cv, &tv) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
  g_have_tables = 1;
}

/* tables straight from the files init_tbl would use (src/jurassic.c:311-416, 612-672), without its 8.8 GB tbl_t: the binary
 * cache "bin.jurassic-fp32-tables-g<NG>-..." in the working directory when ctl->read_binary is set (mapped, not loaded;
 * fatal if it fails and READ_BINARY > 0), else the ASCII .tab/.filt files below ctl->tblbase, after which the cache is
 * written when ctl->write_binary is set */
int jr_b200_init_from_files(ctl_t const *ctl, int device) {
  if (!ctl) JR_FATAL("jr_b200_init_from_files: NULL argument");
  char const *names[NG > 0 ? NG : 1];
  for (int ig = 0; ig < ctl->ng; ig++) names[ig] = ctl->emitter[ig];
  char binname[256];
  jrb_binary_tables_filename(binname, sizeof(binname), NG, TBLNP, TBLNT, TBLNU, ND);
  jrb_host_tables *ht = NULL;
  jrb_tbl_view tv;
  int missing = 0;
  if (ctl->read_binary) {
    if (jrb_tables_read_binary(binname, ctl->ng, names, ctl->nd, ctl->nu, &ht) == JRB_OK) {
      printf("matching binary tables file found\n");
    } else {
      printf("# %s\n", jrb_ingest_last_error());
      if (ctl->read_binary > 0) JR_FATAL("Failed to read binary file while READ_BINARY > 0");
    }
  }
  if (!ht) {
    if (jrb_tables_read_ascii(ctl->tblbase, ctl->ng, names, ctl->nd, ctl->nu, TBLNP, TBLNT, TBLNU, &ht) != JRB_OK)
      JR_FATAL(jrb_ingest_last_error());
    jrb_host_tables_view(ht, &tv, &missing);
    if (missing > 0) printf("Warning! %d files were not found!\n", missing); /* like init_tbl (src/jurassic.c:424-427) */
    if (ctl->write_binary) {
      int const status = jrb_tables_write_binary(binname, &tv, ctl->ng, names, ctl->nd, ctl->nu, NG, TBLNP, TBLNT, TBLNU, ND);
      printf("# jr_write_binary_tables returns status %d\n", status);
    }
  }
  jrb_host_tables_view(ht, &tv, NULL);
  pthread_mutex_lock(&g_lock);
  make_group_locked(ctl, device, 1);
  jrb_ctl_view cv;
  fill_ctl_view(ctl, &cv);
  if (jrb_group_set_tables(g_grp, &cv, &tv) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
  g_have_tables = 1;
  pthread_mutex_unlock(&g_lock);
  jrb_host_tables_free(ht);
  return 0;
}

int jr_b200_init(ctl_t const *ctl, tbl_t const *tbl, int device) {
  if (!ctl || !tbl) JR_FATAL("jr_b200_init: NULL argument");
  pthread_mutex_lock(&g_lock);
  init_locked(ctl, tbl, device, 1);
  pthread_mutex_unlock(&g_lock);
  return 0;
}

/* all devices of the node behind the same calls: one context set and one host thread per device, tables packed once and
 * broadcast with NCCL, batches cut into contiguous package slices (SURVEY.md 8b "needed extension", 8e) */
int jr_b200_init_multi(ctl_t const *ctl, tbl_t const *tbl, int ndevices) {
  if (!ctl || !tbl) JR_FATAL("jr_b200_init_multi: NULL argument");
  pthread_mutex_lock(&g_lock);
  if (g_grp) JR_FATAL("jr_b200_init_multi: already initialised (call jr_b200_finalize first)");
  init_locked(ctl, tbl, -1, ndevices <= 0 ? 0 : ndevices);
  pthread_mutex_unlock(&g_lock);
  int nd_ = 0;
  jrb_group_size(g_grp, &nd_, NULL);
  return nd_;
}

/* ---- rank style (one process per GPU) ---------------------------------------------------------------------------------- */
int jr_b200_dist_unique_id(char id[128]) { return jrb_dist_unique_id(id, 128) == JRB_OK ? 0 : -1; }

int jr_b200_dist_init(ctl_t const *ctl, tbl_t const *tbl, int rank, int nranks, char const id[128], int device, int root) {
  if (!ctl) JR_FATAL("jr_b200_dist_init: NULL argument");
  pthread_mutex_lock(&g_lock);
  if (g_grp) JR_FATAL("jr_b200_dist_init: already initialised (call jr_b200_finalize first)");
  make_group_locked(ctl, device, 1);
  if (jrb_group_dist_init(g_grp, rank, nranks, id, 128) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
  jrb_ctl_view cv;
  fill_ctl_view(ctl, &cv);
  jrb_tbl_view tv;
  if (tbl) fill_tbl_view(tbl, &tv);
  if (jrb_group_dist_set_tables(g_grp, &cv, tbl ? &tv : NULL, root) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
  g_have_tables = 1;
  pthread_mutex_unlock(&g_lock);
  return 0;
}

void jr_b200_dist_gather(obs_t *const obs_all[], int const counts[], int nranks, int root) {
  if (!g_grp) JR_FATAL("jr_b200_dist_gather: not initialised");
  int total = 0;
  for (int r = 0; r < nranks; r++) total += counts[r];
  jrb_obs_view *ov = NULL;
  if (obs_all) {
    ov = (jrb_obs_view *)malloc(sizeof(jrb_obs_view) * (size_t)(total ? total : 1));
    if (!ov) JR_FATAL("Out of memory!");
    for (int i = 0; i < total; i++) fill_obs_view(obs_all[i], &ov[i]);
  }
  if (jrb_group_dist_gather(g_grp, root, counts, total, ov) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
  free(ov);
}

/* ---- page-locking of the caller's structs (direct I/O, see jurassic_b200.h) ------------------------------------------- */
int jr_b200_pin_packages(atm_t *const atm[], obs_t *const obs[], int npackages) {
  for (int i = 0; i < npackages; i++) {
    if (atm && atm[i] && jrb_host_register(atm[i], sizeof(atm_t)) != JRB_OK) return -1;
    if (obs && obs[i] && jrb_host_register(obs[i], sizeof(obs_t)) != JRB_OK) return -1;
  }
  return 0;
}
void jr_b200_unpin_all(void) { jrb_host_unregister_all(); }

void *jr_b200_shared_alloc(char const *name, size_t bytes, int create) {
  void *p = NULL;
  if (jrb_shared_alloc(name, bytes, create, &p) != JRB_OK) return NULL;
  return p;
}
void jr_b200_shared_free(char const *name, void *ptr, size_t bytes, int unlink_it) { jrb_shared_free(name, ptr, bytes, unlink_it); }

/* FOV shape file of ctl->fov, parsed like read_shape (src/jurassic.c:1134-1150); kept for the last file name seen.
 * Returns 1 if the convolution is to be applied. */
static char g_fov_loaded[LEN] = "";
static int push_fov(ctl_t const *ctl, int enable) {
  char *const loaded = g_fov_loaded;
  if (!enable || ctl->fov[0] == '-') return 0; /* "-": do not take the FOV into account (src/jurassic.c:219) */
  pthread_mutex_lock(&g_lock);
  if (strncmp(loaded, ctl->fov, LEN) != 0) {
    static double dz[NSHAPE], w[NSHAPE];
    printf("Read shape function: %s\n", ctl->fov);
    FILE *in = fopen(ctl->fov, "r");
    if (!in) JR_FATAL("Cannot open file!");
    char line[LEN];
    int n = 0;
    while (fgets(line, LEN, in)) {
      double a, b;
      if (sscanf(line, "%lg %lg", &a, &b) == 2) {
        if (n >= NSHAPE) JR_FATAL("Too many data points!");
        dz[n] = a; w[n] = b; n++;
      }
    }
    fclose(in);
    if (n < 1) JR_FATAL("Could not read any data!");
    if (jrb_group_set_fov(g_grp, n, dz, w) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
    strncpy(loaded, ctl->fov, LEN - 1);
  }
  pthread_mutex_unlock(&g_lock);
  return 1;
}

/* internal entry: exported names are reached through these statics so that a formod_GPU/... symbol of another library
 * in the global scope (e.g. the CPU-only stub of the reference, src/CPUdrivers.c:156-176) can never interpose them */
static void formod_batch_impl(ctl_t const *ctl, atm_t *const atm[], obs_t *const obs[], int npackages, int fov) {
  if (ctl->checkmode) { printf("# %s: no operation in checkmode\n", __func__); return; }
  if (npackages <= 0) return;
  struct timespec ts0, ts1, ts2;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  if (!g_have_tables) { /* first call without jr_b200_init: tables from the reference's get_tbl, like its own first call */
    pthread_mutex_lock(&g_lock);
    if (!g_have_tables) {
      get_tbl_fn const gt = find_get_tbl();
      if (!gt) JR_FATAL("tables not initialised: call jr_b200_init() or link the reference's get_tbl()");
      int ndev = 1;
      if (getenv("JRB_NDEVICES")) ndev = atoi(getenv("JRB_NDEVICES"));
      init_locked(ctl, gt(ctl), -1, ndev);
    }
    pthread_mutex_unlock(&g_lock);
  }
  /* control values can change between calls (the reference re-uploads ctl_t on every call, src/GPUdrivers.cu:355): they
   * travel with the call and are applied to the lane that serves it */
  jrb_ctl_view cv;
  fill_ctl_view(ctl, &cv);
  int const use_fov = push_fov(ctl, fov);
  jrb_atm_view *av = (jrb_atm_view *)malloc(sizeof(jrb_atm_view) * (size_t)npackages);
  jrb_obs_view *ov = (jrb_obs_view *)malloc(sizeof(jrb_obs_view) * (size_t)npackages);
  if (!av || !ov) JR_FATAL("Out of memory!");
  for (int i = 0; i < npackages; i++) { fill_atm_view(atm[i], &av[i]); fill_obs_view(obs[i], &ov[i]); }
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  /* concurrent callers (OpenMP host threads of a retrieval) are served by different lanes (src/GPUdrivers.cu:331-334) */
  if (jrb_group_formod_batch(g_grp, &cv, npackages, av, ov, use_fov) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
  free(av); free(ov);
  if (getenv("JRB_DEBUG_TIMING")) {
    clock_gettime(CLOCK_MONOTONIC, &ts2);
    fprintf(stderr, "[jr_dropin] prepare %.2f ms, core call %.2f ms\n", (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6,
            (ts2.tv_sec - ts1.tv_sec) * 1e3 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-6);
  }
}

void jr_b200_formod_batch(ctl_t const *ctl, atm_t *const atm[], obs_t *const obs[], int npackages) {
  formod_batch_impl(ctl, atm, obs, npackages, 0);
}

/* per package: formod(ctl, atm, obs); formod_fov(ctl, obs); with the convolution as a device epilogue */
void jr_b200_formod_fov_batch(ctl_t const *ctl, atm_t *const atm[], obs_t *const obs[], int npackages) {
  formod_batch_impl(ctl, atm, obs, npackages, 1);
}

static void formod_one_impl(ctl_t const *ctl, atm_t *atm, obs_t *obs) {
  atm_t *const a[1] = {atm};
  obs_t *const o[1] = {obs};
  formod_batch_impl(ctl, a, o, 1, 0);
}

void formod_GPU(ctl_t const *ctl, atm_t *atm, obs_t *obs) { formod_one_impl(ctl, atm, obs); }

/* ---- batched finite-difference Jacobian ---------------------------------------------------------
 * What the reference's kernel() computes (src/jurassic.c:812-857): one forward model for the undisturbed atmosphere, one per
 * state-vector element with that element perturbed, K[i][j] = (y_j[i] - y_0[i]) / h_j.  The reference issues the 1+n
 * forward models one after the other; here the n perturbed ones are device batches: every perturbed atmosphere is a view
 * that shares all profiles with the caller's atm_t except the single perturbed one (row-pointer form of jrb_atm_view). */
typedef struct { double *value; int iq; int ip; } jr_state_elem;

/* atm2x / atm2x_help (src/jurassic.c:1491-1513): state vector = retrieved p, T, q[ig], k[iw] inside their altitude ranges */
static size_t state_vector(ctl_t const *ctl, atm_t *atm, jr_state_elem *e) {
  size_t n = 0;
#define JR_ADD(zmin, zmax, arr, code)                                                    \
  for (int ip = 0; ip < atm->np; ip++)                                                   \
    if (atm->z[ip] >= (zmin) && atm->z[ip] <= (zmax)) {                                  \
      if (e) { e[n].value = (arr); e[n].iq = (code); e[n].ip = ip; }                     \
      n++;                                                                               \
    }
  JR_ADD(ctl->retp_zmin, ctl->retp_zmax, atm->p, 0)
  JR_ADD(ctl->rett_zmin, ctl->rett_zmax, atm->t, 1)
  for (int ig = 0; ig < ctl->ng; ig++) JR_ADD(ctl->retq_zmin[ig], ctl->retq_zmax[ig], atm->q[ig], 2 + ig)
  for (int iw = 0; iw < ctl->nw; iw++) JR_ADD(ctl->retk_zmin[iw], ctl->retk_zmax[iw], atm->k[iw], 2 + ctl->ng + iw)
#undef JR_ADD
  return n;
}

/* obs2y (src/jurassic.c:1528-1541): the measurement vector holds the finite radiances, ray-major */
static size_t measurement_count(ctl_t const *ctl, obs_t const *obs) {
  size_t m = 0;
  for (int ir = 0; ir < obs->nr; ir++)
    for (int id = 0; id < ctl->nd; id++) m += isfinite(obs->rad[ir][id]) ? 1 : 0;
  return m;
}

size_t jr_b200_kernel_dims(ctl_t const *ctl, atm_t *atm, obs_t const *obs, size_t *m_out) {
  if (m_out) *m_out = measurement_count(ctl, obs);
  return state_vector(ctl, atm, NULL);
}

void jr_b200_kernel(ctl_t const *ctl, atm_t *atm, obs_t *obs, double *k, size_t m, size_t n) {
  if (ctl->checkmode) { printf("# %s: no operation in checkmode\n", __func__); return; }
  jr_state_elem *el = (jr_state_elem *)malloc(sizeof(jr_state_elem) * (n ? n : 1));
  if (!el) JR_FATAL("Out of memory!");
  if (state_vector(ctl, atm, el) != n) JR_FATAL("jr_b200_kernel: n does not match the state vector of ctl/atm");
  int const nr = obs->nr, nd = ctl->nd, np = atm->np, nrow = ctl->ng + ctl->nw;
  int const private_p = ctl->hydz >= 0; /* the hydrostatic adjustment rewrites p of every perturbed atmosphere */

  /* undisturbed run: fills the caller's obs; its NaN mask is what the perturbed runs inherit (copy_obs, :168-195) */
  formod_one_impl(ctl, atm, obs);
  if (measurement_count(ctl, obs) != m) JR_FATAL("jr_b200_kernel: m does not match the number of finite radiances");
  memset(k, 0, sizeof(double) * m * n);

  size_t const chunk = 256; /* perturbed atmospheres per device batch */
  size_t const nb_max = n < chunk ? (n ? n : 1) : chunk;
  double *hs = (double *)malloc(sizeof(double) * nb_max);
  double *prof = (double *)malloc(sizeof(double) * nb_max * (size_t)np * 2);
  double **rows = (double **)malloc(sizeof(double *) * nb_max * (size_t)(nrow ? nrow : 1));
  double *out = (double *)malloc(sizeof(double) * nb_max * (size_t)nr * nd * 2);
  double *tp = (double *)malloc(sizeof(double) * nb_max * (size_t)nr * 3);
  jrb_atm_view *av = (jrb_atm_view *)malloc(sizeof(jrb_atm_view) * nb_max);
  jrb_obs_view *ov = (jrb_obs_view *)malloc(sizeof(jrb_obs_view) * nb_max);
  if (!hs || !prof || !rows || !out || !tp || !av || !ov) JR_FATAL("Out of memory!");

  for (size_t j0 = 0; j0 < n; j0 += chunk) {
    size_t const nb = (n - j0 < chunk) ? n - j0 : chunk;
    for (size_t b = 0; b < nb; b++) {
      jr_state_elem const *e = &el[j0 + b];
      double const x0 = e->value[e->ip];
      double h; /* perturbation sizes of the reference (src/jurassic.c:832-836) */
      if (e->iq == 0) h = fmax(fabs(0.01 * x0), 1e-7);
      else if (e->iq == 1) h = 1;
      else if (e->iq < 2 + ctl->ng) h = fmax(fabs(0.01 * x0), 1e-15);
      else h = 1e-4;
      hs[b] = h;
      double *mine = prof + b * (size_t)np * 2;
      memcpy(mine, e->value, sizeof(double) * (size_t)np);
      mine[e->ip] = x0 + h;
      fill_atm_view(atm, &av[b]);
      double **r = rows + b * (size_t)(nrow ? nrow : 1);
      for (int ig = 0; ig < ctl->ng; ig++) r[ig] = atm->q[ig];
      for (int iw = 0; iw < ctl->nw; iw++) r[ctl->ng + iw] = atm->k[iw];
      av[b].q_rows = r; av[b].k_rows = r + ctl->ng;
      if (e->iq == 0) av[b].p = mine;
      else if (e->iq == 1) av[b].t = mine;
      else r[e->iq - 2] = mine;
      if (private_p && e->iq != 0) { memcpy(mine + np, atm->p, sizeof(double) * (size_t)np); av[b].p = mine + np; }
      fill_obs_view(obs, &ov[b]);
      ov[b].rad = out + b * (size_t)nr * nd * 2; ov[b].tau = ov[b].rad + (size_t)nr * nd;
      ov[b].row_stride = nd; ov[b].nd_reset = nd;
      ov[b].tpz = tp + b * (size_t)nr * 3; ov[b].tplon = ov[b].tpz + nr; ov[b].tplat = ov[b].tplon + nr;
      for (int ir = 0; ir < nr; ir++) memcpy(ov[b].rad + (size_t)ir * nd, obs->rad[ir], sizeof(double) * (size_t)nd); /* NaN mask */
    }
    {
      jrb_ctl_view cv;
      fill_ctl_view(ctl, &cv);
      if (jrb_group_formod_batch(g_grp, &cv, (int)nb, av, ov, 0) != JRB_OK) JR_FATAL(jrb_group_last_error(g_grp));
    }
    for (size_t b = 0; b < nb; b++) { /* K[:, j] = (y1 - y0) / h over the finite radiances (obs2y order) */
      size_t i = 0;
      for (int ir = 0; ir < nr; ir++)
        for (int id = 0; id < nd; id++)
          if (isfinite(obs->rad[ir][id])) { k[i * n + (j0 + b)] = (ov[b].rad[(size_t)ir * nd + id] - obs->rad[ir][id]) / hs[b]; i++; }
    }
  }
  free(hs); free(prof); free(rows); free(out); free(tp); free(av); free(ov); free(el);
}

void jr_b200_finalize(void) {
  pthread_mutex_lock(&g_lock);
  if (g_grp) jrb_group_destroy(g_grp);
  g_grp = NULL; g_have_tables = 0; g_fov_loaded[0] = 0;
  pthread_mutex_unlock(&g_lock);
}

void jr_b200_dims(int dims[11], long long sizes[4]) {
  int const d[11] = {ND, NG, NP, NR, NW, NLOS, TBLNP, TBLNT, TBLNU, TBLNS, LEN};
  memcpy(dims, d, sizeof(d));
  sizes[0] = sizeof(ctl_t); sizes[1] = sizeof(atm_t); sizes[2] = sizeof(obs_t); sizes[3] = sizeof(tbl_t);
}

void *jr_b200_core_context(void) { return g_grp ? (void *)jrb_group_context(g_grp, 0, 0) : NULL; }
void *jr_b200_core_group(void) { return g_grp; }
