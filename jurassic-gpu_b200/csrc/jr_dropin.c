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
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

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

static jrb_context *g_ctx = NULL;
static int g_have_tables = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

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
}

static void fill_obs_view(obs_t *o, jrb_obs_view *v) {
  v->nr = o->nr;
  v->time = o->time; v->obsz = o->obsz; v->obslon = o->obslon; v->obslat = o->obslat;
  v->vpz = o->vpz; v->vplon = o->vplon; v->vplat = o->vplat;
  v->tpz = o->tpz; v->tplon = o->tplon; v->tplat = o->tplat;
  v->rad = &o->rad[0][0]; v->tau = &o->tau[0][0];
  v->row_stride = ND; v->nd_reset = ND;
}

/* control values can change between calls (the reference re-uploads ctl_t on every call, src/GPUdrivers.cu:355) */
static void push_control(ctl_t const *ctl) {
  jrb_ctl_view cv;
  fill_ctl_view(ctl, &cv);
  if (jrb_set_control(g_ctx, &cv) != JRB_OK) JR_FATAL(jrb_last_error(g_ctx));
}

static void init_locked(ctl_t const *ctl, tbl_t const *tbl, int device) {
  if (!g_ctx) {
    if (device < 0) device = ctl->MPIlocalrank;
    int const ndev = jrb_device_count();
    if (ndev < 1) JR_FATAL("no CUDA device available (there is no CPU fallback in this library)");
    if (device >= ndev) JR_FATAL("More MPI-Ranks on Node than GPUs. Abort."); /* src/GPUdrivers.cu:284-287 */
    if (jrb_create(&g_ctx, device) != JRB_OK) JR_FATAL(jrb_last_error(NULL));
  }
  push_control(ctl);
  jrb_tbl_view tv;
  fill_tbl_view(tbl, &tv);
  if (jrb_set_tables(g_ctx, &tv) != JRB_OK) JR_FATAL(jrb_last_error(g_ctx));
  g_have_tables = 1;
}

int jr_b200_init(ctl_t const *ctl, tbl_t const *tbl, int device) {
  if (!ctl || !tbl) JR_FATAL("jr_b200_init: NULL argument");
  pthread_mutex_lock(&g_lock);
  init_locked(ctl, tbl, device);
  pthread_mutex_unlock(&g_lock);
  return 0;
}

void jr_b200_formod_batch(ctl_t const *ctl, atm_t *const atm[], obs_t *const obs[], int npackages) {
  if (ctl->checkmode) { printf("# %s: no operation in checkmode\n", __func__); return; }
  if (npackages <= 0) return;
  pthread_mutex_lock(&g_lock); /* concurrent callers (OpenMP host threads of a retrieval) are serialised */
  if (!g_have_tables) {
    get_tbl_fn const gt = find_get_tbl();
    if (!gt) JR_FATAL("tables not initialised: call jr_b200_init() or link the reference's get_tbl()");
    init_locked(ctl, gt(ctl), -1);
  } else {
    push_control(ctl);
  }
  jrb_atm_view *av = (jrb_atm_view *)malloc(sizeof(jrb_atm_view) * (size_t)npackages);
  jrb_obs_view *ov = (jrb_obs_view *)malloc(sizeof(jrb_obs_view) * (size_t)npackages);
  if (!av || !ov) JR_FATAL("Out of memory!");
  for (int i = 0; i < npackages; i++) { fill_atm_view(atm[i], &av[i]); fill_obs_view(obs[i], &ov[i]); }
  if (jrb_formod_batch(g_ctx, npackages, av, ov) != JRB_OK) JR_FATAL(jrb_last_error(g_ctx));
  free(av); free(ov);
  pthread_mutex_unlock(&g_lock);
}

void formod_GPU(ctl_t const *ctl, atm_t *atm, obs_t *obs) {
  atm_t *const a[1] = {atm};
  obs_t *const o[1] = {obs};
  jr_b200_formod_batch(ctl, a, o, 1);
}

void jr_b200_finalize(void) {
  pthread_mutex_lock(&g_lock);
  if (g_ctx) jrb_destroy(g_ctx);
  g_ctx = NULL; g_have_tables = 0;
  pthread_mutex_unlock(&g_lock);
}

void jr_b200_dims(int dims[11], long long sizes[4]) {
  int const d[11] = {ND, NG, NP, NR, NW, NLOS, TBLNP, TBLNT, TBLNU, TBLNS, LEN};
  memcpy(dims, d, sizeof(d));
  sizes[0] = sizeof(ctl_t); sizes[1] = sizeof(atm_t); sizes[2] = sizeof(obs_t); sizes[3] = sizeof(tbl_t);
}

void *jr_b200_core_context(void) { return g_ctx; }
