/* ref_hooks.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Compiled TOGETHER WITH the unmodified reference sources (/root/reference/src/jurassic.c and
 * CPUdrivers.c, included where they lie, never copied) into oracle/_ref/libjurassic_ref_nd<ND>_ng<NG>.so.
 * It exports small helpers so that Python tests / bench.py can drive the reference:
 *   - struct sizes and field offsets (ABI test T1 of SURVEY.md section 7)
 *   - a flattening wrapper around the reference's own traceray()      (src/jr_common.h:585-711)
 *   - formod_CPU's call sequence with a caller-supplied tbl_t          (src/CPUdrivers.c:108-151)
 *     (the reference's get_tbl() can only load tables from files; big synthetic cases fill a tbl_t in memory)
 * Everything numerical is done by the reference's own functions; this file only forwards calls.
 */
/* The reference driver is compiled as part of this translation unit, unmodified, from where it lies
 * (jr_common.h uses C99 `extern inline`, so it can be included by exactly one C translation unit). */
#include "CPUdrivers.c"
#include <stddef.h>
#include <omp.h>

/* ---- ABI facts ---------------------------------------------------------------------------- */
void jrref_dims(int out[12]) {
  out[0] = ND; out[1] = NG; out[2] = NP; out[3] = NR; out[4] = NW; out[5] = NLOS;
  out[6] = TBLNP; out[7] = TBLNT; out[8] = TBLNU; out[9] = TBLNS; out[10] = LEN; out[11] = NSHAPE;
}

#define OFF(T, f) ((long long)offsetof(T, f))
/* fills "name=value" pairs into parallel arrays; returns count */
int jrref_layout(const char *names[], long long values[], int max) {
  int n = 0;
#define PUT(nm, v) do { if (n < max) { names[n] = nm; values[n] = (long long)(v); } n++; } while (0)
  PUT("sizeof(ctl_t)", sizeof(ctl_t)); PUT("sizeof(atm_t)", sizeof(atm_t)); PUT("sizeof(obs_t)", sizeof(obs_t));
  PUT("sizeof(pos_t)", sizeof(pos_t)); PUT("sizeof(tbl_t)", sizeof(tbl_t));
  PUT("ctl.ng", OFF(ctl_t, ng)); PUT("ctl.emitter", OFF(ctl_t, emitter)); PUT("ctl.nd", OFF(ctl_t, nd));
  PUT("ctl.nw", OFF(ctl_t, nw)); PUT("ctl.nu", OFF(ctl_t, nu)); PUT("ctl.window", OFF(ctl_t, window));
  PUT("ctl.tblbase", OFF(ctl_t, tblbase)); PUT("ctl.hydz", OFF(ctl_t, hydz)); PUT("ctl.ctm_co2", OFF(ctl_t, ctm_co2));
  PUT("ctl.ctm_h2o", OFF(ctl_t, ctm_h2o)); PUT("ctl.ctm_n2", OFF(ctl_t, ctm_n2)); PUT("ctl.ctm_o2", OFF(ctl_t, ctm_o2));
  PUT("ctl.ip", OFF(ctl_t, ip)); PUT("ctl.cz", OFF(ctl_t, cz)); PUT("ctl.cx", OFF(ctl_t, cx));
  PUT("ctl.refrac", OFF(ctl_t, refrac)); PUT("ctl.rayds", OFF(ctl_t, rayds)); PUT("ctl.raydz", OFF(ctl_t, raydz));
  PUT("ctl.fov", OFF(ctl_t, fov)); PUT("ctl.retp_zmin", OFF(ctl_t, retp_zmin)); PUT("ctl.retq_zmin", OFF(ctl_t, retq_zmin));
  PUT("ctl.retk_zmax", OFF(ctl_t, retk_zmax)); PUT("ctl.write_bbt", OFF(ctl_t, write_bbt));
  PUT("ctl.write_matrix", OFF(ctl_t, write_matrix)); PUT("ctl.formod", OFF(ctl_t, formod));
  PUT("ctl.rfmbin", OFF(ctl_t, rfmbin)); PUT("ctl.rfmhit", OFF(ctl_t, rfmhit)); PUT("ctl.rfmxsc", OFF(ctl_t, rfmxsc));
  PUT("ctl.useGPU", OFF(ctl_t, useGPU)); PUT("ctl.checkmode", OFF(ctl_t, checkmode));
  PUT("ctl.MPIglobrank", OFF(ctl_t, MPIglobrank)); PUT("ctl.MPIlocalrank", OFF(ctl_t, MPIlocalrank));
  PUT("ctl.read_binary", OFF(ctl_t, read_binary)); PUT("ctl.write_binary", OFF(ctl_t, write_binary));
  PUT("ctl.gpu_nbytes_shared_memory", OFF(ctl_t, gpu_nbytes_shared_memory));
  PUT("atm.time", OFF(atm_t, time)); PUT("atm.z", OFF(atm_t, z)); PUT("atm.lon", OFF(atm_t, lon)); PUT("atm.lat", OFF(atm_t, lat));
  PUT("atm.p", OFF(atm_t, p)); PUT("atm.t", OFF(atm_t, t)); PUT("atm.q", OFF(atm_t, q)); PUT("atm.k", OFF(atm_t, k));
  PUT("atm.np", OFF(atm_t, np)); PUT("atm.init", OFF(atm_t, init));
  PUT("obs.time", OFF(obs_t, time)); PUT("obs.obsz", OFF(obs_t, obsz)); PUT("obs.obslon", OFF(obs_t, obslon));
  PUT("obs.obslat", OFF(obs_t, obslat)); PUT("obs.vpz", OFF(obs_t, vpz)); PUT("obs.vplon", OFF(obs_t, vplon));
  PUT("obs.vplat", OFF(obs_t, vplat)); PUT("obs.tpz", OFF(obs_t, tpz)); PUT("obs.tplon", OFF(obs_t, tplon));
  PUT("obs.tplat", OFF(obs_t, tplat)); PUT("obs.tau", OFF(obs_t, tau)); PUT("obs.rad", OFF(obs_t, rad)); PUT("obs.nr", OFF(obs_t, nr));
  PUT("tbl.np", OFF(tbl_t, np)); PUT("tbl.nt", OFF(tbl_t, nt)); PUT("tbl.nu", OFF(tbl_t, nu)); PUT("tbl.p", OFF(tbl_t, p));
  PUT("tbl.t", OFF(tbl_t, t)); PUT("tbl.u", OFF(tbl_t, u)); PUT("tbl.eps", OFF(tbl_t, eps)); PUT("tbl.sr", OFF(tbl_t, sr));
  PUT("tbl.st", OFF(tbl_t, st));
#undef PUT
  return n;
}

/* ---- tables -------------------------------------------------------------------------------- */
tbl_t *jrref_tbl_calloc(void) { return (tbl_t *)calloc(1, sizeof(tbl_t)); } /* lazily-zero pages */
void jrref_tbl_free(tbl_t *t) { free(t); }
/* the reference's own loader (ASCII .tab/.filt files below ctl->tblbase), src/jurassic.c:311 */
void jrref_init_tbl(ctl_t const *ctl, tbl_t *tbl) { init_tbl(ctl, tbl); }
/* the reference's cached singleton (what formod() uses), src/jr_common.h:60 */
tbl_t *jrref_get_tbl(ctl_t const *ctl) { return get_tbl(ctl); }

/* ---- ray tracer ------------------------------------------------------------------------------ */
/* Runs the reference traceray() for ray ir and flattens pos_t[np] into
 * out[ip*stride + {0:z,1:lon,2:lat,3:p,4:t,5:ds, 6..6+NW-1:k, 6+NW..:q[ng], then u[ng]}], stride = 6+NW+2*ng.
 * Returns np; *tsurf as set by the reference; obs->tp* are updated in place. */
int jrref_traceray(ctl_t const *ctl, atm_t const *atm, obs_t *obs, int ir, double *out, double *tsurf) {
  pos_t *los = (pos_t *)malloc(sizeof(pos_t) * NLOS);
  int const np = traceray(ctl, atm, obs, ir, los, tsurf);
  int const ng = ctl->ng, stride = 6 + NW + 2 * ng;
  for (int ip = 0; ip < np; ip++) {
    double *o = out + (size_t)ip * stride;
    o[0] = los[ip].z; o[1] = los[ip].lon; o[2] = los[ip].lat; o[3] = los[ip].p; o[4] = los[ip].t; o[5] = los[ip].ds;
    for (int iw = 0; iw < NW; iw++) o[6 + iw] = los[ip].k[iw];
    for (int ig = 0; ig < ng; ig++) { o[6 + NW + ig] = los[ip].q[ig]; o[6 + NW + ng + ig] = los[ip].u[ig]; }
  }
  free(los);
  return np;
}

/* the reference's atmosphere interpolation dispatch (src/jurassic.c:685-691; 1-D, 2-D, 3-D); out = {p, t, q[ng], k[nw]}.
 * The 2-D / 3-D forms keep their geometry in function statics guarded by atm->init: pass a fresh atm (init == 0). */
void jrref_intpol_atm_geo(ctl_t const *ctl, atm_t *atm, double z0, double lon0, double lat0, double *out) {
  double q[NG], k[NW];
  intpol_atm_geo(ctl, atm, z0, lon0, lat0, &out[0], &out[1], q, k);
  for (int ig = 0; ig < ctl->ng; ig++) out[2 + ig] = q[ig];
  for (int iw = 0; iw < ctl->nw; iw++) out[2 + ctl->ng + iw] = k[iw];
}

/* ---- forward model with a caller-supplied table ------------------------------------------------ */
/* Same call sequence as formod_CPU (src/CPUdrivers.c:108-151), but with `tbl` given instead of get_tbl(),
 * and ig_co2/ig_h2o looked up on every call instead of cached in function statics (Appendix D #16). */
void jrref_formod_tbl(ctl_t const *ctl, atm_t *atm, obs_t *obs, tbl_t const *tbl) {
  char (*mask)[ND] = (char (*)[ND])malloc((size_t)NR * ND);
  save_mask(mask, obs, ctl);
  double *t_surf = (double *)malloc((size_t)obs->nr * sizeof(double));
  int *np = (int *)malloc((size_t)obs->nr * sizeof(int));
  pos_t (*los)[NLOS] = (pos_t (*)[NLOS])malloc((size_t)obs->nr * NLOS * sizeof(pos_t));
  int ig_co2 = -999, ig_h2o = -999;
  if (ctl->ctm_h2o) ig_h2o = find_emitter(ctl, "H2O");
  if (ctl->ctm_co2) ig_co2 = find_emitter(ctl, "CO2");
  char const fourbit = (char)(((1 == ctl->ctm_co2) && (ig_co2 >= 0)) * 8 + ((1 == ctl->ctm_h2o) && (ig_h2o >= 0)) * 4 +
                              (1 == ctl->ctm_n2) * 2 + (1 == ctl->ctm_o2) * 1);
  hydrostatic1d_CPU(ctl, atm, obs->nr, ig_h2o);
  raytrace_rays_CPU(ctl, atm, obs, los, t_surf, np); /* outside a parallel region, i.e. serial, as in the reference (:136) */
#pragma omp parallel
  {
    apply_kernels_CPU(tbl, ctl, obs, los, np, ig_co2, ig_h2o, fourbit);
    surface_terms_CPU(tbl, obs, t_surf, ctl->nd);
  }
  free(los); free(np); free(t_surf);
  if (ctl->write_bbt) radiance_to_brightness_CPU(ctl, obs);
  apply_mask(mask, obs, ctl);
  free(mask);
}

/* unmodified entry point, tables from files via get_tbl() */
void jrref_formod(ctl_t const *ctl, atm_t *atm, obs_t *obs) { formod(ctl, atm, obs); }

/* the reference's finite-difference Jacobian kernel() (src/jurassic.c:812-857), tables via get_tbl(); k is m x n row-major */
size_t jrref_kernel_dims(ctl_t const *ctl, atm_t const *atm, obs_t const *obs, size_t *m) {
  *m = obs2y(ctl, obs, NULL, NULL, NULL);
  return atm2x(ctl, atm, NULL, NULL, NULL);
}
void jrref_kernel(ctl_t const *ctl, atm_t *atm, obs_t *obs, double *k, size_t m, size_t n) {
  gsl_matrix *K = gsl_matrix_alloc(m, n);
  kernel(ctl, atm, obs, K);
  for (size_t i = 0; i < m; i++) for (size_t j = 0; j < n; j++) k[i * n + j] = gsl_matrix_get(K, i, j);
  gsl_matrix_free(K);
}

int jrref_max_threads(void) { return omp_get_max_threads(); }
void jrref_set_threads(int n) { omp_set_num_threads(n); }
