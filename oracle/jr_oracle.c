/* jr_oracle.c -- TEST INFRASTRUCTURE ONLY: scalar CPU restatement of the reference's EGA forward-model path.
 *
 * Purpose: an independent checker for the CUDA path that also exists on machines without the reference checkout.
 * It is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it.
 * It restates, function by function, what the reference computes (paths below /root/reference/src), operating on
 * the dimension-agnostic views of include/jurassic_b200.h instead of the compile-time sized structs:
 *
 *   jro_formod        formod_CPU            CPUdrivers.c:108-151  (mask, hydrostatic, raytrace, kernels, surface, BT)
 *   o_traceray        traceray              jr_common.h:585-711
 *   o_ega_eps         ega_eps               jr_common.h:237-268   (+ locate_id :106, locate_tbl_id :116, get_u :179, get_eps :156)
 *   o_continua        continua_core_bbbb    jr_continua_core.mv4g.h:1-14, jr_common.h:315-390
 *   o_planck          src_planck_core       jr_common.h:220-224
 *   o_hydrostatic     hydrostatic_1d_h2o    jr_common.h:714-761
 *
 * Pinning: validated against the reference itself (oracle/_ref, compiled unmodified from /root/reference/src) by
 * tests/test_oracle_vs_reference.py on the limb/nadir examples and synthetic cases, and against the geometry
 * columns of the reference's golden files example/{limb,nadir}/rad.org.  The radiance columns of rad.org cannot be
 * reproduced by anyone here: their emissivity tables are missing from the reference checkout.
 */
#include "jr_oracle.h"
#include "jrb_ctm_data.h" /* generated at build time from the reference's src/ctm*.tbl */

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define O_C1 1.19104259e-8 /* jurassic.h:111 */
#define O_C2 1.43877506    /* :114 */
#define O_P0 1013.25       /* :120 */
#define O_RE 6367.421      /* :126 */
#define O_KB 1.3806504e-23 /* GSL_CONST_MKSA_BOLTZMANN, GSL 2.5 */
#define O_NA 6.02214199e23 /* GSL_CONST_NUM_AVOGADRO */
#define O_RGAS 8.314472    /* GSL_CONST_MKSA_MOLAR_GAS */
#define O_NLOS 400

typedef struct { double z, lon, lat, p, t, ds; double *q, *k, *u; } o_pos;

static double o_c01(double x) { return (x > 1.) ? 1. : ((x < 0.) ? 0. : x); }                   /* :43-45 */
static double o_lip(double x0, double y0, double x1, double y1, double x) {                   /* :48-50 */
  return y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}
static double o_eip(double x0, double y0, double x1, double y1, double x) {                   /* :53-57 */
  if ((y0 > 0) && (y1 > 0)) return y0 * exp(log(y1 / y0) / (x1 - x0) * (x - x0));
  return o_lip(x0, y0, x1, y1, x);
}

/* locate (:87-104): ascending or descending */
static int o_locate(const double *xx, int n, double x) {
  int ilo = 0, ihi = n - 1, i = (n - 1) >> 1;
  if (xx[i] < xx[i + 1]) {
    while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (xx[i] > x) ihi = i; else ilo = i; }
  } else {
    while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (xx[i] <= x) ihi = i; else ilo = i; }
  }
  return ilo;
}

/* ---- table access through the view ----------------------------------------------------------------------- */
#define TI_NP(v, g, d) ((v)->np[(size_t)(g) * (v)->dim_d + (d)])
#define TI_NT(v, g, ip, d) ((v)->nt[((size_t)(g) * (v)->dim_p + (ip)) * (v)->dim_d + (d)])
#define TI_NU(v, g, ip, it, d) ((v)->nu[(((size_t)(g) * (v)->dim_p + (ip)) * (v)->dim_t + (it)) * (v)->dim_d + (d)])
#define TI_P(v, g, ip, d) ((v)->p[((size_t)(g) * (v)->dim_p + (ip)) * (v)->dim_d + (d)])
#define TI_T(v, g, ip, it, d) ((v)->t[(((size_t)(g) * (v)->dim_p + (ip)) * (v)->dim_t + (it)) * (v)->dim_d + (d)])
#define TI_COL(v, g, ip, it, d) ((((((size_t)(g) * (v)->dim_p + (ip)) * (v)->dim_t + (it)) * (v)->dim_u) * (v)->dim_d) + (d))

/* locate_tbl_id (:116-125) on a float column with stride dim_d */
static int o_locate_col(const float *col, size_t stride, int n, double x) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) { int i = (ihi + ilo) >> 1; if (col[(size_t)i * stride] > x) ihi = i; else ilo = i; }
  return ilo;
}
static double o_get_eps(const jrb_tbl_view *v, int ig, int id, int ip, int it, double u) {      /* :156-177 */
  const size_t c = TI_COL(v, ig, ip, it, id), s = v->dim_d;
  const int idx = o_locate_col(v->u + c, s, TI_NU(v, ig, ip, it, id), u);
  return o_lip(v->u[c + idx * s], v->eps[c + idx * s], v->u[c + (idx + 1) * s], v->eps[c + (idx + 1) * s], u);
}
static double o_get_u(const jrb_tbl_view *v, int ig, int id, int ip, int it, double eps) {      /* :179-185 */
  const size_t c = TI_COL(v, ig, ip, it, id), s = v->dim_d;
  const int idx = o_locate_col(v->eps + c, s, TI_NU(v, ig, ip, it, id), eps);
  return o_lip(v->eps[c + idx * s], v->u[c + idx * s], v->eps[c + (idx + 1) * s], v->u[c + (idx + 1) * s], eps);
}
/* locate_id (:106-114) on the p axis / a T axis of one (gas, channel) */
static int o_locate_p(const jrb_tbl_view *v, int ig, int id, int n, double x) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) { int i = (ihi + ilo) >> 1; if (TI_P(v, ig, i, id) > x) ihi = i; else ilo = i; }
  return ilo;
}
static int o_locate_t(const jrb_tbl_view *v, int ig, int ip, int id, int n, double x) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) { int i = (ihi + ilo) >> 1; if (TI_T(v, ig, ip, i, id) > x) ihi = i; else ilo = i; }
  return ilo;
}

/* ega_eps (:237-268) */
static double o_ega_eps(const jrb_tbl_view *v, double tau, double t, double u, double p, int ig, int id) {
  if (tau < 1e-9) return 0.;
  if (TI_NP(v, ig, id) < 2) return 1.;
  const int ipr = o_locate_p(v, ig, id, TI_NP(v, ig, id), p);
  if (TI_NT(v, ig, ipr, id) < 2 || TI_NT(v, ig, ipr + 1, id) < 2) return 1.;
  const int it0 = o_locate_t(v, ig, ipr, id, TI_NT(v, ig, ipr, id), t);
  if (TI_NU(v, ig, ipr, it0, id) < 2 || TI_NU(v, ig, ipr, it0 + 1, id) < 2) return 1.;
  const int it1 = o_locate_t(v, ig, ipr + 1, id, TI_NT(v, ig, ipr + 1, id), t);
  if (TI_NU(v, ig, ipr + 1, it1, id) < 2 || TI_NU(v, ig, ipr + 1, it1 + 1, id) < 2) return 1.;
  const double eps = 1 - tau;
  const double u00 = o_get_u(v, ig, id, ipr, it0, eps), u01 = o_get_u(v, ig, id, ipr, it0 + 1, eps);
  const double u10 = o_get_u(v, ig, id, ipr + 1, it1, eps), u11 = o_get_u(v, ig, id, ipr + 1, it1 + 1, eps);
  const double e00 = o_c01(o_get_eps(v, ig, id, ipr, it0, u00 + u)), e01 = o_c01(o_get_eps(v, ig, id, ipr, it0 + 1, u01 + u));
  const double e10 = o_c01(o_get_eps(v, ig, id, ipr + 1, it1, u10 + u)), e11 = o_c01(o_get_eps(v, ig, id, ipr + 1, it1 + 1, u11 + u));
  const double ep0 = o_c01(o_lip(TI_T(v, ig, ipr, it0, id), e00, TI_T(v, ig, ipr, it0 + 1, id), e01, t));
  const double ep1 = o_c01(o_lip(TI_T(v, ig, ipr + 1, it1, id), e10, TI_T(v, ig, ipr + 1, it1 + 1, id), e11, t));
  const double ept = o_c01(o_lip(TI_P(v, ig, ipr, id), ep0, TI_P(v, ig, ipr + 1, id), ep1, p));
  return (1. - ept) / tau;
}

/* ---- continua (:315-390) ------------------------------------------------------------------------------------ */
static double o_ctmco2(double nu, double p, double t, double u) {
  if (nu < 0 || nu >= 4000) return 0;
  const double xw = nu * 0.5 + 1;
  const int iw = (int)xw;
  const double dw = xw - iw, ew = 1 - dw;
  const double cw296 = ew * jrb_co2296[iw - 1] + dw * jrb_co2296[iw];
  const double cw260 = ew * jrb_co2260[iw - 1] + dw * jrb_co2260[iw];
  const double cw230 = ew * jrb_co2230[iw - 1] + dw * jrb_co2230[iw];
  const double dt230 = t - 230, dt260 = t - 260, dt296 = t - 296;
  const double ctw = dt260 * 5.050505e-4 * dt296 * cw230 - dt230 * 9.259259e-4 * dt296 * cw260 + dt230 * 4.208754e-4 * dt260 * cw296;
  return u * p * ctw / (O_NA * 1000 * O_P0);
}
static double o_ctmh2o(double nu, double p, double t, double q, double u) {
  if (nu < 0 || nu >= 20000) return 0;
  const double xw = nu / 10 + 1;
  const int iw = (int)xw;
  const double dw = xw - iw, ew = 1 - dw;
  const double cw296 = ew * jrb_h2o296[iw - 1] + dw * jrb_h2o296[iw];
  const double cw260 = ew * jrb_h2o260[iw - 1] + dw * jrb_h2o260[iw];
  const double cwfrn = ew * jrb_h2ofrn[iw - 1] + dw * jrb_h2ofrn[iw];
  double sfac = 1.;
  if ((nu > 820.) && (nu < 960.)) {
    const char xfcrev[16] = {3, 9, 15, 23, 29, 33, 37, 39, 40, 46, 36, 27, 10, 2, 0, 0};
    const float xx = nu * 0.1 - 82;
    const int ix = (int)xx;
    const float dx = xx - ix;
    sfac += .001 * ((1 - dx) * xfcrev[ix] + dx * xfcrev[ix + 1]);
  }
  const double ctwslf = sfac * cw296 * pow(cw260 / cw296, (296. - t) / (296. - 260.));
  const double vf1 = nu - 370.;
  const double vf2 = vf1 * vf1;
  const double vf6 = vf2 * vf2 * vf2;
  const double fscal = 36100. / (vf2 + vf6 * 1e-8 + 36100.) * -.25 + 1.;
  const double ctwfrn = cwfrn * fscal;
  const double a1 = nu * u * tanh(.7193876 / t * nu);
  const double a2 = 296. / t;
  const double a3 = p / O_P0 * (q * ctwslf + (1 - q) * ctwfrn) * 1e-20;
  return a1 * a2 * a3;
}
static double o_ctmn2(double nu, double p, double t) {
  if (nu < 2120 || nu > 2605) return 0;
  const double xnu = nu * 0.2 - 424;
  const int idx = (int)xnu;
  const double a1 = xnu - idx, a0 = 1 - a1;
  /* the reference reads one past the end at nu == 2605 with weight a1 == 0 */
  const double b = a0 * jrb_n2_ba[idx] + a1 * (idx + 1 < 98 ? jrb_n2_ba[idx + 1] : 0.);
  const double beta = a0 * jrb_n2_betaa[idx] + a1 * (idx + 1 < 98 ? jrb_n2_betaa[idx + 1] : 0.);
  const double q_n2 = 0.79, t0 = 273, tr = 296;
  return 0.1 * (p / O_P0) * (p / O_P0) * (t0 / t) * (t0 / t) * exp(beta * (1 / tr - 1 / t)) * q_n2 * b *
         (q_n2 + (1 - q_n2) * (1.294 - 0.4545 * t / tr));
}
static double o_ctmo2(double nu, double p, double t) {
  if (nu < 1360 || nu > 1805) return 0;
  const double xnu = nu * 0.2 - 272;
  const int idx = (int)xnu;
  const double a1 = xnu - idx, a0 = 1 - a1;
  const double b = a0 * jrb_o2_ba[idx] + a1 * (idx + 1 < 90 ? jrb_o2_ba[idx + 1] : 0.);
  const double beta = a0 * jrb_o2_betaa[idx] + a1 * (idx + 1 < 90 ? jrb_o2_betaa[idx + 1] : 0.);
  const double q_o2 = 0.21, t0 = 273, tr = 296;
  return 0.1 * (p / O_P0) * (p / O_P0) * (t0 / t) * (t0 / t) * exp(beta * (1 / tr - 1 / t)) * q_o2 * b;
}
/* continua_core_bbbb (jr_continua_core.mv4g.h:1-14) */
static double o_continua(const jrb_ctl_view *c, int fourbit, const o_pos *los, int id) {
  double beta_ds = los->k[c->window[id]] * los->ds;
  if (fourbit & 8) beta_ds += o_ctmco2(c->nu[id], los->p, los->t, los->u[c->ig_co2]);
  if (fourbit & 4) beta_ds += o_ctmh2o(c->nu[id], los->p, los->t, los->q[c->ig_h2o], los->u[c->ig_h2o]);
  if (fourbit & 2) beta_ds += o_ctmn2(c->nu[id], los->p, los->t) * los->ds;
  if (fourbit & 1) beta_ds += o_ctmo2(c->nu[id], los->p, los->t) * los->ds;
  return beta_ds;
}

/* src_planck_core (:220-224) with locate_st (:82-84) */
static double o_planck(const jrb_tbl_view *v, double t, int id) {
  const int it = (int)(4 * t) - 400;
  return o_lip(v->st[it], v->sr[(size_t)it * v->dim_d + id], v->st[it + 1], v->sr[(size_t)(it + 1) * v->dim_d + id], t);
}

/* ---- geometry (:482-500) -------------------------------------------------------------------------------------- */
static double o_norm(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static void o_cart2geo(const double x[3], double *alt, double *lon, double *lat) {
  const double radius = o_norm(x);
  *lat = asin(x[2] / radius) * (180 / M_PI);
  *lon = atan2(x[1], x[0]) * (180 / M_PI);
  *alt = radius - O_RE;
}
static void o_geo2cart(double alt, double lon, double lat, double x[3]) {
  const double radius = alt + O_RE, clat = cos(lat * (M_PI / 180));
  x[0] = radius * clat * cos(lon * (M_PI / 180));
  x[1] = radius * clat * sin(lon * (M_PI / 180));
  x[2] = radius * sin(lat * (M_PI / 180));
}
static void o_intpol_pt(const jrb_atm_view *a, int idx0, int n, double z0, double *p, double *t) { /* :549-555 */
  const int ip = idx0 + o_locate(a->z + idx0, n, z0);
  *p = o_eip(a->z[ip], a->p[ip], a->z[ip + 1], a->p[ip + 1], z0);
  *t = o_lip(a->z[ip], a->t[ip], a->z[ip + 1], a->t[ip + 1], z0);
}
static void o_intpol_qk(const jrb_ctl_view *c, const jrb_atm_view *a, int idx0, int n, double z0, double *q, double *k) { /* :557-567 */
  const int ip = idx0 + o_locate(a->z + idx0, n, z0);
  for (int ig = 0; ig < c->ng; ig++) {
    const double *qq = a->q_rows ? a->q_rows[ig] : a->q + (size_t)ig * a->q_stride;
    q[ig] = o_lip(a->z[ip], qq[ip], a->z[ip + 1], qq[ip + 1], z0);
  }
  for (int iw = 0; iw < c->nw; iw++) {
    const double *kk = a->k_rows ? a->k_rows[iw] : a->k + (size_t)iw * a->k_stride;
    k[iw] = o_lip(a->z[ip], kk[ip], a->z[ip + 1], kk[ip + 1], z0);
  }
}

/* ---- 2-D / 3-D atmosphere interpolation (src/jurassic.c:685-804) ---------------------------------------------------
 * intpol_atm_geo dispatches on ctl->ip.  The reference's formod() path never gets there: its tracer calls
 * intpol_atm_geo_pt/_qk, which assert ip == 1 (src/jr_common.h:569-583).  The restatement below is what the dispatch of
 * src/jurassic.c:685-691 gives when it is applied to the atmosphere slice [idx0, idx0+n) that locate_atm selected for the
 * ray -- i.e. the tracer of src/jr_common.h:585-711 with its two assert-guarded calls replaced by intpol_atm_geo.
 * The interpolation itself is pinned bit for bit against the reference's intpol_atm_geo (tests/test_oracle_vs_reference.py);
 * the composition with the tracer has no reference behaviour to compare with (the reference aborts).
 * The reference keeps the profile list of the 2-D case in function statics guarded by atm->init; here it is rebuilt per
 * call (same values).  Return: 0, or the reference's fatal conditions: -3 "Cannot identify profiles. Check ordering of
 * data points!", -4 "Distance of profiles is too large!" (:727-728). */
static void o_intpol_1d(const jrb_ctl_view *c, const jrb_atm_view *a, int idx0, int n, double z0, double *p, double *t, double *q, double *k) {
  o_intpol_pt(a, idx0, n, z0, p, t);      /* intpol_atm_1d (:694-701): EXP / LIN macros = eip / lip */
  o_intpol_qk(c, a, idx0, n, z0, q, k);
}
static double o_dist2(const double a[3], const double b[3]) { /* DIST2, src/jurassic.h:61 */
  return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}
static int o_intpol_2d(const jrb_ctl_view *c, const jrb_atm_view *a, int idx0, int n, double z0, double lon0, double lat0,
                       double *p, double *t, double *q, double *k) { /* intpol_atm_2d (:704-760) */
  const double dlat = 10;
  double dhmin0 = 1e99, dhmin1 = 1e99, lat1 = -999, lon1 = -999, x0[3], x1a[3] = {0, 0, 0}, x1b[3] = {0, 0, 0};
  int nx = 0, ix0 = 0, ix1 = 0;
  int *idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1) * 2), *nz = idx + (n > 0 ? n : 1);
  for (int ip = idx0; ip < idx0 + n; ip++) { /* profile list (:713-725) */
    if ((a->lon[ip] != lon1) || (a->lat[ip] != lat1)) {
      ++nx; nz[nx - 1] = 0; lon1 = a->lon[ip]; lat1 = a->lat[ip]; idx[nx - 1] = ip;
    }
    ++nz[nx - 1];
  }
  for (int ix = 0; ix < nx; ix++) {
    if (nz[ix] <= 1) { free(idx); return -3; }
    if ((ix > 0) && (fabs(a->lat[idx[ix - 1]] - a->lat[idx[ix]]) > dlat)) { free(idx); return -4; }
  }
  o_geo2cart(0, lon0, lat0, x0);
  for (int ix = 0; ix < nx; ix++) /* two nearest profiles (:732-745) */
    if (fabs(lat0 - a->lat[idx[ix]]) <= dlat) {
      double xp[3];
      o_geo2cart(0, a->lon[idx[ix]], a->lat[idx[ix]], xp);
      const double dh = o_dist2(x0, xp);
      if (dh <= dhmin0) { dhmin1 = dhmin0; ix1 = ix0; dhmin0 = dh; ix0 = ix; }
      else if (dh <= dhmin1) { dhmin1 = dh; ix1 = ix; }
    }
  double p0, p1, t0, t1, q0[JRB_MAX_NG], q1[JRB_MAX_NG], k0[JRB_MAX_NW], k1[JRB_MAX_NW], r;
  o_intpol_1d(c, a, idx[ix0], nz[ix0], z0, &p0, &t0, q0, k0);
  o_intpol_1d(c, a, idx[ix1], nz[ix1], z0, &p1, &t1, q1, k1);
  o_geo2cart(0, a->lon[idx[ix0]], a->lat[idx[ix0]], x1a);
  o_geo2cart(0, a->lon[idx[ix1]], a->lat[idx[ix1]], x1b);
  const double x2 = o_dist2(x1a, x1b), x = sqrt(x2), r0 = (dhmin0 - dhmin1 + x2) / (2 * x), r1 = x - r0; /* :750-755 */
  if (r0 <= 0) r = 0; else r = (r1 <= 0) ? 1 : r0 / (r0 + r1);
  *p = (1 - r) * p0 + r * p1;
  *t = (1 - r) * t0 + r * t1;
  for (int ig = 0; ig < c->ng; ig++) q[ig] = (1 - r) * q0[ig] + r * q1[ig];
  for (int iw = 0; iw < c->nw; iw++) k[iw] = (1 - r) * k0[iw] + r * k1[iw];
  free(idx);
  return 0;
}
static int o_intpol_3d(const jrb_ctl_view *c, const jrb_atm_view *a, int idx0, int n, double z0, double lon0, double lat0,
                       double *p, double *t, double *q, double *k) { /* intpol_atm_3d (:763-804) */
  const double rm2 = c->cx * c->cx;
  double wsum = 0, x0[3];
  *p = *t = 0.;
  for (int ig = 0; ig < c->ng; ig++) q[ig] = 0;
  for (int iw = 0; iw < c->nw; iw++) k[iw] = 0;
  for (int ip = idx0; ip < idx0 + n; ip++) {
    const double dz = fabs(a->z[ip] - z0);
    if (dz >= c->cz) continue;
    if (fabs(a->lat[ip] - lat0) * 111.13 >= c->cx) continue;
    double xp[3];
    o_geo2cart(0, lon0, lat0, x0);
    o_geo2cart(0, a->lon[ip], a->lat[ip], xp);
    const double dx2 = o_dist2(x0, xp);
    if (dx2 >= rm2) continue;
    const double w = (1 - dz / c->cz) * (rm2 - dx2) / (rm2 + dx2);
    wsum += w;
    *p += w * a->p[ip];
    *t += w * a->t[ip];
    for (int ig = 0; ig < c->ng; ig++) q[ig] += w * (a->q_rows ? a->q_rows[ig] : a->q + (size_t)ig * a->q_stride)[ip];
    for (int iw = 0; iw < c->nw; iw++) k[iw] += w * (a->k_rows ? a->k_rows[iw] : a->k + (size_t)iw * a->k_stride)[ip];
  }
  if (wsum >= 1e-6) {
    *p /= wsum; *t /= wsum;
    for (int ig = 0; ig < c->ng; ig++) q[ig] /= wsum;
    for (int iw = 0; iw < c->nw; iw++) k[iw] /= wsum;
  } else {
    *p = *t = NAN;
    for (int ig = 0; ig < c->ng; ig++) q[ig] = NAN;
    for (int iw = 0; iw < c->nw; iw++) k[iw] = NAN;
  }
  return 0;
}
/* intpol_atm_geo (:685-691) on a slice */
static int o_intpol_geo(const jrb_ctl_view *c, const jrb_atm_view *a, int idx0, int n, double z0, double lon0, double lat0,
                        double *p, double *t, double *q, double *k) {
  if (c->ip == 1) { o_intpol_1d(c, a, idx0, n, z0, p, t, q, k); return 0; }
  if (c->ip == 2) return o_intpol_2d(c, a, idx0, n, z0, lon0, lat0, p, t, q, k);
  if (c->ip == 3) return o_intpol_3d(c, a, idx0, n, z0, lon0, lat0, p, t, q, k);
  return -1; /* "Unknown interpolation method, check IP!" */
}
/* the dispatch applied to the whole atmosphere, as the reference's intpol_atm_geo does; out = {p, t, q[ng], k[nw]} */
int jro_intpol_atm_geo(const jrb_ctl_view *c, const jrb_atm_view *a, double z0, double lon0, double lat0, double *out) {
  double q[JRB_MAX_NG], k[JRB_MAX_NW];
  const int rc = o_intpol_geo(c, a, 0, a->np, z0, lon0, lat0, &out[0], &out[1], q, k);
  for (int ig = 0; ig < c->ng; ig++) out[2 + ig] = q[ig];
  for (int iw = 0; iw < c->nw; iw++) out[2 + c->ng + iw] = k[iw];
  return rc;
}

/* traceray (:585-711).  los[] must hold O_NLOS points whose q/k/u arrays are allocated.  Returns np (or < 0: the fatal
 * conditions of the 2-D interpolation, see above). */
static int o_traceray(const jrb_ctl_view *c, const jrb_atm_view *a, const jrb_obs_view *o, int ir, o_pos *los, double *tsurf) {
  double ex0[3], ex1[3], q[JRB_MAX_NG], k[JRB_MAX_NW], lat, lon, p, t, x[3], xobs[3], xvp[3], z = 1e99, z_low = z, zmax, zmin;
  const double zrefrac = 60;
  *tsurf = -999;
  for (int ig = 0; ig < JRB_MAX_NG; ig++) q[ig] = 0;
  for (int iw = 0; iw < JRB_MAX_NW; iw++) k[iw] = 0;
  o->tpz[ir] = o->vpz[ir]; o->tplon[ir] = o->vplon[ir]; o->tplat[ir] = o->vplat[ir];
  /* locate_atm (:127-154) */
  int lo = 0, hi = a->np - 1, i;
  while (hi > lo + 1) { i = (lo + hi) / 2; if (a->time[i] < o->time[ir]) lo = i; else hi = i; }
  const int lower = (0 == lo) ? lo : hi;
  lo = lower; hi = a->np - 1;
  while (hi > lo + 1) { i = (lo + hi) / 2; if (a->time[i] > o->time[ir]) hi = i; else lo = i; }
  const int upper = (hi == a->np - 1) ? a->np : hi;
  const int idx0 = lower, n = upper - lower;
  /* altitude_range_nn (:411-420) */
  zmax = zmin = a->z[idx0];
  for (int ipp = idx0; (ipp < idx0 + n) && (a->lon[ipp] == a->lon[idx0]) && (a->lat[ipp] == a->lat[idx0]); ++ipp) {
    zmax = fmax(zmax, a->z[ipp]); zmin = fmin(zmin, a->z[ipp]);
  }
  if (o->obsz[ir] < zmin) return 0;
  if (o->vpz[ir] > zmax - 0.001) return 0;
  o_geo2cart(o->obsz[ir], o->obslon[ir], o->obslat[ir], xobs);
  o_geo2cart(o->vpz[ir], o->vplon[ir], o->vplat[ir], xvp);
  for (i = 0; i < 3; i++) ex0[i] = xvp[i] - xobs[i];
  const double norm = o_norm(ex0);
  for (i = 0; i < 3; i++) { ex0[i] /= norm; x[i] = xobs[i]; }
  if (o->obsz[ir] > zmax) {
    double dmax = norm, dmin = 0.;
    while (fabs(dmin - dmax) > 0.001) {
      const double d = 0.5 * (dmax + dmin);
      for (i = 0; i < 3; i++) x[i] = xobs[i] + d * ex0[i];
      z = o_norm(x) - O_RE;
      if ((z <= zmax) && (z > zmax - 0.001)) break;
      if (z < zmax - 0.0005) dmax = d; else dmin = d;
    }
  }
  int np = 0, z_low_idx = -1;
  double qd[JRB_MAX_NG], kd[JRB_MAX_NW]; /* the probe points need p and T only */
#define O_PROBE_PT() do { if (c->ip == 1) o_intpol_pt(a, idx0, n, z, &p, &t); \
                          else { const int rc_ = o_intpol_geo(c, a, idx0, n, z, lon, lat, &p, &t, qd, kd); if (rc_) return rc_; } } while (0)
  for (int stop = 0; np < O_NLOS; ++np) {
    double ds = c->rayds, dz = c->raydz;
    if (dz > 0.) {
      const double norm_x = 1.0 / o_norm(x);
      double dot = 0.;
      for (i = 0; i < 3; i++) dot += ex0[i] * x[i] * norm_x;
      const double cosa = fabs(dot);
      if (cosa != 0.) ds = fmin(ds, dz / cosa);
    }
    o_cart2geo(x, &z, &lon, &lat);
    if ((z < zmin) || (z > zmax)) {
      if (np == 0) return 0; /* the reference would read los[-1] here (undefined); unreachable after the entry search */
      double xh[3];
      stop = (z < zmin) ? 2 : 1;
      o_geo2cart(los[np - 1].z, los[np - 1].lon, los[np - 1].lat, xh);
      const double zfrac = (z < zmin) ? zmin : zmax;
      const double frac = (zfrac - los[np - 1].z) / (z - los[np - 1].z);
      for (i = 0; i < 3; i++) x[i] = xh[i] + frac * (x[i] - xh[i]);
      o_cart2geo(x, &z, &lon, &lat);
      los[np - 1].ds = ds * frac;
      ds = 0.;
    }
    if (c->ip == 1) { o_intpol_pt(a, idx0, n, z, &p, &t); o_intpol_qk(c, a, idx0, n, z, q, k); }
    else { const int rc = o_intpol_geo(c, a, idx0, n, z, lon, lat, &p, &t, q, k); if (rc) return rc; }
    los[np].lon = lon; los[np].lat = lat; los[np].z = z; los[np].p = p; los[np].t = t; los[np].ds = ds; /* write_pos_point :422-434 */
    for (int ig = 0; ig < c->ng; ig++) los[np].q[ig] = q[ig];
    for (int iw = 0; iw < c->nw; iw++) los[np].k[iw] = k[iw];
    if (z < z_low) { z_low = z; z_low_idx = np; }
    if (stop) { *tsurf = (stop == 2 ? t : -999); break; }
    double nn = 1., ng[] = {0., 0., 0.};
    if (c->refrac && z <= zrefrac) {
      nn += 7.753e-05 * p / t;
      double xh[3];
      for (i = 0; i < 3; i++) xh[i] = x[i] + 0.5 * ds * ex0[i];
      o_cart2geo(xh, &z, &lon, &lat);
      O_PROBE_PT();
      const double n2 = 7.753e-05 * p / t;
      for (i = 0; i < 3; i++) {
        const double h = 0.02;
        xh[i] += h;
        o_cart2geo(xh, &z, &lon, &lat);
        O_PROBE_PT();
        ng[i] = (7.753e-05 * p / t - n2) / h;
        xh[i] -= h;
      }
    }
    for (i = 0; i < 3; i++) ex1[i] = ex0[i] * nn + ds * ng[i];
    const double norm_ex1 = o_norm(ex1);
    for (i = 0; i < 3; i++) { ex1[i] /= norm_ex1; x[i] += 0.5 * ds * (ex0[i] + ex1[i]); ex0[i] = ex1[i]; }
  }
  ++np;
  if (np > O_NLOS) np = O_NLOS; /* reference: fatal "Too many LOS points!" when np >= NLOS (:693-695) */

  /* tangent_point (:502-539) */
  {
    const int ip = z_low_idx;
    if (ip <= 0 || ip >= np - 1) {
      o->tpz[ir] = los[np - 1].z; o->tplon[ir] = los[np - 1].lon; o->tplat[ir] = los[np - 1].lat;
    } else {
      const double yy0 = los[ip - 1].z, yy1 = los[ip].z, yy2 = los[ip + 1].z, ds0 = los[ip].ds, ds1 = los[ip + 1].ds,
                   dyy10 = yy1 - yy0, dyy21 = yy2 - yy1, x1 = sqrt(ds0 * ds0 - dyy10 * dyy10),
                   x2 = x1 + sqrt(ds1 * ds1 - dyy21 * dyy21), dx12 = x1 - x2,
                   aa = (dyy10 * x2 + (yy0 - yy2) * x1) / (x1 * x2 * dx12), bb = dyy10 / x1 - aa * x1, cc = yy0,
                   xx = -bb / (2 * aa);
      o->tpz[ir] = (aa * xx + bb) * xx + cc;
      double v[3], v0[3], v2[3], dummy;
      o_geo2cart(los[ip - 1].z, los[ip - 1].lon, los[ip - 1].lat, v0);
      o_geo2cart(los[ip + 1].z, los[ip + 1].lon, los[ip + 1].lat, v2);
      for (i = 0; i < 3; i++) v[i] = o_lip(0.0, v0[i], x2, v2[i], xx);
      o_cart2geo(v, &dummy, &o->tplon[ir], &o->tplat[ir]);
    }
  }
  for (int ip = np - 1; ip >= 1; ip--) los[ip].ds = 0.5 * (los[ip - 1].ds + los[ip].ds); /* trapezoid_rule_pos :437-443 */
  los[0].ds *= 0.5;
  for (int ip = 0; ip < np; ip++)                                                         /* column_density :446-453 */
    for (int ig = 0; ig < c->ng; ig++) los[ip].u[ig] = 10. * los[ip].q[ig] * los[ip].p / (O_KB * los[ip].t) * los[ip].ds;
  return np;
}

/* hydrostatic_1d_h2o (:728-761) with find_reference_parcel (:714-726) and gravity (:213-217) */
static double o_gravity(double z, double lat) {
  const double deg2rad = M_PI / 180., x = sin(lat * deg2rad), y = sin(2 * lat * deg2rad);
  return 9.780318 * (1. + 0.0053024 * x * x - 5.8e-6 * y * y) - 3.086e-3 * z;
}
static void o_hydrostatic(const jrb_ctl_view *c, const jrb_atm_view *a, int ig_h2o) {
  const int npts = 20, ip0 = 0, ip1 = a->np;
  double dzmin = 1e99; int ipref = 0;
  for (int ip = ip0; ip < ip1; ip++) { const double dz = fabs(a->z[ip] - c->hydz); if (dz < dzmin) { dzmin = dz; ipref = ip; } }
  const double lat = a->lat[ipref], mmair = 28.96456e-3, mmh2o = 18.0153e-3;
  const double *qh = ig_h2o >= 0 ? (a->q_rows ? a->q_rows[ig_h2o] : a->q + (size_t)ig_h2o * a->q_stride) : NULL;
  double e = 0.;
  for (int ip = ipref + 1; ip < ip1; ip++) {
    double mean = 0.;
    for (int i = 0; i < npts; i++) {
      const double z = o_lip(0.0, a->z[ip - 1], npts - 1.0, a->z[ip], (double)i), grav = o_gravity(z, lat);
      if (qh) e = o_lip(0.0, qh[ip - 1], npts - 1.0, qh[ip], (double)i);
      const double temp = o_lip(0.0, a->t[ip - 1], npts - 1.0, a->t[ip], (double)i);
      mean += (e * mmh2o + (1 - e) * mmair) * grav / (O_RGAS * temp * npts);
    }
    a->p[ip] = a->p[ip - 1] * exp(-1000 * mean * (a->z[ip] - a->z[ip - 1]));
  }
  for (int ip = ipref - 1; ip >= ip0; ip--) {
    double mean = 0.;
    for (int i = 0; i < npts; i++) {
      const double z = o_lip(0.0, a->z[ip + 1], npts - 1.0, a->z[ip], (double)i), grav = o_gravity(z, lat);
      if (qh) e = o_lip(0.0, qh[ip + 1], npts - 1.0, qh[ip], (double)i);
      const double temp = o_lip(0.0, a->t[ip + 1], npts - 1.0, a->t[ip], (double)i);
      mean += (e * mmh2o + (1 - e) * mmair) * grav / (O_RGAS * temp * npts);
    }
    a->p[ip] = a->p[ip + 1] * exp(-1000 * mean * (a->z[ip] - a->z[ip + 1]));
  }
}

static o_pos *o_alloc_los(int ng, int nw) {
  o_pos *los = (o_pos *)calloc(O_NLOS, sizeof(o_pos));
  double *buf = (double *)calloc((size_t)O_NLOS * (2 * (ng + 1) + nw + 1), sizeof(double));
  for (int i = 0; i < O_NLOS; i++) {
    los[i].q = buf + (size_t)i * (2 * (ng + 1) + nw + 1);
    los[i].u = los[i].q + ng + 1;
    los[i].k = los[i].u + ng + 1;
  }
  return los;
}
static void o_free_los(o_pos *los) { free(los[0].q); free(los); }

static int o_fourbit(const jrb_ctl_view *c) { /* CPUdrivers.c:130-134 */
  return ((1 == c->ctm_co2) && (c->ig_co2 >= 0)) * 8 + ((1 == c->ctm_h2o) && (c->ig_h2o >= 0)) * 4 + (1 == c->ctm_n2) * 2 +
         (1 == c->ctm_o2) * 1;
}

/* formod_CPU (CPUdrivers.c:108-151) for one package */
int jro_formod(const jrb_ctl_view *c, const jrb_tbl_view *v, const jrb_atm_view *a, const jrb_obs_view *o) {
  if (c->ng > JRB_MAX_NG || c->nw > JRB_MAX_NW || c->formod != 2 || c->ip < 1 || c->ip > 3) return -1;
  const int nr = o->nr, nd = c->nd, fourbit = o_fourbit(c);
  char *mask = (char *)malloc((size_t)nr * nd + 1);
  for (int ir = 0; ir < nr; ir++)
    for (int id = 0; id < nd; id++) mask[(size_t)ir * nd + id] = !isfinite(o->rad[(size_t)ir * o->row_stride + id]); /* save_mask :193-200 */
  if (c->hydz >= 0) o_hydrostatic(c, a, c->ig_h2o); /* hydrostatic1d_CPU :97-103 (idempotent, so once) */
  int fatal = 0;
  int too_many = 0; /* the reference is fatal ("Too many LOS points!") when a ray needs NLOS points or more (:693-695) */
#pragma omp parallel
  {
    o_pos *los = o_alloc_los(c->ng, c->nw);
    double *tau_path = (double *)malloc(sizeof(double) * (size_t)nd * (c->ng + 1));
#pragma omp for schedule(dynamic, 1)
    for (int ir = 0; ir < nr; ir++) {
      double tsurf;
      int np = o_traceray(c, a, o, ir, los, &tsurf);
      if (np < 0) { /* fatal in the reference (2-D profile list, src/jurassic.c:727-728) */
#pragma omp atomic write
        fatal = np;
        np = 0;
      }
      if (np >= O_NLOS) {
#pragma omp atomic write
        too_many = 1;
      }
      double *rad = o->rad + (size_t)ir * o->row_stride, *tau = o->tau + (size_t)ir * o->row_stride;
      for (int id = 0; id < o->nd_reset; id++) { rad[id] = 0.0; tau[id] = 1.0; } /* apply_kernels_CPU :57-64 */
      for (int j = 0; j < nd * (c->ng + 1); j++) tau_path[j] = 1.0;
      for (int ip = 0; ip < np; ++ip)
        for (int id = 0; id < nd; id++) {
          const double beta_ds = o_continua(c, fourbit, &los[ip], id);
          double tau_gas = 1.0; /* apply_ega_core :270-280 */
          for (int ig = 0; ig < c->ng; ig++) {
            const double e = o_ega_eps(v, tau_path[id * (c->ng + 1) + ig], los[ip].t, los[ip].u[ig], los[ip].p, ig, id);
            tau_path[id * (c->ng + 1) + ig] *= e;
            tau_gas *= e;
          }
          const double planck = o_planck(v, los[ip].t, id);
          if (tau_gas > 1e-50) { /* new_obs_core :293-300 */
            const double eps = 1. - tau_gas * exp(-beta_ds);
            rad[id] += planck * eps * tau[id];
            tau[id] *= (1. - eps);
          }
        }
      if (tsurf > 0.) /* add_surface_core :227-234 */
        for (int id = 0; id < nd; id++) rad[id] += o_planck(v, tsurf, id) * tau[id];
      if (c->write_bbt) /* brightness_core :188-190 */
        for (int id = 0; id < nd; id++) rad[id] = O_C2 * c->nu[id] / log1p((O_C1 * c->nu[id] * c->nu[id] * c->nu[id]) / rad[id]);
      for (int id = 0; id < nd; id++) if (mask[(size_t)ir * nd + id]) rad[id] = NAN; /* apply_mask :203-210 */
    }
    free(tau_path);
    o_free_los(los);
  }
  free(mask);
  return too_many ? -2 : fatal;
}

/* LOS of one ray flattened like oracle/ref_hooks.c:jrref_traceray:
 * out[ip*stride + {0:z,1:lon,2:lat,3:p,4:t,5:ds, 6..:k[nw], then q[ng], then u[ng]}], stride = 6+nw+2*ng */
int jro_traceray(const jrb_ctl_view *c, const jrb_atm_view *a, const jrb_obs_view *o, int ir, double *out, double *tsurf) {
  o_pos *los = o_alloc_los(c->ng, c->nw);
  const int np = o_traceray(c, a, o, ir, los, tsurf);
  const int ng = c->ng, nw = c->nw, stride = 6 + nw + 2 * ng;
  for (int ip = 0; ip < np; ip++) {
    double *q = out + (size_t)ip * stride;
    q[0] = los[ip].z; q[1] = los[ip].lon; q[2] = los[ip].lat; q[3] = los[ip].p; q[4] = los[ip].t; q[5] = los[ip].ds;
    for (int iw = 0; iw < nw; iw++) q[6 + iw] = los[ip].k[iw];
    for (int ig = 0; ig < ng; ig++) { q[6 + nw + ig] = los[ip].q[ig]; q[6 + nw + ng + ig] = los[ip].u[ig]; }
  }
  o_free_los(los);
  return np;
}

/* formod_fov (src/jurassic.c:214-258) for one package, applied to the results already in obs (i.e. what a caller gets
 * from formod(); formod_fov();).  n, dz, w: the shape file as read_shape returns it.  -1: "Cannot apply FOV convolution!" */
int jro_formod_fov(const jrb_obs_view *o, int nd, int n, const double *dz, const double *w) {
  enum { NFOV = 5 }; /* src/jurassic.h:175 */
  const int nr = o->nr;
  double *rad0 = (double *)malloc(sizeof(double) * (size_t)(nr ? nr : 1) * nd * 2), *tau0 = rad0 + (size_t)nr * nd;
  for (int ir = 0; ir < nr; ir++) /* copy_obs (:224) */
    for (int id = 0; id < nd; id++) {
      rad0[(size_t)ir * nd + id] = o->rad[(size_t)ir * o->row_stride + id];
      tau0[(size_t)ir * nd + id] = o->tau[(size_t)ir * o->row_stride + id];
    }
  int rc = 0;
  for (int ir = 0; ir < nr && rc == 0; ir++) {
    double z[2 * NFOV + 1];
    int src[2 * NFOV + 1], nz = 0;
    for (int ir2 = (ir - NFOV > 0 ? ir - NFOV : 0); ir2 < (ir + 1 + NFOV < nr ? ir + 1 + NFOV : nr); ir2++) /* :227-235 */
      if (o->time[ir2] == o->time[ir]) { z[nz] = o->vpz[ir2]; src[nz] = ir2; nz++; }
    if (nz < 2) { rc = -1; break; } /* :236 */
    double *rr = o->rad + (size_t)ir * o->row_stride, *tt = o->tau + (size_t)ir * o->row_stride;
    double wsum = 0;
    for (int id = 0; id < nd; id++) { rr[id] = 0; tt[id] = 0; }
    for (int i = 0; i < n; i++) { /* :243-251 */
      const double zfov = o->vpz[ir] + dz[i];
      const int k = o_locate(z, nz, zfov);
      for (int id = 0; id < nd; id++) {
        const double r0 = rad0[(size_t)src[k] * nd + id], r1 = rad0[(size_t)src[k + 1] * nd + id];
        const double t0 = tau0[(size_t)src[k] * nd + id], t1 = tau0[(size_t)src[k + 1] * nd + id];
        rr[id] += w[i] * (r0 + (zfov - z[k]) * (r1 - r0) / (z[k + 1] - z[k])); /* LIN, src/jurassic.h:81 */
        tt[id] += w[i] * (t0 + (zfov - z[k]) * (t1 - t0) / (z[k + 1] - z[k]));
      }
      wsum += w[i];
    }
    for (int id = 0; id < nd; id++) { rr[id] /= wsum; tt[id] /= wsum; } /* :253-256 */
  }
  free(rad0);
  return rc;
}

int jro_max_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
