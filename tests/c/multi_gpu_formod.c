/* multi_gpu_formod.c -- C caller of the drop-in library (no Python anywhere in the data path).
 *
 * What a JURASSIC host program does with more than one GPU behind formod(): build ctl_t / tbl_t / atm_t[] / obs_t[] in
 * host memory, call jr_b200_init_multi() once and jr_b200_formod_batch() per batch.  The test
 *   1. runs the batch on ONE device,
 *   2. runs it again on ALL visible devices (tables broadcast by NCCL inside the library, contiguous package slices, every
 *      device storing straight into the caller's obs_t rows) and requires bit-identical rad / tau / tangent points
 *      (SURVEY.md section 7, T8),
 *   3. repeats (2) with page-locked structs (jr_b200_pin_packages: direct I/O) -- same bits again,
 *   4. checks packages of every device slice against the CPU restatement oracle (checker only, tolerance 1e-6),
 *   5. hammers formod_GPU from several host threads (lanes),
 * and prints one JSON line with the timings.  Exit code 0 = all checks passed.
 *
 * Build (tests/c/Makefile): gcc -DND=32 -DNG=5 -I include -I jurassic-gpu_b200/csrc -I oracle ... -ljurassic_b200_dropin_nd32_ng5
 * usage: multi_gpu_formod [packages=16] [devices=0 (all)]
 */
#include "jr_structs.h"
#include <jurassic_b200.h>
#include <jurassic_b200_dropin.h>
#include <jr_oracle.h>

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static double planck(double t, double nu) { return 1.19104259e-8 * nu * nu * nu / expm1(1.43877506 * nu / t); }

/* analytic emissivity tables (the recipe of SURVEY.md 8d): monotone in u and eps, geometric u axis */
static void fill_tables(ctl_t const *ctl, tbl_t *tbl) {
  static double const kappa_g[5] = {3e-22, 2e-23, 5e-21, 4e-19, 6e-19};
  for (int ig = 0; ig < ctl->ng; ig++)
    for (int id = 0; id < ctl->nd; id++) {
      double const k0 = kappa_g[ig % 5] * (1 + 0.5 * sin(0.37 * id + ig));
      tbl->np[ig][id] = 36;
      for (int ip = 0; ip < 36; ip++) {
        double const p = 1e-3 * pow(10., 6.2 * ip / 35.);
        tbl->p[ig][ip][id] = p;
        tbl->nt[ig][ip][id] = 12;
        for (int it = 0; it < 12; it++) {
          double const T = 180. + 12. * it;
          tbl->t[ig][ip][it][id] = T;
          double const kap = k0 * (0.3 + 0.7 * pow(p / 1013.25, 0.6)) * (1 + 0.004 * (T - 250));
          int n = 0;
          float eps_old = -1, u_old = -1;
          for (int iu = 0; iu < 400 && n < 300; iu++) {
            double const u = 1e12 * pow(10., 0.05 * iu);
            double const e = 1 - 0.5 * exp(-kap * u) - 0.5 * exp(-0.05 * kap * u);
            if (e <= 1e-7) continue;
            float const uf = (float)u, ef = (float)e;
            if (n > 0 && !(uf > u_old && ef > eps_old)) continue;
            tbl->u[ig][ip][it][n][id] = uf; tbl->eps[ig][ip][it][n][id] = ef;
            u_old = uf; eps_old = ef; n++;
            if (e > 0.99999) break;
          }
          tbl->nu[ig][ip][it][id] = n;
        }
      }
    }
  for (int it = 0; it < TBLNS; it++) {
    tbl->st[it] = 100. + 0.25 * it;
    for (int id = 0; id < ctl->nd; id++) tbl->sr[it][id] = planck(tbl->st[it], ctl->nu[id]);
  }
}

static unsigned long long rng_state = 88172645463325252ull;
static double urand(void) { /* xorshift64 */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}

/* one package of the limb-sounder shape: 17 profiles x 64 tangent heights */
static void fill_package(ctl_t const *ctl, atm_t *atm, obs_t *obs) {
  int n = 0;
  for (int j = 0; j < 17; j++) {
    double const fp = 1 + 0.05 * (2 * urand() - 1), dT = 30 * (2 * urand() - 1);
    for (int iz = 0; iz <= 90; iz++, n++) {
      double const z = iz;
      double T = z < 11 ? 288.15 - 6.5 * z : (z < 20 ? 216.65 : (z < 47 ? 216.65 + 2.0 * (z - 20) : (z < 51 ? 270.65 : 270.65 - 2.5 * (z - 51))));
      if (T < 170) T = 170;
      atm->time[n] = j; atm->z[n] = z; atm->lon[n] = 0; atm->lat[n] = 0;
      atm->p[n] = 1013.25 * exp(-z / 7.0) * fp;
      atm->t[n] = T + dT;
      atm->q[0][n] = 3.9e-4;
      atm->q[1][n] = 1e-2 * exp(-z / 2.0) + 4e-6;
      atm->q[2][n] = 1e-7 + 6e-6 * exp(-(z - 32) * (z - 32) / 120.);
      atm->q[3][n] = 2.5e-10 * exp(-z / 12.0);
      atm->q[4][n] = 1e-10 * exp(-z / 10.0);
      atm->k[0][n] = 0;
    }
  }
  atm->np = n; atm->init = 0;
  int r = 0;
  for (int j = 0; j < 17; j++)
    for (int k = 0; k < 64; k++, r++) {
      obs->time[r] = j; obs->obsz[r] = 780; obs->obslon[r] = 0; obs->obslat[r] = 0;
      obs->vpz[r] = 3 + k; obs->vplon[r] = 0;
      obs->vplat[r] = acos((6367.421 + obs->vpz[r]) / (6367.421 + 780.)) * 180 / M_PI; /* src/limb.c:56 */
    }
  obs->nr = r;
  (void)ctl;
}

static int same_bits(obs_t *const a[], obs_t *const b[], int n, char const *what) {
  for (int i = 0; i < n; i++)
    if (memcmp(a[i]->rad, b[i]->rad, sizeof(a[i]->rad)) || memcmp(a[i]->tau, b[i]->tau, sizeof(a[i]->tau)) ||
        memcmp(a[i]->tpz, b[i]->tpz, sizeof(double) * (size_t)a[i]->nr) || memcmp(a[i]->tplat, b[i]->tplat, sizeof(double) * (size_t)a[i]->nr)) {
      printf("FAIL %s: package %d differs\n", what, i);
      return 0;
    }
  return 1;
}

static void clear_outputs(obs_t *const o[], int n) {
  for (int i = 0; i < n; i++) { memset(o[i]->rad, 0, sizeof(o[i]->rad)); memset(o[i]->tau, 0, sizeof(o[i]->tau)); memset(o[i]->tpz, 0, sizeof(o[i]->tpz)); }
}

int main(int argc, char **argv) {
  int const npk = argc > 1 ? atoi(argv[1]) : 16;
  int const want_dev = argc > 2 ? atoi(argv[2]) : 0;
  ctl_t *ctl = calloc(1, sizeof(ctl_t));
  tbl_t *tbl = calloc(1, sizeof(tbl_t));
  if (!ctl || !tbl) return 2;
  char const *gases[5] = {"CO2", "H2O", "O3", "F11", "CCl4"};
  ctl->ng = 5; ctl->nd = 32; ctl->nw = 1;
  for (int ig = 0; ig < 5; ig++) strcpy(ctl->emitter[ig], gases[ig]);
  for (int id = 0; id < 32; id++) { ctl->nu[id] = 785 + id; ctl->window[id] = 0; }
  ctl->hydz = -999; ctl->ctm_co2 = 1; ctl->ctm_h2o = 1; ctl->ctm_n2 = 0; ctl->ctm_o2 = 0; /* read_ctl's auto switch-off (src/jurassic.c:954-968) */
  ctl->ip = 1; ctl->refrac = 1; ctl->rayds = 10; ctl->raydz = 0.5; ctl->formod = 2; ctl->useGPU = 1;
  strcpy(ctl->fov, "-");
  fill_tables(ctl, tbl);

  atm_t **atm = malloc(sizeof(*atm) * (size_t)npk);
  obs_t **obs = malloc(sizeof(*obs) * (size_t)npk), **one = malloc(sizeof(*one) * (size_t)npk);
  for (int i = 0; i < npk; i++) {
    atm[i] = calloc(1, sizeof(atm_t)); obs[i] = calloc(1, sizeof(obs_t)); one[i] = calloc(1, sizeof(obs_t));
    if (!atm[i] || !obs[i] || !one[i]) return 2;
    fill_package(ctl, atm[i], obs[i]);
    memcpy(one[i], obs[i], sizeof(obs_t));
  }
  obs[npk / 2]->rad[17][3] = NAN; one[npk / 2]->rad[17][3] = NAN; /* NaN mask travels through every path */
  int ok = 1;

  /* 1. one device */
  int nd1 = jr_b200_init_multi(ctl, tbl, 1);
  jr_b200_formod_batch(ctl, atm, one, npk); /* warm-up (allocations) */
  double t0 = now_ms();
  jr_b200_formod_batch(ctl, atm, one, npk);
  double const ms_one = now_ms() - t0;
  jr_b200_finalize();

  /* 2. all devices, staged I/O */
  int const ndev = jr_b200_init_multi(ctl, tbl, want_dev);
  jrb_group_stats gs;
  jrb_group_get_stats((jrb_group *)jr_b200_core_group(), &gs);
  jr_b200_formod_batch(ctl, atm, obs, npk);
  t0 = now_ms();
  jr_b200_formod_batch(ctl, atm, obs, npk);
  double const ms_all = now_ms() - t0;
  ok &= same_bits(obs, one, npk, "all devices (staged) vs one device");
  if (!isnan(obs[npk / 2]->rad[17][3])) { printf("FAIL NaN mask lost\n"); ok = 0; }
  jrb_group_get_stats((jrb_group *)jr_b200_core_group(), &gs);
  if (ndev > 1 && (gs.nccl_nranks != ndev || gs.n_slices != ndev)) { printf("FAIL expected %d NCCL ranks / slices, saw %d / %d\n", ndev, gs.nccl_nranks, gs.n_slices); ok = 0; }

  /* 3. all devices, page-locked structs: direct I/O */
  clear_outputs(obs, npk);
  obs[npk / 2]->rad[17][3] = NAN;
  if (jr_b200_pin_packages(atm, obs, npk) != 0) { printf("FAIL jr_b200_pin_packages\n"); ok = 0; }
  jr_b200_formod_batch(ctl, atm, obs, npk);
  jrb_stats cs;
  jrb_get_stats((jrb_context *)jr_b200_core_context(), &cs);
  if (!cs.io_direct) { printf("FAIL pinned packages did not select direct I/O\n"); ok = 0; }
  t0 = now_ms();
  jr_b200_formod_batch(ctl, atm, obs, npk);
  double const ms_pin = now_ms() - t0;
  ok &= same_bits(obs, one, npk, "all devices (direct) vs one device");

  /* 4. oracle on the first package of every device slice and the last package (checker only) */
  double worst_rad = 0, worst_tau = 0;
  {
    jrb_ctl_view cv;
    memset(&cv, 0, sizeof(cv));
    cv.ng = 5; cv.nd = 32; cv.nw = 1; cv.nu = ctl->nu; cv.window = ctl->window;
    cv.ctm_co2 = 1; cv.ctm_h2o = 1; cv.ig_co2 = 0; cv.ig_h2o = 1; cv.refrac = 1; cv.rayds = 10; cv.raydz = 0.5; cv.hydz = -999; cv.formod = 2; cv.ip = 1;
    jrb_tbl_view tv = {NG, TBLNP, TBLNT, TBLNU, ND, TBLNS, &tbl->np[0][0], &tbl->nt[0][0][0], &tbl->nu[0][0][0][0], &tbl->p[0][0][0],
                       &tbl->t[0][0][0][0], &tbl->u[0][0][0][0][0], &tbl->eps[0][0][0][0][0], &tbl->sr[0][0], &tbl->st[0]};
    obs_t *chk = calloc(1, sizeof(obs_t));
    for (int s = 0; s <= ndev; s++) {
      int const i = s == ndev ? npk - 1 : (int)((long long)s * npk / ndev);
      memcpy(chk, obs[i], sizeof(obs_t));
      memset(chk->rad, 0, sizeof(chk->rad));
      jrb_atm_view av = {atm[i]->np, atm[i]->time, atm[i]->z, atm[i]->lon, atm[i]->lat, atm[i]->p, atm[i]->t, &atm[i]->q[0][0], NP, &atm[i]->k[0][0], NP, NULL, NULL};
      jrb_obs_view ov = {chk->nr, chk->time, chk->obsz, chk->obslon, chk->obslat, chk->vpz, chk->vplon, chk->vplat, chk->tpz, chk->tplon, chk->tplat,
                         &chk->rad[0][0], &chk->tau[0][0], ND, ND};
      if (jro_formod(&cv, &tv, &av, &ov) != 0) { printf("FAIL oracle\n"); ok = 0; break; }
      double rmax = 0;
      for (int ir = 0; ir < chk->nr; ir++) for (int id = 0; id < 32; id++) rmax = fmax(rmax, fabs(chk->rad[ir][id]));
      for (int ir = 0; ir < chk->nr; ir++)
        for (int id = 0; id < 32; id++) {
          if (isnan(obs[i]->rad[ir][id])) continue;
          worst_rad = fmax(worst_rad, fabs(obs[i]->rad[ir][id] - chk->rad[ir][id]) / (fabs(chk->rad[ir][id]) + 1e-12 * rmax));
          worst_tau = fmax(worst_tau, fabs(obs[i]->tau[ir][id] - chk->tau[ir][id]) / (fabs(chk->tau[ir][id]) + 1e-12));
        }
    }
    free(chk);
    if (!(worst_rad <= 1e-6 && worst_tau <= 1e-6)) { printf("FAIL oracle parity: rad %.3e tau %.3e\n", worst_rad, worst_tau); ok = 0; }
  }

  /* 5. concurrent single-package callers: formod_GPU from 4 host threads, like OpenMP threads of a retrieval */
#pragma omp parallel for num_threads(4) schedule(dynamic, 1)
  for (int i = 0; i < npk; i++) formod_GPU(ctl, atm[i], obs[i]); /* warm-up: every lane allocates its buffers once */
  clear_outputs(obs, npk);
  obs[npk / 2]->rad[17][3] = NAN;
  t0 = now_ms();
#pragma omp parallel for num_threads(4) schedule(dynamic, 1)
  for (int i = 0; i < npk; i++) formod_GPU(ctl, atm[i], obs[i]);
  double const ms_lanes = now_ms() - t0;
  ok &= same_bits(obs, one, npk, "concurrent formod_GPU vs batch");
  t0 = now_ms();
  for (int i = 0; i < npk; i++) formod_GPU(ctl, atm[i], obs[i]);
  double const ms_serial = now_ms() - t0;

  jr_b200_finalize();
  jr_b200_unpin_all();
  long long const rc = (long long)npk * 1088 * 32;
  printf("{\"test\": \"multi_gpu_formod\", \"ok\": %s, \"packages\": %d, \"devices_first\": %d, \"devices\": %d, \"nccl_nranks\": %d, "
         "\"ms_one_device\": %.2f, \"ms_all_devices_staged\": %.2f, \"ms_all_devices_direct\": %.2f, \"ray_channels_per_s_direct\": %.4g, "
         "\"ms_formod_GPU_4_threads\": %.2f, \"ms_formod_GPU_serial\": %.2f, \"max_rel_err_rad\": %.3e, \"max_rel_err_tau\": %.3e}\n",
         ok ? "true" : "false", npk, nd1, ndev, gs.nccl_nranks, ms_one, ms_all, ms_pin, rc / (ms_pin * 1e-3), ms_lanes, ms_serial, worst_rad, worst_tau);
  return ok ? 0 : 1;
}
