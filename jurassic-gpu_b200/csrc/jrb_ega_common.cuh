// jrb_ega_common.cuh -- per-(segment,channel) physics shared by the generic and the specialised EGA kernels:
// continuum optical depth, Planck source lookup, radiance/transmittance update, surface and brightness epilogues.
#pragma once
#include "jrb_internal.h"

namespace jrb {

// 1/(N_A * 1000 * P0): scale of the CO2 continuum (src/jr_common.h:330)
constexpr double kInvCo2Scale = 1.0 / (kAvogadro * 1000 * kP0);

// Continuum + extinction optical depth of one segment for one channel (continua_core_bbbb,
// src/jr_continua_core.mv4g.h:1-14; formulas src/jr_common.h:315-390).  `mask` = CO2*8+H2O*4+N2*2+O2; when it is a
// compile-time constant the untaken branches vanish.  Channel-only factors were hoisted to the host (chan[]).
__device__ __forceinline__ double continuum_beta_ds(const int mask, const double *__restrict__ chan, const int nd,
                                                    const int id, const double p, const double t, const double ds,
                                                    const double kext, const double u_co2, const double u_h2o,
                                                    const double q_h2o) {
  double beta_ds = kext * ds;
  if (mask & 8) {
    const double cw296 = chan[CH_CO2_296 * nd + id], cw260 = chan[CH_CO2_260 * nd + id],
                 cw230 = chan[CH_CO2_230 * nd + id];
    const double dt230 = t - 230, dt260 = t - 260, dt296 = t - 296;
    const double ctw = dt260 * 5.050505e-4 * dt296 * cw230 - dt230 * 9.259259e-4 * dt296 * cw260 +
                       dt230 * 4.208754e-4 * dt260 * cw296;
    beta_ds += u_co2 * p * ctw * kInvCo2Scale;
  }
  if (mask & 4) {
    const double nu = chan[CH_NU * nd + id];
    const double s296 = chan[CH_H2O_S296 * nd + id], lnr = chan[CH_H2O_LNRATIO * nd + id],
                 frn = chan[CH_H2O_FRN * nd + id];
    const double ctwslf = s296 * exp(lnr * ((296. - t) / (296. - 260.)));
    const double a1 = nu * u_h2o * tanh(.7193876 / t * nu);
    const double a2 = 296. / t;
    const double a3 = p / kP0 * (q_h2o * ctwslf + (1 - q_h2o) * frn) * 1e-20;
    beta_ds += a1 * a2 * a3;
  }
  if (mask & 3) {
    const double pr = p / kP0, tr = 273.0 / t;
    const double common = 0.1 * pr * pr * tr * tr;
    const double dinv = 1.0 / 296.0 - 1.0 / t;
    if (mask & 2) {
      const double b = chan[CH_N2_B * nd + id], beta = chan[CH_N2_BETA * nd + id];
      const double q_n2 = 0.79;
      beta_ds += common * exp(beta * dinv) * q_n2 * b * (q_n2 + (1 - q_n2) * (1.294 - 0.4545 * t / 296.0)) * ds;
    }
    if (mask & 1) {
      const double b = chan[CH_O2_B * nd + id], beta = chan[CH_O2_BETA * nd + id];
      beta_ds += common * exp(beta * dinv) * 0.21 * b * ds;
    }
  }
  return beta_ds;
}

// Band-averaged Planck radiance at temperature t (src_planck_core, src/jr_common.h:220-224): the source axis is
// st[it] = 100 + 0.25 it, so the index is (int)(4t) - 400 (locate_st, :82-84) and the divisor of the linear
// interpolation is exactly 0.25.  The reference indexes without a bounds check (valid for 100 <= t < 400 K, outside it
// reads whatever lies next to the array); here the index is clipped to the table, i.e. temperatures outside the source
// table extrapolate its first / last interval -- identical inside the range, defined (and no stray access) outside.
__device__ __forceinline__ double planck_source(const double *__restrict__ sr, const int nd, const int id,
                                                const double t) {
  const int it = min(max((int)(4 * t) - 400, 0), kTBLNS - 2);
  const double st = 100.0 + 0.25 * it;
  const double y0 = sr[(size_t)it * nd + id], y1 = sr[(size_t)(it + 1) * nd + id];
  return y0 + (t - st) * (y1 - y0) * 4.0;
}

// new_obs_core (src/jr_common.h:293-300)
__device__ __forceinline__ void accumulate(double &rad, double &tau, const double beta_ds, const double src,
                                           const double tau_gas) {
  if (tau_gas > 1e-50) {
    const double eps = 1. - tau_gas * exp(-beta_ds);
    rad += src * eps * tau;
    tau *= (1. - eps);
  }
}

// add_surface_core (:227-234) and brightness_core (:188-190)
__device__ __forceinline__ void epilogue(double &rad, const double tau, const double tsurf,
                                         const double *__restrict__ sr, const int nd, const int id,
                                         const int write_bbt, const double nu) {
  if (tsurf > 0.) rad += planck_source(sr, nd, id, tsurf) * tau;
  if (write_bbt) rad = kC2 * nu / log1p((kC1 * nu * nu * nu) / rad);
}

} // namespace jrb
