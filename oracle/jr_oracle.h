/* jr_oracle.h -- TEST INFRASTRUCTURE ONLY: interface of the CPU restatement (see jr_oracle.c). */
#ifndef JR_ORACLE_H
#define JR_ORACLE_H
#include <jurassic_b200.h> /* only for the view structs */
#ifdef __cplusplus
extern "C" {
#endif
/* complete forward model for one package; same semantics as formod_CPU (src/CPUdrivers.c:108-151). 0 on success,
 * -2 where the reference is fatal with "Too many LOS points!" (a ray of NLOS = 400 points or more, src/jr_common.h:693-695),
 * -3 / -4 for the fatal conditions of the 2-D profile list (ctl->ip == 2, src/jurassic.c:727-728) */
int jro_formod(const jrb_ctl_view *ctl, const jrb_tbl_view *tbl, const jrb_atm_view *atm, const jrb_obs_view *obs);
/* ray tracer only; flattened LOS, see jr_oracle.c */
int jro_traceray(const jrb_ctl_view *ctl, const jrb_atm_view *atm, const jrb_obs_view *obs, int ir, double *out, double *tsurf);
/* formod_fov (src/jurassic.c:214-258) applied in place to the rad/tau already in obs; shape = (dz[n], w[n]) */
int jro_formod_fov(const jrb_obs_view *obs, int nd, int n, const double *dz, const double *w);
/* intpol_atm_geo (src/jurassic.c:685-804; ctl->ip = 1, 2, 3) over the whole atmosphere; out = {p, t, q[ng], k[nw]} */
int jro_intpol_atm_geo(const jrb_ctl_view *ctl, const jrb_atm_view *atm, double z0, double lon0, double lat0, double *out);
int jro_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
