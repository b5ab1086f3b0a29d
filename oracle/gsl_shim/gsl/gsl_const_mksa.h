/* Stand-in for <gsl/gsl_const_mksa.h>: GSL 2.5 values (CODATA 2006). TEST INFRASTRUCTURE ONLY. */
#ifndef JRB_GSL_SHIM_CONST_MKSA_H
#define JRB_GSL_SHIM_CONST_MKSA_H
#define GSL_CONST_MKSA_BOLTZMANN (1.3806504e-23) /* kg m^2 / K s^2 */
#define GSL_CONST_MKSA_MOLAR_GAS (8.314472e0)    /* kg m^2 / K mol s^2 */
#endif
