/* Stand-in for <gsl/gsl_const_num.h>: GSL 2.5 value. TEST INFRASTRUCTURE ONLY. */
#ifndef JRB_GSL_SHIM_CONST_NUM_H
#define JRB_GSL_SHIM_CONST_NUM_H
#define GSL_CONST_NUM_AVOGADRO (6.02214199e23) /* 1 / mol */
#endif
