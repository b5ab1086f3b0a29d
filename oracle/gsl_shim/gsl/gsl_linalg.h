/* Empty stand-in for <gsl/gsl_linalg.h> (nothing from it is used). TEST INFRASTRUCTURE ONLY. */
#ifndef JRB_GSL_SHIM_LINALG_H
#define JRB_GSL_SHIM_LINALG_H
#endif
