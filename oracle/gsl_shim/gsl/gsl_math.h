/* Minimal stand-in for <gsl/gsl_math.h> -- TEST INFRASTRUCTURE ONLY.
 * GSL 2.5 (lib/build.sh:15 of the reference) is neither installed nor buildable offline;
 * the reference hot path needs only the few names below.  Values/semantics follow GSL 2.5. */
#ifndef JRB_GSL_SHIM_MATH_H
#define JRB_GSL_SHIM_MATH_H
#include <math.h>
#ifndef GSL_NAN
#define GSL_NAN (NAN)
#endif
#define GSL_POSINF (INFINITY)
#define GSL_NEGINF (-INFINITY)
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))
#define GSL_MAX_DBL(a, b) GSL_MAX(a, b)
#define GSL_MIN_DBL(a, b) GSL_MIN(a, b)
#define GSL_MAX_INT(a, b) GSL_MAX(a, b)
#define GSL_MIN_INT(a, b) GSL_MIN(a, b)
static inline int gsl_finite(double x) { return isfinite(x) ? 1 : 0; }
static inline int gsl_isnan(double x) { return isnan(x) ? 1 : 0; }
static inline double gsl_pow_2(double x) { return x * x; }
static inline double gsl_pow_3(double x) { return x * x * x; }
/* libm versions; GSL's own differ by <= 1 ulp and are off the hot path (planck/brightness helpers) */
static inline double gsl_log1p(double x) { return log1p(x); }
static inline double gsl_expm1(double x) { return expm1(x); }
#endif
