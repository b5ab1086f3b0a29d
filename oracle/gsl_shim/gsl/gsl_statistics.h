/* Stand-in for <gsl/gsl_statistics.h>. TEST INFRASTRUCTURE ONLY. */
#ifndef JRB_GSL_SHIM_STATISTICS_H
#define JRB_GSL_SHIM_STATISTICS_H
#include <stddef.h>
static inline size_t gsl_stats_min_index(const double data[], size_t stride, size_t n) {
  size_t k = 0; double m = data[0];
  for (size_t i = 1; i < n; i++) if (data[i * stride] < m) { m = data[i * stride]; k = i; }
  return k;
}
#endif
