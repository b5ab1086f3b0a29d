/* Stand-in for <gsl/gsl_blas.h>: just enough gsl_vector/gsl_matrix for the reference's
 * retrieval helpers (off the hot path) to compile.  TEST INFRASTRUCTURE ONLY. */
#ifndef JRB_GSL_SHIM_BLAS_H
#define JRB_GSL_SHIM_BLAS_H
#include <stdlib.h>
#include <string.h>
typedef struct { size_t size, stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { size_t size1, size2, tda; double *data; void *block; int owner; } gsl_matrix;
static inline gsl_vector *gsl_vector_alloc(size_t n) {
  gsl_vector *v = (gsl_vector *)malloc(sizeof(gsl_vector));
  v->size = n; v->stride = 1; v->data = (double *)calloc(n ? n : 1, sizeof(double)); v->block = 0; v->owner = 1;
  return v;
}
static inline void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
static inline double gsl_vector_get(gsl_vector const *v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i * v->stride] = x; }
static inline int gsl_vector_memcpy(gsl_vector *d, gsl_vector const *s) {
  for (size_t i = 0; i < s->size; i++) d->data[i * d->stride] = s->data[i * s->stride];
  return 0;
}
static inline gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) {
  gsl_matrix *m = (gsl_matrix *)malloc(sizeof(gsl_matrix));
  m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = (double *)calloc(n1 * n2 ? n1 * n2 : 1, sizeof(double));
  m->block = 0; m->owner = 1;
  return m;
}
static inline void gsl_matrix_free(gsl_matrix *m) { if (m) { free(m->data); free(m); } }
static inline double gsl_matrix_get(gsl_matrix const *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
static inline void gsl_matrix_set_zero(gsl_matrix *m) {
  for (size_t i = 0; i < m->size1; i++) memset(m->data + i * m->tda, 0, m->size2 * sizeof(double));
}
#endif
