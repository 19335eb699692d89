/* Header-only stand-in for <gsl/gsl_matrix.h> (see gsl_vector.h). */
#ifndef MB_ORACLE_STUB_GSL_MATRIX_H
#define MB_ORACLE_STUB_GSL_MATRIX_H
#include <cstdlib>
#include <cstddef>
struct gsl_matrix { size_t size1, size2; double* data; };
static inline gsl_matrix* gsl_matrix_alloc (size_t r, size_t c) {
  gsl_matrix* m = (gsl_matrix*) std::malloc (sizeof (gsl_matrix));
  m->size1 = r; m->size2 = c; m->data = (double*) std::calloc (r * c ? r * c : 1, sizeof (double));
  return m;
}
static inline void gsl_matrix_free (gsl_matrix* m) { if (m) { std::free (m->data); std::free (m); } }
static inline double gsl_matrix_get (const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
static inline void gsl_matrix_set (gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }
#endif
