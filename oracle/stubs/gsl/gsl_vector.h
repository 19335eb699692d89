/* Header-only stand-in for <gsl/gsl_vector.h>: test infrastructure only.
 * GSL is a system dependency of the reference that is absent from this image; the DP hot path
 * (forward/backward/viterbi/counts) never calls it.  Only the container type is real. */
#ifndef MB_ORACLE_STUB_GSL_VECTOR_H
#define MB_ORACLE_STUB_GSL_VECTOR_H
#include <cstdlib>
#include <cstddef>
struct gsl_vector { size_t size; double* data; };
static inline gsl_vector* gsl_vector_alloc (size_t n) {
  gsl_vector* v = (gsl_vector*) std::malloc (sizeof (gsl_vector));
  v->size = n; v->data = (double*) std::calloc (n ? n : 1, sizeof (double));
  return v;
}
static inline void gsl_vector_free (gsl_vector* v) { if (v) { std::free (v->data); std::free (v); } }
static inline double gsl_vector_get (const gsl_vector* v, size_t i) { return v->data[i]; }
static inline void gsl_vector_set (gsl_vector* v, size_t i, double x) { v->data[i] = x; }
#endif
