/* Stand-in for <gsl/gsl_multimin.h>: the BFGS M-step (MachineObjective::optimize) is outside the
 * accelerated path; the declarations let counts.cpp compile, and any call aborts. */
#ifndef MB_ORACLE_STUB_GSL_MULTIMIN_H
#define MB_ORACLE_STUB_GSL_MULTIMIN_H
#include "gsl_vector.h"
#include "gsl_randist.h"
#define GSL_SUCCESS 0
#define GSL_CONTINUE (-2)
struct gsl_multimin_function_fdf {
  double (*f) (const gsl_vector*, void*);
  void (*df) (const gsl_vector*, void*, gsl_vector*);
  void (*fdf) (const gsl_vector*, void*, double*, gsl_vector*);
  size_t n;
  void* params;
};
struct gsl_multimin_fdfminimizer_type { const char* name; };
struct gsl_multimin_fdfminimizer { gsl_vector* x; gsl_vector* gradient; double f; };
static const gsl_multimin_fdfminimizer_type mb_stub_bfgs2 = { "stub" };
static const gsl_multimin_fdfminimizer_type* const gsl_multimin_fdfminimizer_vector_bfgs2 = &mb_stub_bfgs2;
static inline gsl_multimin_fdfminimizer* gsl_multimin_fdfminimizer_alloc (const gsl_multimin_fdfminimizer_type*, size_t) { mb_stub_gsl_unavailable ("gsl_multimin"); return 0; }
static inline int gsl_multimin_fdfminimizer_set (gsl_multimin_fdfminimizer*, gsl_multimin_function_fdf*, const gsl_vector*, double, double) { return 0; }
static inline int gsl_multimin_fdfminimizer_iterate (gsl_multimin_fdfminimizer*) { return 0; }
static inline int gsl_multimin_test_gradient (const gsl_vector*, double) { return 0; }
static inline void gsl_multimin_fdfminimizer_free (gsl_multimin_fdfminimizer*) { }
static inline const char* gsl_strerror (int) { return "stub"; }
#endif
