/* Stand-in for <gsl/gsl_randist.h>: the pdf helpers are not on the DP path; calling one aborts. */
#ifndef MB_ORACLE_STUB_GSL_RANDIST_H
#define MB_ORACLE_STUB_GSL_RANDIST_H
#include <cstdio>
#include <cstdlib>
#include <cstddef>
static inline double mb_stub_gsl_unavailable (const char* what) {
  std::fprintf (stderr, "oracle stub: %s needs GSL, which is not in this image\n", what);
  std::abort();
  return 0;
}
static inline double gsl_ran_beta_pdf (double, double, double) { return mb_stub_gsl_unavailable ("gsl_ran_beta_pdf"); }
static inline double gsl_ran_gamma_pdf (double, double, double) { return mb_stub_gsl_unavailable ("gsl_ran_gamma_pdf"); }
static inline double gsl_ran_dirichlet_pdf (size_t, const double*, const double*) { return mb_stub_gsl_unavailable ("gsl_ran_dirichlet_pdf"); }
static inline double gsl_ran_gaussian_pdf (double, double) { return mb_stub_gsl_unavailable ("gsl_ran_gaussian_pdf"); }
#endif
