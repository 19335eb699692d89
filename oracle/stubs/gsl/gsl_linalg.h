/* Stand-in for <gsl/gsl_linalg.h>: only EvaluatedMachine::sumInTrans uses it (not on the DP path). */
#ifndef MB_ORACLE_STUB_GSL_LINALG_H
#define MB_ORACLE_STUB_GSL_LINALG_H
#include "gsl_matrix.h"
#include "gsl_randist.h"
struct gsl_permutation { size_t size; };
static inline gsl_permutation* gsl_permutation_alloc (size_t n) { gsl_permutation* p = new gsl_permutation; p->size = n; return p; }
static inline void gsl_permutation_free (gsl_permutation* p) { delete p; }
static inline int gsl_linalg_LU_decomp (gsl_matrix*, gsl_permutation*, int*) { mb_stub_gsl_unavailable ("gsl_linalg_LU_decomp"); return 0; }
static inline int gsl_linalg_LU_invert (const gsl_matrix*, const gsl_permutation*, gsl_matrix*) { mb_stub_gsl_unavailable ("gsl_linalg_LU_invert"); return 0; }
#endif
