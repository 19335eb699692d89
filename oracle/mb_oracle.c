/* oracle/mb_oracle.c -- TEST INFRASTRUCTURE (not product code).  See mb_oracle.h.
 *
 * Restates, in plain C over a flat machine, what these reference functions compute (paths relative
 * to the reference root):
 *   src/logsumexp.h:48-90, src/logsumexp.cpp:10-18,49-51   log_sum_exp + lookup table
 *   src/forward.defs.h:22-55                               MappedForwardMatrix::fill / logLike
 *   src/backward.cpp:18-50                                 BackwardMatrix::fill / logLike
 *   src/backward.cpp:62-87, src/backward.h:37-42           getCounts / accumulateCounts
 *   src/viterbi.cpp:18-47                                  ViterbiMatrix::fill / logLike
 *   src/dpmatrix.defs.h:82-110,171-174                     traceBack / selectMaxTrans
 *   src/dpmatrix.h:106-115                                 iterate (candidate order)
 * The order of floating-point operations follows the reference exactly, so Forward / Backward /
 * Viterbi values are expected to be bit-identical to oracle/_ref/refdrv in MBO_LSE_TABLE mode.
 */
#include "mb_oracle.h"
#include "synth.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- log_sum_exp (logsumexp.h:20-23: max 10, precision 1e-4, 100001 entries) ---- */
#define LSE_MAX 10
#define LSE_PREC .0001
#define LSE_ENTRIES (((int) (LSE_MAX / LSE_PREC)) + 1)

static double* lse_table = NULL;

static double lse_unary_slow (double x) { return log (1. + exp (-x)); }   /* logsumexp.cpp:49-51 */

static void lse_init (void) {
  if (lse_table) return;
  double* t = (double*) malloc (sizeof (double) * LSE_ENTRIES);
  for (int n = 0; n < LSE_ENTRIES; ++n)   /* logsumexp.cpp:10-18 */
    t[n] = lse_unary_slow (n * LSE_PREC);
  lse_table = t;
}

static inline double lse_unary (double x, int mode) {   /* logsumexp.h:48-70 */
  if (mode == MBO_LSE_EXACT) return lse_unary_slow (x);
  if (x >= LSE_MAX || isnan (x) || isinf (x)) return 0;
  if (x < 0) return -x;
  const int n = (int) (x / LSE_PREC);
  const double f0 = lse_table[n];
  const double dx = x - (n * LSE_PREC);
  const double f1 = lse_table[n + 1];
  const double df = f1 - f0;
  return f0 + df * (dx / LSE_PREC);
}

static inline double lse2 (double a, double b, int mode) {   /* logsumexp.h:72-90 */
  double max, diff;
  if (mode == MBO_LSE_EXACT) {   /* logsumexp.cpp:24-39 */
    double min;
    if (a < b) { min = a; max = b; } else { min = b; max = a; }
    if (min == -INFINITY) return max;
    return max + lse_unary_slow (max - min);
  }
  if (a == b) { max = a; diff = 0; }
  else if (a < b) { max = b; diff = b - a; }
  else { max = a; diff = a - b; }
  return max + lse_unary (diff, mode);
}

double mbo_log_sum_exp (double a, double b, int lse_mode) { lse_init(); return lse2 (a, b, lse_mode); }

/* ---- transition index: the role of EvaluatedMachineState::incoming / outgoing (eval.h:68-76) ----
 * For key (state, inTok, outTok) a contiguous range of entries in the reference's multimap order:
 * incoming: source state ascending, ties in insertion order = transIndex ascending (eval.cpp:55-64)
 * outgoing: destination ascending, ties by transIndex ascending.                                  */
typedef struct {
  int64_t* off;     /* [nKeys+1] */
  int32_t* other;   /* source (incoming) or destination (outgoing) state */
  int32_t* id;      /* global transition id */
  double* lw;
  int nIn1, nOut1;
} trans_index;

static inline int64_t key_of (const trans_index* ix, int state, int in, int out) {
  return ((int64_t) state * ix->nIn1 + in) * ix->nOut1 + out;
}

static void index_build (trans_index* ix, const mbo_machine* m, int incoming) {
  const int64_t T = m->nTrans;
  ix->nIn1 = m->nInTok + 1;
  ix->nOut1 = m->nOutTok + 1;
  const int64_t nKeys = (int64_t) m->nStates * ix->nIn1 * ix->nOut1;
  ix->off = (int64_t*) calloc ((size_t) nKeys + 1, sizeof (int64_t));
  ix->other = (int32_t*) malloc (sizeof (int32_t) * (size_t) (T ? T : 1));
  ix->id = (int32_t*) malloc (sizeof (int32_t) * (size_t) (T ? T : 1));
  ix->lw = (double*) malloc (sizeof (double) * (size_t) (T ? T : 1));
  /* first order transitions by (other state, id): ids already ascend with src, so for incoming the
   * identity order is right; for outgoing do a stable counting sort by destination. */
  int64_t* order = (int64_t*) malloc (sizeof (int64_t) * (size_t) (T ? T : 1));
  if (incoming) {
    for (int64_t t = 0; t < T; ++t) order[t] = t;
  } else {
    int64_t* cnt = (int64_t*) calloc ((size_t) m->nStates + 1, sizeof (int64_t));
    for (int64_t t = 0; t < T; ++t) cnt[m->dst[t] + 1]++;
    for (int s = 0; s < m->nStates; ++s) cnt[s + 1] += cnt[s];
    for (int64_t t = 0; t < T; ++t) order[cnt[m->dst[t]]++] = t;
    free (cnt);
  }
  /* then a stable counting sort by key */
  for (int64_t n = 0; n < T; ++n) {
    const int64_t t = order[n];
    ix->off[key_of (ix, incoming ? m->dst[t] : m->src[t], m->in[t], m->out[t]) + 1]++;
  }
  for (int64_t k = 0; k < nKeys; ++k) ix->off[k + 1] += ix->off[k];
  int64_t* pos = (int64_t*) malloc (sizeof (int64_t) * (size_t) (nKeys ? nKeys : 1));
  memcpy (pos, ix->off, sizeof (int64_t) * (size_t) nKeys);
  for (int64_t n = 0; n < T; ++n) {
    const int64_t t = order[n];
    const int64_t p = pos[key_of (ix, incoming ? m->dst[t] : m->src[t], m->in[t], m->out[t])]++;
    ix->other[p] = incoming ? m->src[t] : m->dst[t];
    ix->id[p] = (int32_t) t;
    ix->lw[p] = m->logWeight[t];
  }
  free (pos);
  free (order);
}

static void index_free (trans_index* ix) { free (ix->off); free (ix->other); free (ix->id); free (ix->lw); }

/* ---- envelope (seqpair.h:75-113); NULL = full ---- */
static __thread const int64_t* env_start = NULL;
static __thread const int64_t* env_end = NULL;
void mbo_set_envelope (const int64_t* envStart, const int64_t* envEnd) { env_start = envStart; env_end = envEnd; }
static inline int env_contains (int64_t i, int64_t o) { return !env_start || (i >= env_start[o] && i < env_end[o]); }

/* ---- storage: full [o][i][s] or two rolling output rows (dpmatrix.h:35-58) ---- */
typedef struct { double* c; int64_t Li; int S; int rolling; const double* ninf; } cells;
static inline double* cellp (const cells* m, int64_t i, int64_t o) {
  return m->c + ((m->rolling ? (o & 1) : o) * (m->Li + 1) + i) * m->S;
}
/* const read access: -inf outside the envelope (dpmatrix.h:142-144) */
static inline const double* cellr (const cells* m, int64_t i, int64_t o) {
  return env_contains (i, o) ? cellp (m, i, o) : m->ninf;
}
/* DPMatrix::alloc fills the storage with -inf (dpmatrix.defs.h:36): a read of a cell that has not been written
   yet -- the silent self-loop isAdvancingMachine allows on state 0 (machine.cpp:759) -- sees -inf in a full
   matrix, and whatever row o-2 left behind in a rolling one */
static void fill_ninf (double* c, size_t n) { for (size_t q = 0; q < n; ++q) c[q] = -INFINITY; }
static double* make_ninf (int S) {
  double* v = (double*) malloc (sizeof (double) * (size_t) (S ? S : 1));
  for (int s = 0; s < S; ++s) v[s] = -INFINITY;
  return v;
}

/* ================= Forward (forward.defs.h:22-55) ================= */
double mbo_forward (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                    int mode, double* matrix) {
  lse_init();
  trans_index ix; index_build (&ix, m, 1);
  const int S = m->nStates;
  double* ninf = make_ninf (S);
  cells M = { matrix, Li, S, matrix == NULL, ninf };
  if (!matrix) M.c = (double*) malloc (sizeof (double) * 2 * (size_t) (Li + 1) * S);
  fill_ninf (M.c, (size_t) (matrix ? Lo + 1 : 2) * (size_t) (Li + 1) * S);
  for (int64_t o = 0; o <= Lo; ++o) {
    const int outTok = o ? y[o - 1] : 0;
    for (int64_t i = 0; i <= Li; ++i) {
      const int inTok = i ? x[i - 1] : 0;
      double* cur = cellp (&M, i, o);
      if (!env_contains (i, o)) { for (int d = 0; d < S; ++d) cur[d] = -INFINITY; continue; }   /* storage stays -inf (dpmatrix.defs.h:36) */
      for (int d = 0; d < S; ++d) {
        double ll = (i || o || d != 0) ? -INFINITY : 0;
#define ACC(IN, OUT, SRC)                                                           \
        do { const int64_t k_ = key_of (&ix, d, (IN), (OUT));                       \
             const double* s_ = (SRC);                                              \
             for (int64_t p = ix.off[k_]; p < ix.off[k_ + 1]; ++p)                  \
               ll = lse2 (ll, s_[ix.other[p]] + ix.lw[p], mode); } while (0)
        if (i && o) ACC (inTok, outTok, cellr (&M, i - 1, o - 1));
        if (i) ACC (inTok, 0, cellr (&M, i - 1, o));
        if (o) ACC (0, outTok, cellr (&M, i, o - 1));
        ACC (0, 0, cur);
#undef ACC
        cur[d] = ll;
      }
    }
  }
  const double result = cellr (&M, Li, Lo)[S - 1];
  if (!matrix) free (M.c);
  free (ninf);
  index_free (&ix);
  return result;
}

/* ================= Backward (backward.cpp:18-50) ================= */
double mbo_backward (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                     int mode, double* matrix) {
  lse_init();
  trans_index ix; index_build (&ix, m, 0);
  const int S = m->nStates;
  double* ninf = make_ninf (S);
  cells M = { matrix, Li, S, matrix == NULL, ninf };
  if (!matrix) M.c = (double*) malloc (sizeof (double) * 2 * (size_t) (Li + 1) * S);
  fill_ninf (M.c, (size_t) (matrix ? Lo + 1 : 2) * (size_t) (Li + 1) * S);
  for (int64_t o = Lo; o >= 0; --o) {
    const int endO = (o == Lo);
    const int outTok = endO ? 0 : y[o];
    for (int64_t i = Li; i >= 0; --i) {
      const int endI = (i == Li);
      const int inTok = endI ? 0 : x[i];
      double* cur = cellp (&M, i, o);
      if (!env_contains (i, o)) { for (int s = 0; s < S; ++s) cur[s] = -INFINITY; continue; }
      for (int s = S - 1; s >= 0; --s) {
        double ll = (endI && endO && s == S - 1) ? 0 : -INFINITY;
#define ACC(IN, OUT, DST)                                                           \
        do { const int64_t k_ = key_of (&ix, s, (IN), (OUT));                       \
             const double* d_ = (DST);                                              \
             for (int64_t p = ix.off[k_]; p < ix.off[k_ + 1]; ++p)                  \
               ll = lse2 (ll, d_[ix.other[p]] + ix.lw[p], mode); } while (0)
        if (!endI && !endO) ACC (inTok, outTok, cellr (&M, i + 1, o + 1));
        if (!endI) ACC (inTok, 0, cellr (&M, i + 1, o));
        if (!endO) ACC (0, outTok, cellr (&M, i, o + 1));
        ACC (0, 0, cur);
#undef ACC
        cur[s] = ll;
      }
    }
  }
  const double result = cellr (&M, 0, 0)[0];
  if (!matrix) free (M.c);
  free (ninf);
  index_free (&ix);
  return result;
}

/* ================= Viterbi (viterbi.cpp:18-47) + traceback (dpmatrix.defs.h:82-110) ================= */
double mbo_viterbi (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                    double* matrix, int32_t* path, int64_t pathCap, int64_t* pathLen) {
  trans_index ix; index_build (&ix, m, 1);
  const int S = m->nStates;
  const int needFull = (matrix != NULL) || (path != NULL);
  double* ninf = make_ninf (S);
  cells M = { matrix, Li, S, !needFull, ninf };
  if (!matrix) M.c = (double*) malloc (sizeof (double) * (size_t) (needFull ? Lo + 1 : 2) * (size_t) (Li + 1) * S);
  fill_ninf (M.c, (size_t) (needFull ? Lo + 1 : 2) * (size_t) (Li + 1) * S);
  for (int64_t o = 0; o <= Lo; ++o) {
    const int outTok = o ? y[o - 1] : 0;
    for (int64_t i = 0; i <= Li; ++i) {
      const int inTok = i ? x[i - 1] : 0;
      double* cur = cellp (&M, i, o);
      if (!env_contains (i, o)) { for (int d = 0; d < S; ++d) cur[d] = -INFINITY; continue; }
      for (int d = 0; d < S; ++d) {
        double ll = (i || o || d) ? -INFINITY : 0;
#define ACC(IN, OUT, SRC)                                                           \
        do { const int64_t k_ = key_of (&ix, d, (IN), (OUT));                       \
             const double* s_ = (SRC);                                              \
             for (int64_t p = ix.off[k_]; p < ix.off[k_ + 1]; ++p) {                \
               const double t_ = s_[ix.other[p]] + ix.lw[p];                        \
               ll = ll < t_ ? t_ : ll; } } while (0)      /* std::max (dpmatrix.h:122) */
        if (i && o) ACC (inTok, outTok, cellr (&M, i - 1, o - 1));
        if (i) ACC (inTok, 0, cellr (&M, i - 1, o));
        if (o) ACC (0, outTok, cellr (&M, i, o - 1));
        ACC (0, 0, cur);
#undef ACC
        cur[d] = ll;
      }
    }
  }
  const double score = cellr (&M, Li, Lo)[S - 1];
  double result = score;
  if (path) {
    int64_t n = 0;
    if (score > -INFINITY) {   /* boss.cpp:831 only traces finite scores; dpmatrix.defs.h:84 asserts it */
      int64_t i = Li, o = Lo;
      int s = S - 1;
      while (i > 0 || o > 0 || s != 0) {
        const int inTok = i ? x[i - 1] : 0, outTok = o ? y[o - 1] : 0;
        /* candidates in the order match, delete, insert, silent (dpmatrix.defs.h:93-99); the first
         * maximum wins (std::max_element, dpmatrix.defs.h:171-174) */
        double best = 0; int haveBest = 0; int64_t bestP = -1; int bestType = 0;
#define CAND(IN, OUT, SRC, TYPE)                                                    \
        do { const int64_t k_ = key_of (&ix, s, (IN), (OUT));                       \
             const double* s_ = (SRC);                                              \
             for (int64_t p = ix.off[k_]; p < ix.off[k_ + 1]; ++p) {                \
               const double t_ = s_[ix.other[p]] + ix.lw[p];                        \
               if (!haveBest || best < t_) { best = t_; haveBest = 1; bestP = p; bestType = (TYPE); } } } while (0)
        if (i && o) CAND (inTok, outTok, cellr (&M, i - 1, o - 1), 0);
        if (i) CAND (inTok, 0, cellr (&M, i - 1, o), 1);
        if (o) CAND (0, outTok, cellr (&M, i, o - 1), 2);
        CAND (0, 0, cellr (&M, i, o), 3);
#undef CAND
        if (!haveBest || n >= pathCap) { result = NAN; break; }
        path[n++] = ix.id[bestP];
        if (bestType == 0 || bestType == 1) --i;
        if (bestType == 0 || bestType == 2) --o;
        s = ix.other[bestP];
      }
      /* collected end -> start; reverse (the reference push_fronts, dpmatrix.defs.h:75) */
      for (int64_t a = 0, b = n - 1; a < b; ++a, --b) { const int32_t t = path[a]; path[a] = path[b]; path[b] = t; }
    }
    if (pathLen) *pathLen = n;
  }
  if (!matrix) free (M.c);
  free (ninf);
  index_free (&ix);
  return result;
}

/* ================= counts (counts.cpp:57-64, backward.cpp:62-87) ================= */
double mbo_counts (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                   int mode, double* counts, double* backLL) {
  const int S = m->nStates;
  const size_t n = (size_t) (Li + 1) * (size_t) (Lo + 1) * (size_t) S;
  double* F = (double*) malloc (sizeof (double) * n);
  double* B = (double*) malloc (sizeof (double) * n);
  const double fll = mbo_forward (m, x, Li, y, Lo, mode, F);
  const double ll = mbo_backward (m, x, Li, y, Lo, mode, B);   /* getCounts uses backward.logLike() (backward.cpp:66) */
  if (backLL) *backLL = ll;
  trans_index ix; index_build (&ix, m, 0);
  double* ninf = make_ninf (S);
  cells MF = { F, Li, S, 0, ninf }, MB = { B, Li, S, 0, ninf };
  for (int64_t o = Lo; o >= 0; --o) {
    const int endO = (o == Lo);
    const int outTok = endO ? 0 : y[o];
    for (int64_t i = Li; i >= 0; --i) {
      const int endI = (i == Li);
      const int inTok = endI ? 0 : x[i];
      if (!env_contains (i, o)) continue;   /* getCounts walks the envelope's cells only (backward.cpp:70) */
      for (int s = S - 1; s >= 0; --s) {
        const double logOdds = cellp (&MF, i, o)[s] - ll;
#define CNT(IN, OUT, DST)                                                           \
        do { const int64_t k_ = key_of (&ix, s, (IN), (OUT));                       \
             const double* d_ = (DST);                                              \
             for (int64_t p = ix.off[k_]; p < ix.off[k_ + 1]; ++p)                  \
               counts[ix.id[p]] += exp (logOdds + (d_[ix.other[p]] + ix.lw[p])); } while (0)
        if (!endI && !endO) CNT (inTok, outTok, cellr (&MB, i + 1, o + 1));
        if (!endI) CNT (inTok, 0, cellr (&MB, i + 1, o));
        if (!endO) CNT (0, outTok, cellr (&MB, i, o + 1));
        CNT (0, 0, cellr (&MB, i, o));
#undef CNT
      }
    }
  }
  index_free (&ix);
  free (F); free (B); free (ninf);
  return fll;
}

void mbo_synth (uint64_t seed, uint64_t pairIndex, int which, int64_t len, int nSym, uint8_t* tokens) {
  for (int64_t p = 0; p < len; ++p)
    tokens[p] = (uint8_t) mb_synth_token (seed, pairIndex, which, (uint64_t) p, nSym);
}
