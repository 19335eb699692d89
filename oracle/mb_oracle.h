/* oracle/mb_oracle.h -- TEST INFRASTRUCTURE (not product code).
 *
 * Plain-C, single-threaded restatement of the reference's Forward / Backward / Viterbi /
 * traceback / posterior-count recurrences on a FLAT evaluated machine (the same flat form the
 * product's C ABI takes, include/machineboss_b200.h).  Each function cites the reference code it
 * follows.  Pinned against (a) the reference's own golden vectors under t/expect and (b) the
 * reference itself, compiled unmodified into oracle/_ref/refdrv -- see tests/test_oracle.py and
 * tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 */
#ifndef MB_ORACLE_H
#define MB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat evaluated machine.  Transition t (0 <= t < nTrans) is listed in the reference's enumeration
 * order (src/eval.cpp:49-69): source state ascending, then position in the source's TransList, so
 * t = transOffset[src] + transIndex.  Token 0 is epsilon (src/eval.h:25); input tokens are
 * 1..nInTok, output tokens 1..nOutTok.  Start state is 0, end state nStates-1 (eval.cpp:76-84). */
typedef struct {
  int32_t nStates, nInTok, nOutTok;
  int64_t nTrans;
  const int32_t* src;
  const int32_t* dst;
  const int32_t* in;
  const int32_t* out;
  const double* logWeight;
} mbo_machine;

/* lse_mode: 0 = the reference's lookup table with interpolation (logsumexp.h:48-90),
 *           1 = exact log(1+exp(-x)) (the reference's LOG_SUM_EXP_SLOW build). */
#define MBO_LSE_TABLE 0
#define MBO_LSE_EXACT 1

/* Matrices, when requested (non-NULL), are (Lo+1)*(Li+1)*nStates doubles laid out
 * [outPos][inPos][state], the reference's IdentityIndexMapper order (dpmatrix.h:38-40,89-95). */

/* forward.defs.h:22-55.  Returns cell(Li,Lo,nStates-1). */
double mbo_forward (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                    int lse_mode, double* matrix);

/* backward.cpp:18-50.  Returns cell(0,0,0). */
double mbo_backward (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                     int lse_mode, double* matrix);

/* viterbi.cpp:18-47 + dpmatrix.defs.h:82-110,171-174.  Returns the Viterbi score; if path != NULL
 * and the score is finite, writes the traceback as global transition ids, start -> end, into
 * path[0..*pathLen) (capacity pathCap; returns NaN if it does not fit). */
double mbo_viterbi (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                    double* matrix, int32_t* path, int64_t pathCap, int64_t* pathLen);

/* counts.cpp:57-64 + backward.cpp:62-87.  Adds this pair's posterior transition counts into
 * counts[nTrans]; returns forward.logLike(); *backLL (optional) receives backward.logLike(). */
double mbo_counts (const mbo_machine* m, const uint8_t* x, int64_t Li, const uint8_t* y, int64_t Lo,
                   int lse_mode, double* counts, double* backLL);

/* Envelope (src/seqpair.h:75-113): cells of output row o exist only for inStart[o] <= inPos < inEnd[o];
 * reads outside return -inf (dpmatrix.h:142-144) and the fills only visit cells inside
 * (forward.defs.h:30, backward.cpp:26, viterbi.cpp:24).  Call with envStart == NULL for the full
 * matrix.  Sets the envelope used by all subsequent mbo_* calls on this thread. */
void mbo_set_envelope (const int64_t* envStart, const int64_t* envEnd);

/* Synthetic tokens (oracle/synth.h). */
void mbo_synth (uint64_t seed, uint64_t pairIndex, int which, int64_t len, int nSym, uint8_t* tokens);

/* log_sum_exp itself, for unit tests (logsumexp.h:72-90). */
double mbo_log_sum_exp (double a, double b, int lse_mode);

#ifdef __cplusplus
}
#endif
#endif
