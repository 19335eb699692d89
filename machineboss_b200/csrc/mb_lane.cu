// mb_lane.cu -- the lane engine: Forward log-likelihood and Viterbi (+ traceback) of batches WITHOUT
// input sequences (generator machines: a profile HMM, or a profile composed with an error model,
// scoring a million reads -- SURVEY.md section 8, config 5) on machines of any size.
//
//   reference                                              here
//   MappedForwardMatrix::fill / logLike  forward.defs.h:22-55   lane_kernel<OP_SUM>  (scaled linear domain)
//                                                               lane_kernel<OP_LSE>  (log domain: flagged reads, extreme weights)
//   ViterbiMatrix::fill / logLike        viterbi.cpp:18-47      lane_kernel<OP_MAX>  (FP64 add + compare: bit-exact)
//   DPMatrix::traceBack                  dpmatrix.defs.h:82-110 wide_traceback_kernel over the lane-interleaved pointers
//
// Why another mapping.  With no input sequence the matrix of a read is one column of Lo+1 cells, and
// a cell is a sparse matrix-vector product through the silent transitions in dependency order.  A
// profile HMM's delete chain makes that order hundreds of levels deep (488 levels for PF00516, 979
// once composed with protpsw), so spreading the STATES of a cell over lanes (mb_wide.cu) leaves the
// lanes waiting on one another.  Here every LANE owns a READ instead: the 32 reads of a warp walk
// the same transition program in lockstep -- destination states in index order, which is a
// topological order of the silent transitions (eval.cpp:44 requires an advancing machine) -- one
// multiply-add per transition per lane, no barrier anywhere.  The program is a flat stream of
// 16-byte records read through uniform (broadcast) loads; the emitting transitions of a
// destination are stored as the union over output tokens with a token-indexed weight table, so
// that reads with different residues still share the stream.  State vectors live in global memory
// interleaved by lane, v[state][lane], so every load and store of the warp is one 256-byte line;
// two vectors per lane (previous cell, this cell).
//
// Arithmetic as in mb_wide.cu.  Forward: probabilities with one power-of-two frame per cell and
// lane; the previous cell enters through one exact power-of-two factor; a cell whose values span
// more than 2^600 flags the read, and flagged reads are re-run in the log domain.  Viterbi:
// log-weights, FP64 add + strict '<' in the reference's candidate order (insert sources, then
// silent sources, each by ascending source state then transition index), so scores and paths are
// identical to the reference's.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "mb_internal.h"

namespace mb {

enum { L_SUM = 0, L_MAX = 1, L_LSE = 2 };
#define L_SENT (-(1 << 29))
#define L_SPREAD 600
// record control word: low bits flags, high 16 bits the candidate index (silent terms, Viterbi)
#define L_EMIT 1u      // term reads the previous cell; w holds the row of the token-indexed tables
#define L_SCALE 2u     // last emitting term of its destination: bring the sum into this cell's frame
#define L_END 4u       // last term of its destination: store the state
#define L_NOTERM 8u    // a destination without incoming transitions

struct LRec { double w; uint32_t src; uint32_t ctl; };      // src: source state * 32 (element offset in a lane-interleaved vector)

struct LHost {
  std::vector<LRec> recLin, recLog;       // identical but for the weights
  std::vector<double> emLin, emLog;       // [row][nOut] token-indexed weights of the emitting terms
  std::vector<uint16_t> emIdx;            // [row][nOut] candidate index in the token-selected list (Viterbi back-pointer)
  std::vector<int64_t> recPerm;           // silent record n carries hInc entry recPerm[n] (-1 otherwise)
  std::vector<int64_t> emPerm;            // table entry -> hInc entry (-1: no such transition)
  LRec* dRecLin = nullptr; LRec* dRecLog = nullptr;
  double* dEmLin = nullptr; double* dEmLog = nullptr;
  uint16_t* dEmIdx = nullptr;
  int64_t nRec = 0;
  int S = 0, nOut = 0, bpBytes = 2;
  bool linearOk = false;
  int numSMs = 148;
};

struct LParams {
  const LRec* rec; int64_t nRec;
  const double* em; const uint16_t* emIdx;
  int32_t S, nOut, bpBytes;
  DevBatch b;
  const int64_t* order; int64_t nWork;       // reads, longest first; task n = reads [32n, 32n+32)
  unsigned long long* counter;
  double* result; int32_t* flag;
  double* vec;                                // per resident warp: 2 vectors of S * 32 doubles
  unsigned char* bp; const int64_t* bpOff;   // back-pointers of task n at bpOff[n]: [o][state][lane]
};

__device__ __forceinline__ double l_ninf() { return __longlong_as_double (0xfff0000000000000LL); }
__device__ __forceinline__ double l_pow2 (int e) { return e < -1022 ? 0. : __longlong_as_double ((long long) (e > 1023 ? 2046 : e + 1023) << 52); }
__device__ __forceinline__ double l_lse (double a, double b) {      // as w_lse in mb_wide.cu
  const double mx = fmax (a, b), mn = fmin (a, b);
  const float d = (float) (mn - mx);
  if (!(d > -40.f)) return mx;
  return mx + (double) __logf (1.f + __expf (d));
}

// R reads per lane (reads lane, lane+32, ... of the task): R independent multiply-add chains and R
// loads in flight per thread hide the latency of the state vectors, which stream from L2 / HBM, and
// every record of the program is fetched and decoded once for 32*R reads.
template<int OP, int R>
__global__ void __launch_bounds__(128) lane_kernel (const __grid_constant__ LParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  const int S = p.S;
  constexpr int LPT = 32 * R;      // reads (vector lanes) per task
  double* v0 = p.vec + ((size_t) blockIdx.x * nWarps + warp) * 2 * (size_t) S * LPT + lane;
  double* v1 = v0 + (size_t) S * LPT;
  const double ZERO = OP == L_SUM ? 0. : l_ninf(), ONE = OP == L_SUM ? 1. : 0.;
  const unsigned kb = p.bpBytes == 1 ? 6 : 14;
  const double LN2 = 0.693147180559945309417232121458;
  for (;;) {
    long long task = 0;
    if (lane == 0) task = (long long) atomicAdd (p.counter, 1ULL);
    task = __shfl_sync (0xffffffffu, task, 0);
    if (task * LPT >= p.nWork) break;
    int64_t k[R];
    const uint8_t* y[R];
    int Lo[R], maxLo = -1;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const long long rd = task * LPT + q * 32 + lane;
      const bool have = rd < p.nWork;
      k[q] = have ? p.order[rd] : 0;
      y[q] = p.b.y + p.b.yOff[k[q]];
      Lo[q] = have ? (int) (p.b.yOff[k[q] + 1] - p.b.yOff[k[q]]) : -1;
      maxLo = max (maxLo, Lo[q]);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) maxLo = max (maxLo, __shfl_xor_sync (0xffffffffu, maxLo, off));
    unsigned char* bp = OP == L_MAX && p.bp ? p.bp + p.bpOff[task] + (size_t) lane * p.bpBytes : nullptr;
    bool bad[R];
    int Fprev[R], Gprev[R];
#pragma unroll
    for (int q = 0; q < R; ++q) { bad[q] = false; Fprev[q] = 0; Gprev[q] = L_SENT; }
    for (int o = 0; o <= maxLo; ++o) {
      double* cur = (o & 1) ? v1 : v0;
      const double* prev = (o & 1) ? v0 : v1;
      int tok[R], F[R], mx[R];
      unsigned mn[R], best[R];
      double f[R], acc[R], res[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        tok[q] = (o <= Lo[q] && o > 0) ? y[q][o - 1] - 1 : 0;
        F[q] = 0; f[q] = 1.;
        if (OP == L_SUM && o > 0) { if (Gprev[q] == L_SENT) f[q] = 0.; else { F[q] = Gprev[q]; f[q] = l_pow2 (Fprev[q] - F[q]); } }
        mx[q] = 0; mn[q] = 0xffffffffu; best[q] = 0xffffu;
        acc[q] = o == 0 ? ONE : ZERO;      // the origin cell's start state (forward.defs.h:36)
        res[q] = ZERO;
      }
      const bool usePrev = o > 0;
      double* out = cur;
      unsigned char* bpRow = bp ? bp + (size_t) o * S * LPT * p.bpBytes : nullptr;
      const LRec* r = p.rec;
      uint4 nx = __ldg (reinterpret_cast<const uint4*> (r));
      for (int64_t n = 0; n < p.nRec; ++n) {
        const uint4 u = nx;
        nx = __ldg (reinterpret_cast<const uint4*> (r + n + 1));      // the stream ends with a spare record
        const unsigned ctl = u.w;
        if (!(ctl & L_NOTERM)) {
          if (ctl & L_EMIT) {
            if (usePrev) {
              double x[R], w[R];
              unsigned ix[R];
#pragma unroll
              for (int q = 0; q < R; ++q) {
                x[q] = prev[(size_t) u.z * R + q * 32];
                w[q] = __ldg (p.em + (size_t) u.x * p.nOut + tok[q]);
                if (OP == L_MAX) ix[q] = __ldg (p.emIdx + (size_t) u.x * p.nOut + tok[q]);
              }
#pragma unroll
              for (int q = 0; q < R; ++q) {
                if (OP == L_SUM) acc[q] = fma (w[q], x[q], acc[q]);
                else if (OP == L_LSE) acc[q] = l_lse (acc[q], x[q] + w[q]);
                else { const double c = x[q] + w[q]; if (acc[q] < c) { acc[q] = c; best[q] = ((unsigned) T_INSERT << kb) | ix[q]; } }
              }
            }
            if (OP == L_SUM && (ctl & L_SCALE)) {
#pragma unroll
              for (int q = 0; q < R; ++q) acc[q] *= f[q];
            }
          } else {
            const double w = __hiloint2double ((int) u.y, (int) u.x);
            double x[R];
#pragma unroll
            for (int q = 0; q < R; ++q) x[q] = cur[(size_t) u.z * R + q * 32];
#pragma unroll
            for (int q = 0; q < R; ++q) {
              if (OP == L_SUM) acc[q] = fma (w, x[q], acc[q]);
              else if (OP == L_LSE) acc[q] = l_lse (acc[q], x[q] + w);
              else { const double c = x[q] + w; if (acc[q] < c) { acc[q] = c; best[q] = ((unsigned) T_SILENT << kb) | (ctl >> 16); } }
            }
          }
        }
        if (ctl & L_END) {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            out[q * 32] = acc[q];
            if (OP == L_SUM) {
              const int hi = __double2hiint (acc[q]);
              mx[q] = max (mx[q], hi);
              mn[q] = min (mn[q], (unsigned) (hi - 0x00100000));      // zeros and denormals wrap to the top and drop out
            } else if (OP == L_MAX && bpRow) {
              if (p.bpBytes == 1) bpRow[q * 32] = (unsigned char) best[q]; else reinterpret_cast<uint16_t*> (bpRow)[q * 32] = (uint16_t) best[q];
            }
            res[q] = acc[q];      // after the last record: the end state
            acc[q] = ZERO;
            best[q] = 0xffffu;
          }
          out += LPT;
          if (OP == L_MAX && bpRow) bpRow += LPT * p.bpBytes;
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (OP == L_SUM) {
          int Gc = L_SENT;
          if (mx[q] >= 0x00100000) {
            const int emx = mx[q] >> 20;
            Gc = F[q] + emx - 1023;
            if (o <= Lo[q] && (emx == 0x7ff || (mn[q] != 0xffffffffu && emx - (int) ((mn[q] >> 20) + 1) > L_SPREAD))) bad[q] = true;
          }
          Fprev[q] = F[q]; Gprev[q] = Gc;
        }
        if (o == Lo[q]) {
          if (OP == L_SUM) { p.result[k[q]] = res[q] > 0. ? log (res[q]) + F[q] * LN2 : l_ninf(); p.flag[k[q]] = bad[q] || !(res[q] > 0.) || !(res[q] < 1e300); }
          else p.result[k[q]] = res[q];
        }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static LHost* lh (const mb_machine* m) { return static_cast<LHost*> (m->lane); }

static void lane_fill_weights (const mb_machine* m, LHost* h) {
  bool ok = true;
  const double lim = 30. * 0.6931471805599453;
  auto check = [&] (double lw) { if (std::isnan (lw) || (std::isfinite (lw) && fabs (lw) > lim) || lw == INFINITY) ok = false; };
  for (size_t n = 0; n < h->recPerm.size(); ++n) {
    if (h->recPerm[n] < 0) continue;
    const double lw = m->hInc.lw[h->recPerm[n]];
    h->recLog[n].w = lw;
    h->recLin[n].w = exp (lw);
    check (lw);
  }
  for (size_t n = 0; n < h->emPerm.size(); ++n) {
    if (h->emPerm[n] < 0) { h->emLin[n] = 0.; h->emLog[n] = -INFINITY; continue; }
    const double lw = m->hInc.lw[h->emPerm[n]];
    h->emLog[n] = lw;
    h->emLin[n] = exp (lw);
    check (lw);
  }
  h->linearOk = ok;
}

void lane_destroy (mb_machine* m) {
  LHost* h = lh (m);
  if (!h) return;
  for (void* p: { (void*) h->dRecLin, (void*) h->dRecLog, (void*) h->dEmLin, (void*) h->dEmLog, (void*) h->dEmIdx }) if (p) cudaFree (p);
  delete h;
  m->lane = nullptr;
}

static int lane_upload (mb_machine* m, LHost* h, bool all) {
  MB_CUDA (cudaSetDevice (m->device));
  const size_t rb = h->recLin.size() * sizeof (LRec), eb = std::max<size_t> (h->emLin.size(), 1) * 8;
  if (all) {
    MB_CUDA (cudaMalloc (&h->dRecLin, rb)); MB_CUDA (cudaMalloc (&h->dRecLog, rb));
    MB_CUDA (cudaMalloc (&h->dEmLin, eb)); MB_CUDA (cudaMalloc (&h->dEmLog, eb));
    MB_CUDA (cudaMalloc (&h->dEmIdx, std::max<size_t> (h->emIdx.size(), 1) * 2));
    if (!h->emIdx.empty()) MB_CUDA (cudaMemcpy (h->dEmIdx, h->emIdx.data(), h->emIdx.size() * 2, cudaMemcpyHostToDevice));
  }
  MB_CUDA (cudaMemcpy (h->dRecLin, h->recLin.data(), rb, cudaMemcpyHostToDevice));
  MB_CUDA (cudaMemcpy (h->dRecLog, h->recLog.data(), rb, cudaMemcpyHostToDevice));
  if (!h->emLin.empty()) {
    MB_CUDA (cudaMemcpy (h->dEmLin, h->emLin.data(), h->emLin.size() * 8, cudaMemcpyHostToDevice));
    MB_CUDA (cudaMemcpy (h->dEmLog, h->emLog.data(), h->emLog.size() * 8, cudaMemcpyHostToDevice));
  }
  return 0;
}

int lane_update_weights (mb_machine* m) {
  LHost* h = lh (m);
  if (!h) return 0;
  lane_fill_weights (m, h);
  return lane_upload (m, h, false);
}

// The transition program: destinations in index order; per destination the emitting terms (union
// over output tokens of the token-selected insert lists, in list order), then the silent terms.
int lane_prepare (mb_machine* m) {
  LHost* h = new LHost;
  m->lane = h;
  const int S = m->S, nOut = m->nOut, nIn1 = m->nIn + 1, nOut1 = nOut + 1;
  const HostCsr& inc = m->hInc;
  h->S = S; h->nOut = nOut;
  int maxList = 0;
  LRec blank; blank.w = 0; blank.src = 0; blank.ctl = 0;
  for (int d = 0; d < S; ++d) {
    const size_t first = h->recLin.size();
    // union of the insert lists of d over the output tokens: key (source, rank among equals) keeps each list's order
    std::map<std::pair<int, int>, int> rowOf;      // -> table row
    std::vector<std::pair<int, int>> keys;
    for (int c = 1; c <= nOut; ++c) {
      const int64_t key = ((int64_t) d * nIn1) * nOut1 + c;
      maxList = std::max<int> (maxList, (int) (inc.off[key + 1] - inc.off[key]));
      std::map<int, int> seen;
      for (int64_t q = inc.off[key]; q < inc.off[key + 1]; ++q) {
        const std::pair<int, int> kk (inc.other[q], seen[inc.other[q]]++);
        if (!rowOf.count (kk)) { rowOf[kk] = -1; keys.push_back (kk); }
      }
    }
    std::sort (keys.begin(), keys.end());
    for (auto& kk: keys) {
      const int row = (int) (h->emPerm.size() / std::max (nOut, 1));
      rowOf[kk] = row;
      h->emPerm.resize (h->emPerm.size() + nOut, -1);
      h->emIdx.resize (h->emIdx.size() + nOut, 0x3fff);
      LRec r = blank;
      const uint64_t rowBits = (uint64_t) row;
      memcpy (&r.w, &rowBits, 8);
      r.src = (uint32_t) kk.first * 32u;
      r.ctl = L_EMIT;
      h->recLin.push_back (r);
      h->recPerm.push_back (-1);
    }
    for (int c = 1; c <= nOut; ++c) {
      const int64_t key = ((int64_t) d * nIn1) * nOut1 + c;
      std::map<int, int> seen;
      for (int64_t q = inc.off[key]; q < inc.off[key + 1]; ++q) {
        const int row = rowOf[std::make_pair (inc.other[q], seen[inc.other[q]]++)];
        h->emPerm[(size_t) row * nOut + (c - 1)] = q;
        h->emIdx[(size_t) row * nOut + (c - 1)] = (uint16_t) (q - inc.off[key]);
      }
    }
    if (!keys.empty()) h->recLin.back().ctl |= L_SCALE;
    if (d > 0) {      // state 0's only possible silent source is its own self-loop, which contributes nothing (machine.cpp:759)
      const int64_t key = (int64_t) d * nIn1 * nOut1;
      maxList = std::max<int> (maxList, (int) (inc.off[key + 1] - inc.off[key]));
      for (int64_t q = inc.off[key]; q < inc.off[key + 1]; ++q) {
        LRec r = blank;
        r.src = (uint32_t) inc.other[q] * 32u;
        r.ctl = (uint32_t) (q - inc.off[key]) << 16;
        h->recLin.push_back (r);
        h->recPerm.push_back (q);
      }
    }
    if (h->recLin.size() == first) { LRec r = blank; r.ctl = L_NOTERM; h->recLin.push_back (r); h->recPerm.push_back (-1); }
    h->recLin.back().ctl |= L_END;
  }
  if (maxList > 16382) { set_error ("lane engine: a transition list has more than 16382 entries"); return 1; }
  h->bpBytes = maxList <= 63 ? 1 : 2;
  h->nRec = (int64_t) h->recLin.size();
  { LRec r = blank; r.ctl = L_NOTERM; h->recLin.push_back (r); h->recPerm.push_back (-1); }      // spare record: the sweep keeps one in flight
  h->recLog = h->recLin;
  h->emLin.assign (h->emPerm.size(), 0.);
  h->emLog.assign (h->emPerm.size(), -INFINITY);
  lane_fill_weights (m, h);
  if (lane_upload (m, h, true)) return 1;
  MB_CUDA (cudaDeviceGetAttribute (&h->numSMs, cudaDevAttrMultiProcessorCount, m->device));
  if (m->opt.get ("verbose", 0))
    fprintf (stderr, "lane engine: S=%d records=%lld (emitting rows %zu x %d tokens), bp %d bytes\n", S, (long long) h->nRec, h->emPerm.size() / std::max (nOut, 1), nOut, h->bpBytes);
  return 0;
}

struct LBuf {
  void* p = nullptr;
  ~LBuf() { if (p) cudaFree (p); }
  int alloc (size_t bytes) { MB_CUDA (cudaMalloc (&p, bytes ? bytes : 8)); return 0; }
  template<class T> T* as() { return (T*) p; }
};

// reads per lane.  Measured on B200 (PF00516, reads of 275): with enough reads to give every SM some 40
// warps, one read per lane is fastest (262 144 reads: 175 GCUPS Forward, 132 Viterbi); with fewer, a second
// independent chain per thread makes up for the missing warps in the sums (65 536 reads: 40 -> 89 GCUPS),
// while the max-plus sweep, which carries a pointer per chain, stays at one.
static int lane_reads_per_lane (const mb_machine* m, const LHost* h, int64_t nWork, int op) {
  if (m->opt.has ("lane_r")) { const int r = m->opt.get ("lane_r", 1); return r >= 4 ? 4 : r >= 2 ? 2 : 1; }
  if (op == L_MAX || nWork >= (int64_t) h->numSMs * 40 * 32) return 1;
  return nWork >= (int64_t) h->numSMs * 8 * 64 ? 2 : 1;
}

template<int OP, int R>
static int lane_launch_r (mb_machine* m, mb_batch* b, const std::vector<int64_t>& order, double* dResult, int32_t* dFlag,
                          unsigned char* dBp, const int64_t* dBpOff) {
  LHost* h = lh (m);
  constexpr int LPT = 32 * R;
  const int64_t nWork = (int64_t) order.size(), nTasks = (nWork + LPT - 1) / LPT;
  int ctas = 0;
  MB_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&ctas, lane_kernel<OP, R>, 128, 0));
  int warpsPerSM = 48;      // the sweep is bound by the latency of its state-vector loads: as many warps as the registers allow
  if (m->opt.has ("lane_warps")) warpsPerSM = std::max (4, m->opt.get ("lane_warps", 48));
  ctas = std::max (1, std::min (ctas, warpsPerSM / 4));
  // few tasks: one warp per CTA spreads them over the SMs
  const int threads = nTasks >= (int64_t) ctas * h->numSMs * 4 ? 128 : nTasks >= (int64_t) h->numSMs * 2 ? 64 : 32;
  const int wpc = threads / 32;
  const int grid = (int) std::max<int64_t> (1, std::min<int64_t> ((nTasks + wpc - 1) / wpc, (int64_t) ctas * h->numSMs));
  // scratch from the batch's grow-only workspace: a repeated call allocates nothing
  b->wsOrderHoldsFull = false;
  int64_t* dOrder = (int64_t*) ws_reserve (b, WS_ORDER, order.size() * 8);
  unsigned long long* dCounter = (unsigned long long*) ws_reserve (b, WS_COUNTER, 8);
  double* dVec = (double*) ws_reserve (b, WS_BND, (size_t) grid * wpc * 2 * h->S * LPT * 8);
  if (!dOrder || !dCounter || !dVec) return 1;
  MB_CUDA (cudaMemcpyAsync (dOrder, order.data(), order.size() * 8, cudaMemcpyHostToDevice, b->stream));
  MB_CUDA (cudaMemsetAsync (dCounter, 0, 8, b->stream));
  LParams p {};
  p.rec = OP == L_SUM ? h->dRecLin : h->dRecLog; p.nRec = h->nRec;
  p.em = OP == L_SUM ? h->dEmLin : h->dEmLog; p.emIdx = h->dEmIdx;
  p.S = h->S; p.nOut = h->nOut; p.bpBytes = h->bpBytes;
  p.b = b->dev;
  p.order = dOrder; p.nWork = nWork; p.counter = dCounter;
  p.result = dResult; p.flag = dFlag; p.vec = dVec;
  p.bp = dBp; p.bpOff = dBpOff;
  lane_kernel<OP, R><<<grid, threads, 0, b->stream>>> (p);
  MB_CUDA (cudaGetLastError());
  return 0;
}

template<int OP>
static int lane_launch (mb_machine* m, mb_batch* b, int R, const std::vector<int64_t>& order, double* dResult, int32_t* dFlag,
                        unsigned char* dBp, const int64_t* dBpOff) {
  return R == 4 ? lane_launch_r<OP, 4> (m, b, order, dResult, dFlag, dBp, dBpOff)
       : R == 2 ? lane_launch_r<OP, 2> (m, b, order, dResult, dFlag, dBp, dBpOff)
                : lane_launch_r<OP, 1> (m, b, order, dResult, dFlag, dBp, dBpOff);
}

static std::vector<int64_t> lane_order (const mb_batch* b, const std::vector<int64_t>* subset) {
  std::vector<int64_t> order;
  if (subset) order = *subset;
  else { order.resize ((size_t) b->nPairs); for (int64_t k = 0; k < b->nPairs; ++k) order[k] = k; }
  std::stable_sort (order.begin(), order.end(), [&] (int64_t a, int64_t c) { return b->yOff[a + 1] - b->yOff[a] > b->yOff[c + 1] - b->yOff[c]; });
  return order;
}

bool lane_wanted (const mb_machine* m, const mb_batch* b) {
  if (b->hasEnv || m->opt.get ("no_lane", 0)) return false;
  for (int64_t k = 0; k < b->nPairs; ++k) if (b->xOff[k + 1] != b->xOff[k]) return false;
  return true;
}

int lane_forward (mb_machine* m, mb_batch* b, double* loglike) {
  b->lastRedo = 0;
  if (b->nPairs == 0) return 0;
  LHost* h = lh (m);
  const std::vector<int64_t> order = lane_order (b, nullptr);
  LBuf dRes, dFlag;
  if (dRes.alloc ((size_t) b->nPairs * 8) || dFlag.alloc ((size_t) b->nPairs * 4)) return 1;
  MB_CUDA (cudaMemsetAsync (dFlag.p, 0, (size_t) b->nPairs * 4, b->stream));
  if (timing_begin (b)) return 1;
  int64_t launches = 1;
  const int R = lane_reads_per_lane (m, h, b->nPairs, L_SUM);
  if (h->linearOk) {
    if (lane_launch<L_SUM> (m, b, R, order, dRes.as<double>(), dFlag.as<int32_t>(), nullptr, nullptr)) return 1;
    std::vector<int32_t> flag ((size_t) b->nPairs);
    MB_CUDA (cudaMemcpyAsync (flag.data(), dFlag.p, flag.size() * 4, cudaMemcpyDeviceToHost, b->stream));
    MB_CUDA (cudaStreamSynchronize (b->stream));
    std::vector<int64_t> redo;
    for (int64_t k = 0; k < b->nPairs; ++k) if (flag[k]) redo.push_back (k);
    if (!redo.empty()) {
      if (lane_launch<L_LSE> (m, b, lane_reads_per_lane (m, h, (int64_t) redo.size(), L_LSE), lane_order (b, &redo), dRes.as<double>(), dFlag.as<int32_t>(), nullptr, nullptr)) return 1;
      ++launches;
    }
    b->lastRedo = (int64_t) redo.size();
  } else if (lane_launch<L_LSE> (m, b, R, order, dRes.as<double>(), dFlag.as<int32_t>(), nullptr, nullptr)) return 1;
  if (timing_end (b, launches)) return 1;
  MB_CUDA (cudaMemcpy (loglike, dRes.p, (size_t) b->nPairs * 8, cudaMemcpyDeviceToHost));
  return 0;
}

// traceback over lane-interleaved back-pointers (mb_wide.cu)
int wide_traceback_launch (mb_machine* m, mb_batch* b, const int64_t* dOrder, int64_t nWork, const unsigned char* dBp, const int64_t* dBpOff,
                           int bpBytes, int laneLayout, const double* dScore, int64_t* dLen, int32_t* dOut, const int64_t* dOutOff);

int lane_viterbi (mb_machine* m, mb_batch* b, double* score, int64_t* pathLen) {
  b->pathStart.clear();
  b->pathLen.clear();
  if (b->nPairs == 0) return 0;
  LHost* h = lh (m);
  const bool trace = pathLen != nullptr;
  const std::vector<int64_t> order = lane_order (b, nullptr);
  LBuf dRes;
  if (dRes.alloc ((size_t) b->nPairs * 8)) return 1;
  if (!trace) {
    if (timing_begin (b)) return 1;
    if (lane_launch<L_MAX> (m, b, lane_reads_per_lane (m, h, b->nPairs, L_MAX), order, dRes.as<double>(), nullptr, nullptr, nullptr)) return 1;
    if (timing_end (b, 1)) return 1;
    MB_CUDA (cudaMemcpy (score, dRes.p, (size_t) b->nPairs * 8, cudaMemcpyDeviceToHost));
    return 0;
  }
  // chunks of whole tasks (LPT reads) whose back-pointers, (maxLo+1) * S * LPT bytes or half-words per task, fit in free memory
  const int R = lane_reads_per_lane (m, h, b->nPairs, L_MAX), LPT = 32 * R;
  size_t freeB = 0, totalB = 0;
  MB_CUDA (cudaMemGetInfo (&freeB, &totalB));
  const double vecBytes = (double) h->numSMs * 48 * 2 * h->S * LPT * 8;
  const double budget = 0.75 * ((double) freeB - vecBytes);
  b->pathStart.assign ((size_t) b->nPairs, 0);
  b->pathLen.assign ((size_t) b->nPairs, 0);
  int64_t packed = 0, launches = 0;
  double ms = 0;
  for (size_t c0 = 0; c0 < order.size();) {
    std::vector<int64_t> bpOff;
    double bytes = 0;
    size_t c1 = c0;
    while (c1 < order.size()) {
      const int64_t k = order[c1];      // the task's longest read
      const double need = (double) (b->yOff[k + 1] - b->yOff[k] + 1) * h->S * LPT * h->bpBytes;
      if (need > budget) { set_error ("a task of reads needs more device memory for its back-pointers than is free (" + std::to_string (need) + " bytes)"); return 1; }
      if (c1 > c0 && bytes + need > budget) break;
      bpOff.push_back ((int64_t) bytes);
      bytes += need;
      c1 = std::min (order.size(), c1 + (size_t) LPT);
    }
    std::vector<int64_t> chunk (order.begin() + c0, order.begin() + c1);
    c0 = c1;
    std::vector<int64_t> bpOffRead (chunk.size());      // the traceback kernel indexes by read
    for (size_t n = 0; n < chunk.size(); ++n) bpOffRead[n] = bpOff[n / LPT];
    LBuf dBp, dBpOff, dBpOffRead, dOrder, dLen, dOutOff;
    if (dBp.alloc ((size_t) bytes) || dBpOff.alloc (bpOff.size() * 8) || dBpOffRead.alloc (chunk.size() * 8) || dOrder.alloc (chunk.size() * 8) || dLen.alloc (chunk.size() * 8)) return 1;
    MB_CUDA (cudaMemcpyAsync (dBpOff.p, bpOff.data(), bpOff.size() * 8, cudaMemcpyHostToDevice, b->stream));
    MB_CUDA (cudaMemcpyAsync (dBpOffRead.p, bpOffRead.data(), bpOffRead.size() * 8, cudaMemcpyHostToDevice, b->stream));
    MB_CUDA (cudaMemcpyAsync (dOrder.p, chunk.data(), chunk.size() * 8, cudaMemcpyHostToDevice, b->stream));
    if (timing_begin (b)) return 1;
    if (lane_launch<L_MAX> (m, b, R, chunk, dRes.as<double>(), nullptr, dBp.as<unsigned char>(), dBpOff.as<int64_t>())) return 1;
    const int64_t nWork = (int64_t) chunk.size();
    if (wide_traceback_launch (m, b, dOrder.as<int64_t>(), nWork, dBp.as<unsigned char>(), dBpOffRead.as<int64_t>(), h->bpBytes, LPT,
                               dRes.as<double>(), dLen.as<int64_t>(), nullptr, nullptr)) return 1;
    std::vector<int64_t> len (chunk.size()), off (chunk.size());
    MB_CUDA (cudaMemcpyAsync (len.data(), dLen.p, len.size() * 8, cudaMemcpyDeviceToHost, b->stream));
    MB_CUDA (cudaStreamSynchronize (b->stream));
    for (size_t n = 0; n < len.size(); ++n) {
      off[n] = packed;
      b->pathStart[chunk[n]] = packed;
      b->pathLen[chunk[n]] = len[n];
      packed += len[n];
    }
    if (paths_reserve (b, packed)) return 1;
    if (dOutOff.alloc (off.size() * 8)) return 1;
    MB_CUDA (cudaMemcpyAsync (dOutOff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, b->stream));
    if (wide_traceback_launch (m, b, dOrder.as<int64_t>(), nWork, dBp.as<unsigned char>(), dBpOffRead.as<int64_t>(), h->bpBytes, LPT,
                               dRes.as<double>(), dLen.as<int64_t>(), b->dPaths, dOutOff.as<int64_t>())) return 1;
    launches += 3;
    if (timing_end (b, launches)) return 1;
    ms += b->lastMs;
  }
  b->lastMs = ms;
  b->lastLaunches = launches;
  MB_CUDA (cudaMemcpy (score, dRes.p, (size_t) b->nPairs * 8, cudaMemcpyDeviceToHost));
  for (int64_t k = 0; k < b->nPairs; ++k) pathLen[k] = b->pathLen[k];
  return 0;
}

}  // namespace mb
