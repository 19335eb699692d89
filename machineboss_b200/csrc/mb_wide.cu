// mb_wide.cu -- the wide engine: Forward log-likelihood and Viterbi (+ traceback) for machines too
// large for the register-resident strip kernels of mb_jit.cu (composed transducers with tens to
// thousands of states: SURVEY.md section 8 configs 4-5).
//
//   reference                                              here
//   MappedForwardMatrix::fill / logLike  forward.defs.h:22-55   wide_kernel<OP_SUM>  (scaled linear domain)
//   ViterbiMatrix::fill / logLike        viterbi.cpp:18-47      wide_kernel<OP_MAX>  (FP64 add + compare: bit-exact)
//   DPMatrix::traceBack                  dpmatrix.defs.h:82-110 wide_traceback_kernel (stored back-pointers)
//
// Mapping.  One WARP owns one input position (a matrix column); the states of a cell are spread
// over its lanes.  A CTA of W column warps (+1 loader warp) sweeps a strip of W columns down the
// output rows in a skew (warp j works on output row t-j at step t), so the three neighbour cells a
// cell needs -- (i,o-1) own previous step, (i-1,o) left neighbour's previous step, (i-1,o-1) left
// neighbour's step before that -- sit in a ring of 2-3 cells per column in SHARED memory, with one
// CTA barrier per step.  The last column of a strip hands its live states (sources of input-
// consuming transitions) to the next strip through an L2-resident row buffer that the loader warp
// streams back in one row ahead.
//
// Within a cell the work is a sparse matrix-vector product over the token-selected transition
// lists.  Transitions that consume a token only read neighbour cells, so they run as three flat,
// dependency-free passes (match, delete, insert: "jobs" = one destination state and its source list,
// dealt round-robin to the lanes); silent transitions follow in dependency levels, one __syncwarp
// apart; destinations with long source lists (a profile HMM's end state) are split over all 32
// lanes and combined with shuffles.  The job tables live in shared memory when they fit.
//
// Arithmetic.  Forward: probabilities, one FMA per transition; every cell carries a power-of-two
// frame (F) and the exponent of its largest value (G); a cell is computed in the frame
// max(G of its neighbours), the neighbour sums entering through one exact power-of-two factor each,
// so nothing is ever renormalised in place.  A cell whose values span more than 2^600, or whose
// neighbours' frames are that far apart, flags the pair; flagged pairs are re-run by the log-domain
// generic engine (mb_last_redo counts them).  Viterbi: log-weights, FP64 add + compare in the
// reference's candidate order with a strict '<' (first maximum wins), so scores and paths are
// identical; a back-pointer (kind, index in the token-selected list) is stored per cell-state.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "mb_internal.h"

namespace mb {

enum { OP_SUM = 0, OP_MAX = 1, OP_LSE = 2 };      // scaled linear sum, max-plus with back-pointers, log-sum-exp
#define W_SENT (-(1 << 29))      // G of an empty (all-zero) cell
#define W_SPREAD 600             // exponent spread that hands a pair to the log-domain engine
#define W_COOP_LEN 48            // silent source lists at least this long are split over the warp

struct WJob { uint16_t dst; uint16_t len; uint32_t start; };

struct WideTables {       // byte offsets into one blob (weights first: two blobs differ only there)
  uint32_t oW, oSrc, oJobOffM, oJobOffD, oJobOffI, oJobs, oSil, oSilCoop, oLive, bytes;
  int32_t S, nIn, nOut, nSilRounds, hasMatch, nLiveIn, bpBytes;
};

struct WParams {
  WideTables t;
  const char* blob;
  int32_t tabInSmem, W, R, oneD;
  DevBatch b;
  const int64_t* order;
  int64_t nWork;
  unsigned long long* counter;
  double* result;
  int32_t* flag;
  double* bnd;          // per CTA: 2 buffers of bndRows * nLiveIn doubles
  int2* bndFG;          // per CTA: 2 buffers of bndRows (F, G)
  int64_t bndRows;
  unsigned char* bp;    // back-pointers, pair p at bpOff[p] (bytes), layout [o][i][s]
  const int64_t* bpOff; // indexed by work item
};

__device__ __forceinline__ double w_ninf() { return __longlong_as_double (0xfff0000000000000LL); }
__device__ __forceinline__ double w_pow2 (int e) {   // 2^e, 0 below the normal range
  return e < -1022 ? 0. : __longlong_as_double ((long long) (e > 1023 ? 2046 : e + 1023) << 52);
}

// log(exp(a)+exp(b)): max, difference and final add in FP64, softplus in FP32 (MUFU ex2/lg2);
// <= 1.2e-7 absolute per operation (the reference's table deviates from the exact function by 4.5e-5)
__device__ __forceinline__ double w_lse (double a, double b) {
  const double mx = fmax (a, b), mn = fmin (a, b);
  const float d = (float) (mn - mx);      // <= 0; -inf or NaN when an operand is -inf
  if (!(d > -40.f)) return mx;
  return mx + (double) __logf (1.f + __expf (d));
}

struct WTab {
  const double* w; const uint16_t* src; const uint32_t* jobOffM; const uint32_t* jobOffD; const uint32_t* jobOffI;
  const WJob* jobs; const WJob* sil; const unsigned char* silCoop; const uint16_t* live;
};

__device__ __forceinline__ WJob load_job (const WJob* p) {
  const uint2 r = *reinterpret_cast<const uint2*> (p);
  WJob j; j.dst = (uint16_t) (r.x & 0xffff); j.len = (uint16_t) (r.x >> 16); j.start = r.y;
  return j;
}

// one dependency-free pass over the jobs [j0, j1) reading neighbour cell `from`
template<int OP>
__device__ __forceinline__ void emit_pass (const WTab& T, uint32_t j0, uint32_t j1, const double* __restrict__ from, double* cur,
                                           uint16_t* bpS, double f, unsigned kindBits, int lane) {
  for (uint32_t q = j0 + lane; q < j1; q += 32) {
    const WJob jb = load_job (T.jobs + q);
    if (OP == OP_SUM) {
      double acc = 0;
      for (uint32_t p = jb.start, e = jb.start + jb.len; p < e; ++p) acc = fma (T.w[p], from[T.src[p]], acc);
      cur[jb.dst] = fma (f, acc, cur[jb.dst]);
    } else if (OP == OP_LSE) {
      double acc = cur[jb.dst];
      for (uint32_t p = jb.start, e = jb.start + jb.len; p < e; ++p) acc = w_lse (acc, from[T.src[p]] + T.w[p]);
      cur[jb.dst] = acc;
    } else {
      double best = cur[jb.dst];
      unsigned bp = bpS[jb.dst];
      for (uint32_t n = 0; n < jb.len; ++n) {
        const double v = from[T.src[jb.start + n]] + T.w[jb.start + n];
        if (best < v) { best = v; bp = kindBits | n; }      // strict: the first maximum wins (dpmatrix.defs.h:171-174)
      }
      cur[jb.dst] = best;
      bpS[jb.dst] = (uint16_t) bp;
    }
  }
  __syncwarp();
}

// One cell.  up/left/diag: neighbour cells (null when outside the matrix).  Linear domain: fU/fL/fD
// scale the neighbour sums into this cell's frame.  a, c: tokens (0 at the matrix edge).
template<int OP>
__device__ __forceinline__ void compute_cell (const WTab& T, const WideTables& t, double* cur, uint16_t* bpS,
                                              const double* up, const double* left, const double* diag,
                                              double fU, double fL, double fD, int a, int c, bool origin, int lane) {
  const int S = t.S;
  for (int d = lane; d < S; d += 32) { cur[d] = OP == OP_SUM ? 0. : w_ninf(); if (OP == OP_MAX) bpS[d] = 0xffff; }
  __syncwarp();
  if (origin && lane == 0) cur[0] = OP == OP_SUM ? 1. : 0.;
  __syncwarp();
  const unsigned kb = t.bpBytes == 1 ? 6 : 14;
  if (diag && t.hasMatch) { const int k = (a - 1) * t.nOut + (c - 1); emit_pass<OP> (T, T.jobOffM[k], T.jobOffM[k + 1], diag, cur, bpS, fD, (unsigned) T_MATCH << kb, lane); }
  if (left) emit_pass<OP> (T, T.jobOffD[a - 1], T.jobOffD[a], left, cur, bpS, fL, (unsigned) T_DELETE << kb, lane);
  if (up) emit_pass<OP> (T, T.jobOffI[c - 1], T.jobOffI[c], up, cur, bpS, fU, (unsigned) T_INSERT << kb, lane);
  // silent transitions, in dependency levels (sources are lower states already final)
  for (int r = 0; r < t.nSilRounds; ++r) {
    const WJob jb = load_job (T.sil + r * 32 + lane);
    const bool coop = T.silCoop[r];      // one destination, its source list cut into 32 contiguous pieces (lane order = list order)
    const bool own = jb.len && (!coop || lane == 0);
    if (OP != OP_MAX) {
      double acc = own ? cur[jb.dst] : (OP == OP_SUM ? 0. : w_ninf());
      if (OP == OP_SUM) for (uint32_t p = jb.start, e = jb.start + jb.len; p < e; ++p) acc = fma (T.w[p], cur[T.src[p]], acc);
      else for (uint32_t p = jb.start, e = jb.start + jb.len; p < e; ++p) acc = w_lse (acc, cur[T.src[p]] + T.w[p]);
      if (coop) {
        for (int off = 16; off; off >>= 1) { const double o2 = __shfl_down_sync (0xffffffffu, acc, off); acc = OP == OP_SUM ? acc + o2 : w_lse (acc, o2); }
        if (lane == 0) cur[jb.dst] = acc;
      } else if (jb.len) cur[jb.dst] = acc;
    } else {
      double best = own ? cur[jb.dst] : w_ninf();
      unsigned bp = own ? bpS[jb.dst] : 0xffffu;
      const uint32_t n0 = coop ? jb.start - __shfl_sync (0xffffffffu, jb.start, 0) : 0;
      for (uint32_t n = 0; n < jb.len; ++n) {
        const double v = cur[T.src[jb.start + n]] + T.w[jb.start + n];
        if (best < v) { best = v; bp = ((unsigned) T_SILENT << kb) | (n0 + n); }
      }
      if (coop) {
        for (int off = 16; off; off >>= 1) {      // lower lanes hold earlier candidates: a tie keeps the lower lane
          const double ob = __shfl_down_sync (0xffffffffu, best, off);
          const unsigned obp = __shfl_down_sync (0xffffffffu, bp, off);
          if (best < ob) { best = ob; bp = obp; }
        }
        if (lane == 0) { cur[jb.dst] = best; bpS[jb.dst] = (uint16_t) bp; }
      } else if (jb.len) { cur[jb.dst] = best; bpS[jb.dst] = (uint16_t) bp; }
    }
    __syncwarp();
  }
}

// exponent bookkeeping of a finished linear-domain cell: returns G (W_SENT if empty); sets bad on
// overflow or an exponent spread beyond W_SPREAD
__device__ __forceinline__ int cell_frame (const double* cur, int S, int F, int lane, bool& bad) {
  int mx = 0, mn = 0x7fffffff;
  bool dust = false;
  for (int d = lane; d < S; d += 32) {
    const double v = cur[d];
    const int hi = __double2hiint (v) & 0x7fffffff;
    mx = max (mx, hi);
    if (hi >= 0x00100000) mn = min (mn, hi);
    else if (v != 0.) dust = true;      // a denormal: something is underflowing
  }
  mx = __reduce_max_sync (0xffffffffu, mx);
  mn = __reduce_min_sync (0xffffffffu, mn);
  if (__any_sync (0xffffffffu, dust)) bad = true;
  if (mx < 0x00100000) return W_SENT;     // an empty cell
  const int emx = mx >> 20, emn = mn >> 20;
  if (emx == 0x7ff || emx - emn > W_SPREAD) bad = true;
  return F + emx - 1023;
}

struct EnvD {
  const int64_t* start; const int64_t* end;
  __device__ __forceinline__ bool contains (int64_t i, int64_t o) const { return !start || (i >= start[o] && i < end[o]); }
};

template<int OP>
__global__ void __launch_bounds__(544) wide_kernel (const __grid_constant__ WParams p) {
  extern __shared__ __align__(16) char smem[];
  const WideTables& t = p.t;
  const int S = t.S, W = p.W, R = p.R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nWarps = blockDim.x >> 5;
  // ---- shared memory: [tables] [ring (nWarps x R cells of S doubles)] [FG] [bp stage] [work slot]
  size_t at = 0;
  const char* tab = p.blob;
  if (p.tabInSmem) {
    for (uint32_t n = threadIdx.x * 16; n < t.bytes; n += blockDim.x * 16) *reinterpret_cast<uint4*> (smem + n) = *reinterpret_cast<const uint4*> (p.blob + n);
    tab = smem;
    at = (t.bytes + 15) & ~(size_t) 15;
  }
  double* ring = reinterpret_cast<double*> (smem + at); at += (size_t) nWarps * R * S * 8;
  int2* fg = reinterpret_cast<int2*> (smem + at); at += (size_t) nWarps * R * 8;
  uint16_t* bpStage = reinterpret_cast<uint16_t*> (smem + at); at += OP == OP_MAX ? (size_t) nWarps * S * 2 : 0;
  at = (at + 7) & ~(size_t) 7;
  volatile long long* workSlot = reinterpret_cast<volatile long long*> (smem + at);
  const WTab T = { reinterpret_cast<const double*> (tab + t.oW), reinterpret_cast<const uint16_t*> (tab + t.oSrc),
                   reinterpret_cast<const uint32_t*> (tab + t.oJobOffM), reinterpret_cast<const uint32_t*> (tab + t.oJobOffD),
                   reinterpret_cast<const uint32_t*> (tab + t.oJobOffI), reinterpret_cast<const WJob*> (tab + t.oJobs),
                   reinterpret_cast<const WJob*> (tab + t.oSil), reinterpret_cast<const unsigned char*> (tab + t.oSilCoop),
                   reinterpret_cast<const uint16_t*> (tab + t.oLive) };
  __syncthreads();
  double* myRing = ring + (size_t) warp * R * S;
  int2* myFG = fg + warp * R;
  uint16_t* bpS = bpStage + (OP == OP_MAX ? (size_t) warp * S : 0);
  const double LN2 = 0.693147180559945309417232121458;

  if (p.oneD) {
    // ---- generator / recogniser-free batches (no input): every warp sweeps its own pair, no CTA barriers
    for (;;) {
      long long wk = 0;
      if (lane == 0) wk = (long long) atomicAdd (p.counter, 1ULL);
      wk = __shfl_sync (0xffffffffu, wk, 0);
      if (wk >= p.nWork) break;
      const int64_t k = p.order[wk];
      const uint8_t* y = p.b.y + p.b.yOff[k];
      const int64_t Lo = p.b.yOff[k + 1] - p.b.yOff[k];
      EnvD env { nullptr, nullptr };
      if (p.b.envOff && p.b.envOff[k + 1] != p.b.envOff[k]) env = EnvD { p.b.envStart + p.b.envOff[k], p.b.envEnd + p.b.envOff[k] };
      unsigned char* bp = OP == OP_MAX && p.bp ? p.bp + p.bpOff[wk] : nullptr;
      bool bad = false;
      int slot = 0;
      int Fprev = 0, Gprev = W_SENT;
      double res = OP == OP_SUM ? 0. : w_ninf();
      int Fres = 0;
      for (int64_t o = 0; o <= Lo; ++o) {
        double* cur = myRing + (size_t) slot * S;
        const double* up = o ? myRing + (size_t) (slot ^ 1) * S : nullptr;
        const bool inside = env.contains (0, o);
        int F = 0;
        double fU = 1.;
        if (OP == OP_SUM && up) { if (Gprev == W_SENT) up = nullptr; else { F = Gprev; fU = w_pow2 (Fprev - F); } }
        if (inside) compute_cell<OP> (T, t, cur, bpS, up, nullptr, nullptr, fU, 0., 0., 0, o ? y[o - 1] : 0, o == 0, lane);
        else { for (int d = lane; d < S; d += 32) cur[d] = OP == OP_SUM ? 0. : w_ninf(); __syncwarp(); }
        if (OP == OP_SUM) { Fprev = F; Gprev = inside ? cell_frame (cur, S, F, lane, bad) : W_SENT; }
        else if (bp && inside) {
          unsigned char* row = bp + (size_t) o * S * t.bpBytes;
          if (t.bpBytes == 1) for (int d = lane; d < S; d += 32) row[d] = (unsigned char) bpS[d];
          else for (int d = lane; d < S; d += 32) reinterpret_cast<uint16_t*> (row)[d] = bpS[d];
        }
        if (o == Lo) { res = cur[S - 1]; Fres = F; }
        __syncwarp();
        slot ^= 1;
      }
      if (lane == 0) {
        if (OP == OP_SUM) { p.result[k] = res > 0. ? log (res) + Fres * LN2 : w_ninf(); p.flag[k] = bad || !(res > 0.) || !(res < 1e300); }
        else p.result[k] = res;
      }
    }
    return;
  }

  // ---- two-dimensional sweep: warp 0 streams the left boundary in, warps 1..W own the strip's columns
  double* bndBase = p.bnd + (size_t) blockIdx.x * 2 * p.bndRows * t.nLiveIn;
  int2* bndFGBase = p.bndFG + (size_t) blockIdx.x * 2 * p.bndRows;
  for (;;) {
    if (threadIdx.x == 0) *workSlot = (long long) atomicAdd (p.counter, 1ULL);
    __syncthreads();
    const long long wk = *workSlot;
    __syncthreads();
    if (wk >= p.nWork) break;
    const int64_t k = p.order[wk];
    const uint8_t* x = p.b.x + p.b.xOff[k];
    const uint8_t* y = p.b.y + p.b.yOff[k];
    const int64_t Li = p.b.xOff[k + 1] - p.b.xOff[k], Lo = p.b.yOff[k + 1] - p.b.yOff[k];
    EnvD env { nullptr, nullptr };
    if (p.b.envOff && p.b.envOff[k + 1] != p.b.envOff[k]) env = EnvD { p.b.envStart + p.b.envOff[k], p.b.envEnd + p.b.envOff[k] };
    unsigned char* bp = OP == OP_MAX && p.bp ? p.bp + p.bpOff[wk] : nullptr;
    bool bad = false;
    const int64_t nStrips = (Li + W) / W;     // ceil ((Li + 1) / W)
    for (int64_t strip = 0; strip < nStrips; ++strip) {
      const int64_t i0 = strip * W;
      const int nCols = (int) min ((int64_t) W, Li + 1 - i0);
      const int col = warp - 1;
      const bool active = warp >= 1 && col < nCols;
      const int64_t i = i0 + col;
      const int a = (active && i > 0) ? x[i - 1] : 0;
      const bool writesBnd = active && col == nCols - 1 && strip + 1 < nStrips;
      const double* bndIn = bndBase + (size_t) ((strip & 1) ^ 1) * p.bndRows * t.nLiveIn;
      double* bndOut = bndBase + (size_t) (strip & 1) * p.bndRows * t.nLiveIn;
      const int2* fgIn = bndFGBase + (size_t) ((strip & 1) ^ 1) * p.bndRows;
      int2* fgOut = bndFGBase + (size_t) (strip & 1) * p.bndRows;
      if (warp == 0 && strip > 0) {      // boundary row 0 -> the slot step 0 reads
        double* dst = myRing + (size_t) (R - 1) * S;
        for (int q = lane; q < t.nLiveIn; q += 32) dst[T.live[q]] = bndIn[q];
        if (lane == 0 && OP == OP_SUM) myFG[R - 1] = fgIn[0];
      }
      __syncthreads();
      const int64_t nSteps = Lo + nCols;
      int slot = 0;      // t % R
      for (int64_t ts = 0; ts < nSteps; ++ts) {
        const int prev = slot == 0 ? R - 1 : slot - 1;            // (t-1) % R
        const int prev2 = prev == 0 ? R - 1 : prev - 1;           // (t-2) % R (only read when R == 3)
        if (warp == 0) {
          if (strip > 0 && ts + 1 <= Lo) {      // the virtual column left of the strip "computes" row t+1 at step t
            double* dst = myRing + (size_t) slot * S;
            const double* srcRow = bndIn + (size_t) (ts + 1) * t.nLiveIn;
            for (int q = lane; q < t.nLiveIn; q += 32) dst[T.live[q]] = srcRow[q];
            if (lane == 0 && OP == OP_SUM) myFG[slot] = fgIn[ts + 1];
          }
        } else {
          const int64_t o = ts - col;
          if (active && o >= 0 && o <= Lo) {
            double* cur = myRing + (size_t) slot * S;
            const double* up = o > 0 ? myRing + (size_t) prev * S : nullptr;
            const double* left = i > 0 ? myRing - (size_t) R * S + (size_t) prev * S : nullptr;
            const double* diag = (i > 0 && o > 0 && t.hasMatch) ? myRing - (size_t) R * S + (size_t) prev2 * S : nullptr;
            const bool inside = env.contains (i, o);
            int F = 0;
            double fU = 1., fL = 1., fD = 1.;
            if (OP == OP_SUM) {
              const int2 gU = up ? myFG[prev] : make_int2 (0, W_SENT);
              const int2 gL = left ? myFG[prev - R] : make_int2 (0, W_SENT);
              const int2 gD = diag ? myFG[prev2 - R] : make_int2 (0, W_SENT);
              const int G = max (gU.y, max (gL.y, gD.y));
              F = G == W_SENT ? 0 : G;
              if (gU.y == W_SENT) up = nullptr; else { fU = w_pow2 (gU.x - F); if (F - gU.y > W_SPREAD) bad = true; }
              if (gL.y == W_SENT) left = nullptr; else { fL = w_pow2 (gL.x - F); if (F - gL.y > W_SPREAD) bad = true; }
              if (gD.y == W_SENT) diag = nullptr; else { fD = w_pow2 (gD.x - F); if (F - gD.y > W_SPREAD) bad = true; }
            }
            if (inside) compute_cell<OP> (T, t, cur, bpS, up, left, diag, fU, fL, fD, a, o ? y[o - 1] : 0, i == 0 && o == 0, lane);
            else { for (int d = lane; d < S; d += 32) cur[d] = OP == OP_SUM ? 0. : w_ninf(); __syncwarp(); }
            if (OP == OP_SUM) {
              const int G = inside ? cell_frame (cur, S, F, lane, bad) : W_SENT;
              if (lane == 0) myFG[slot] = make_int2 (F, G);
              if (writesBnd && lane == 0) fgOut[o] = make_int2 (F, G);
            } else if (bp && inside) {
              unsigned char* row = bp + ((size_t) o * (Li + 1) + i) * S * t.bpBytes;
              if (t.bpBytes == 1) for (int d = lane; d < S; d += 32) row[d] = (unsigned char) bpS[d];
              else for (int d = lane; d < S; d += 32) reinterpret_cast<uint16_t*> (row)[d] = bpS[d];
            }
            if (writesBnd) { double* dstRow = bndOut + (size_t) o * t.nLiveIn; for (int q = lane; q < t.nLiveIn; q += 32) dstRow[q] = cur[T.live[q]]; }
            if (i == Li && o == Lo && lane == 0) {
              const double res = cur[S - 1];
              if (OP == OP_SUM) { p.result[k] = res > 0. ? log (res) + F * LN2 : w_ninf(); if (!(res > 0.) || !(res < 1e300)) bad = true; }
              else p.result[k] = res;
            }
          }
        }
        __syncthreads();
        slot = slot + 1 == R ? 0 : slot + 1;
      }
    }
    if (OP == OP_SUM && bad && lane == 0) p.flag[k] = 1;      // flag[] is cleared by the host before the launch
  }
}

// DPMatrix::traceBack (dpmatrix.defs.h:82-110) over the stored back-pointers, one thread per pair.
// lenOut only (out == nullptr) or the path written start -> end.
__global__ void wide_traceback_kernel (DevMachine m, DevBatch b, const int64_t* __restrict__ order, int64_t nWork,
                                       const unsigned char* __restrict__ bp, const int64_t* __restrict__ bpOff, int bpBytes,
                                       const double* __restrict__ score, int64_t* __restrict__ lenOut, int32_t* __restrict__ out,
                                       const int64_t* __restrict__ outOff) {
  const int64_t wk = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (wk >= nWork) return;
  const int64_t k = order[wk];
  const uint8_t* x = b.x + b.xOff[k];
  const uint8_t* y = b.y + b.yOff[k];
  const int64_t Li = b.xOff[k + 1] - b.xOff[k], Lo = b.yOff[k + 1] - b.yOff[k];
  const int S = m.S;
  const unsigned char* base = bp + bpOff[wk];
  const unsigned kb = bpBytes == 1 ? 6 : 14, none = bpBytes == 1 ? 0xffu : 0xffffu;
  int64_t n = 0;
  if (score[k] > __longlong_as_double (0xfff0000000000000LL)) {      // boss.cpp:831
    const int64_t total = out ? lenOut[wk] : 0;
    int64_t i = Li, o = Lo;
    int s = S - 1;
    while (i > 0 || o > 0 || s != 0) {
      const size_t cell = ((size_t) o * (Li + 1) + i) * S + s;
      const unsigned v = bpBytes == 1 ? base[cell] : reinterpret_cast<const uint16_t*> (base)[cell];
      if (v == none) break;      // cannot happen on a finite path
      const unsigned kind = v >> kb, idx = v & ((1u << kb) - 1);
      const int a = i ? x[i - 1] : 0, c = o ? y[o - 1] : 0;
      const int64_t ks = (int64_t) s * m.nIn1;
      const int64_t key = kind == T_MATCH ? (ks + a) * m.nOut1 + c : kind == T_DELETE ? (ks + a) * m.nOut1 : kind == T_INSERT ? ks * m.nOut1 + c : ks * m.nOut1;
      const int64_t q = m.inc.off[key] + idx;
      if (out) out[outOff[wk] + total - 1 - n] = m.inc.id[q];
      ++n;
      if (kind == T_MATCH || kind == T_DELETE) --i;
      if (kind == T_MATCH || kind == T_INSERT) --o;
      s = m.inc.other[q];
    }
  }
  if (!out) lenOut[wk] = n;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct WHost {
  WideTables t {};
  std::vector<char> blobLin, blobLog;      // identical except for the weights: exp(lw) and lw
  std::vector<int64_t> entPerm;            // table entry n is hInc entry entPerm[n]
  char* dLin = nullptr;
  char* dLog = nullptr;
  bool linearOk = false;                   // every finite log-weight within +-30 ln 2
  int maxList = 0;
  int numSMs = 148;
};

static WHost* wh (const mb_machine* m) { return static_cast<WHost*> (m->wide); }

template<class T> static uint32_t put_vec (std::vector<char>& blob, const std::vector<T>& v) {
  const size_t at = (blob.size() + 15) & ~(size_t) 15;
  blob.resize (at + std::max<size_t> (v.size(), 1) * sizeof (T), 0);
  if (!v.empty()) memcpy (blob.data() + at, v.data(), v.size() * sizeof (T));
  return (uint32_t) at;
}

bool wide_supported (const mb_machine* m, std::string* why) {
  if (m->S > 60000) { if (why) *why = "more than 60000 states"; return false; }
  if (m->T > 100000000) { if (why) *why = "too many transitions"; return false; }
  // one column (a ring of 3 cells + the back-pointer stage) must fit in shared memory
  if ((size_t) m->S * (3 * 8 + 2) + 1024 > 227 * 1024) { if (why) *why = "a cell does not fit in shared memory"; return false; }
  return true;
}

static void wide_fill_weights (const mb_machine* m, WHost* h) {
  double* lin = reinterpret_cast<double*> (h->blobLin.data() + h->t.oW);
  double* lg = reinterpret_cast<double*> (h->blobLog.data() + h->t.oW);
  bool ok = true;
  const double lim = 30. * 0.6931471805599453;
  for (size_t n = 0; n < h->entPerm.size(); ++n) {
    const double lw = m->hInc.lw[h->entPerm[n]];
    lg[n] = lw;
    lin[n] = exp (lw);
    if (std::isnan (lw) || (std::isfinite (lw) && fabs (lw) > lim) || lw == INFINITY) ok = false;
  }
  h->linearOk = ok;
}

void wide_destroy (mb_machine* m) {
  WHost* h = wh (m);
  if (!h) return;
  if (h->dLin) cudaFree (h->dLin);
  if (h->dLog) cudaFree (h->dLog);
  delete h;
  m->wide = nullptr;
}

int wide_update_weights (mb_machine* m) {
  WHost* h = wh (m);
  if (!h) return 0;
  wide_fill_weights (m, h);
  const size_t wBytes = std::max<size_t> (h->entPerm.size(), 1) * 8;
  MB_CUDA (cudaMemcpy (h->dLin + h->t.oW, h->blobLin.data() + h->t.oW, wBytes, cudaMemcpyHostToDevice));
  MB_CUDA (cudaMemcpy (h->dLog + h->t.oW, h->blobLog.data() + h->t.oW, wBytes, cudaMemcpyHostToDevice));
  return 0;
}

int wide_prepare (mb_machine* m) {
  WHost* h = new WHost;
  m->wide = h;
  const int S = m->S, nIn = m->nIn, nOut = m->nOut, nIn1 = nIn + 1, nOut1 = nOut + 1;
  const HostCsr& inc = m->hInc;
  std::vector<uint16_t> entSrc;
  std::vector<WJob> jobs, sil;
  std::vector<uint32_t> jobOffM, jobOffD, jobOffI;
  std::vector<unsigned char> silCoop;
  std::vector<char> isLive ((size_t) S, 0);
  int maxList = 0;
  auto add_list = [&] (int64_t key, int dst, std::vector<WJob>& to) {
    const int64_t p0 = inc.off[key], p1 = inc.off[key + 1];
    if (p0 == p1) return;
    WJob j; j.dst = (uint16_t) dst; j.len = (uint16_t) std::min<int64_t> (p1 - p0, 65535); j.start = (uint32_t) entSrc.size();
    for (int64_t p = p0; p < p1; ++p) { entSrc.push_back ((uint16_t) inc.other[p]); h->entPerm.push_back (p); }
    maxList = std::max<int> (maxList, (int) (p1 - p0));
    to.push_back (j);
  };
  auto by_len = [] (const WJob& a, const WJob& b) { return a.len > b.len; };
  // token-consuming lists, per token context, longest lists first so that a round of 32 lanes is even
  bool hasMatch = false;
  for (int a = 1; a <= nIn; ++a) for (int c = 1; c <= nOut; ++c) {
    jobOffM.push_back ((uint32_t) jobs.size());
    const size_t j0 = jobs.size();
    for (int d = 0; d < S; ++d) add_list (((int64_t) d * nIn1 + a) * nOut1 + c, d, jobs);
    std::stable_sort (jobs.begin() + j0, jobs.end(), by_len);
    if (jobs.size() > j0) hasMatch = true;
  }
  jobOffM.push_back ((uint32_t) jobs.size());
  for (int a = 1; a <= nIn; ++a) {
    jobOffD.push_back ((uint32_t) jobs.size());
    const size_t j0 = jobs.size();
    for (int d = 0; d < S; ++d) add_list (((int64_t) d * nIn1 + a) * nOut1, d, jobs);
    std::stable_sort (jobs.begin() + j0, jobs.end(), by_len);
  }
  jobOffD.push_back ((uint32_t) jobs.size());
  const size_t nInputEntries = entSrc.size();
  for (size_t n = 0; n < nInputEntries; ++n) isLive[entSrc[n]] = 1;      // sources of match / delete transitions cross strip boundaries
  for (int c = 1; c <= nOut; ++c) {
    jobOffI.push_back ((uint32_t) jobs.size());
    const size_t j0 = jobs.size();
    for (int d = 0; d < S; ++d) add_list ((int64_t) d * nIn1 * nOut1 + c, d, jobs);
    std::stable_sort (jobs.begin() + j0, jobs.end(), by_len);
  }
  jobOffI.push_back ((uint32_t) jobs.size());
  // silent lists in dependency levels (level 0 has no silent sources); state 0's only possible silent
  // source is its own self-loop, which contributes nothing (it reads the cell being computed)
  const int nLevels = (int) m->fwdLevelOff.size() - 1;
  for (int l = 1; l < nLevels; ++l) {
    std::vector<WJob> lv;
    for (int n = m->fwdLevelOff[l]; n < m->fwdLevelOff[l + 1]; ++n) {
      const int d = m->fwdLevelStates[n];
      if (d == 0) continue;
      add_list ((int64_t) d * nIn1 * nOut1, d, lv);
    }
    std::stable_sort (lv.begin(), lv.end(), by_len);
    size_t n = 0;
    for (; n < lv.size() && lv[n].len >= W_COOP_LEN; ++n) {      // one cooperative round per long list
      const uint32_t chunk = (lv[n].len + 31u) / 32u;
      for (uint32_t lane = 0; lane < 32; ++lane) {
        const uint32_t b0 = std::min<uint32_t> (lane * chunk, lv[n].len), b1 = std::min<uint32_t> (b0 + chunk, lv[n].len);
        WJob j; j.dst = lv[n].dst; j.len = (uint16_t) (b1 - b0); j.start = lv[n].start + b0;
        sil.push_back (j);
      }
      silCoop.push_back (1);
    }
    for (; n < lv.size(); n += 32) {
      for (size_t q = n; q < n + 32; ++q) { WJob j; j.dst = 0; j.len = 0; j.start = 0; sil.push_back (q < lv.size() ? lv[q] : j); }
      silCoop.push_back (0);
    }
  }
  std::vector<uint16_t> live;
  for (int s = 0; s < S; ++s) if (isLive[s]) live.push_back ((uint16_t) s);
  if (maxList > 16383) { set_error ("wide engine: a transition list has more than 16383 entries"); return 1; }
  WideTables& t = h->t;
  std::vector<char>& blob = h->blobLin;
  std::vector<double> wZero (h->entPerm.size(), 0.);
  t.oW = put_vec (blob, wZero); t.oSrc = put_vec (blob, entSrc);
  t.oJobOffM = put_vec (blob, jobOffM); t.oJobOffD = put_vec (blob, jobOffD); t.oJobOffI = put_vec (blob, jobOffI);
  t.oJobs = put_vec (blob, jobs); t.oSil = put_vec (blob, sil); t.oSilCoop = put_vec (blob, silCoop); t.oLive = put_vec (blob, live);
  blob.resize ((blob.size() + 15) & ~(size_t) 15, 0);
  t.bytes = (uint32_t) blob.size();
  t.S = S; t.nIn = nIn; t.nOut = nOut; t.nSilRounds = (int32_t) silCoop.size(); t.hasMatch = hasMatch ? 1 : 0;
  t.nLiveIn = (int32_t) live.size(); t.bpBytes = maxList <= 63 ? 1 : 2;
  h->maxList = maxList;
  h->blobLog = h->blobLin;
  wide_fill_weights (m, h);
  MB_CUDA (cudaSetDevice (m->device));
  MB_CUDA (cudaMalloc (&h->dLin, t.bytes));
  MB_CUDA (cudaMalloc (&h->dLog, t.bytes));
  MB_CUDA (cudaMemcpy (h->dLin, h->blobLin.data(), t.bytes, cudaMemcpyHostToDevice));
  MB_CUDA (cudaMemcpy (h->dLog, h->blobLog.data(), t.bytes, cudaMemcpyHostToDevice));
  MB_CUDA (cudaDeviceGetAttribute (&h->numSMs, cudaDevAttrMultiProcessorCount, m->device));
  return 0;
}

// launch shape: column warps per CTA (W), ring depth, where the tables live, CTAs per SM
struct WShape { int W = 0, R = 2, tabInSmem = 0, ctasPerSM = 0; size_t smem = 0; bool oneD = false; };

template<int OP>
static int choose_shape (const WHost* h, bool oneD, WShape& best) {
  const WideTables& t = h->t;
  const int R = t.hasMatch ? 3 : 2;
  const size_t kMaxSmem = 227 * 1024;
  MB_CUDA (cudaFuncSetAttribute (wide_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxSmem));
  double bestScore = -1;
  static const int cand[] = { 16, 12, 8, 6, 4, 3, 2, 1 };
  for (int tabIn = 1; tabIn >= 0; --tabIn)
    for (int W: cand) {
      const int nWarps = W + 1;
      size_t smem = (tabIn ? ((size_t) t.bytes + 15) & ~(size_t) 15 : 0) + (size_t) nWarps * R * t.S * 8 + (size_t) nWarps * R * 8
        + (OP == OP_MAX ? (size_t) nWarps * t.S * 2 : 0) + 32;
      if (smem > kMaxSmem) continue;
      int ctas = 0;
      MB_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&ctas, wide_kernel<OP>, nWarps * 32, smem));
      if (ctas < 1) continue;
      const double score = (double) (oneD ? nWarps : W) * ctas * (tabIn ? 1.0 : 0.6);
      if (score > bestScore) { bestScore = score; best.W = W; best.R = R; best.tabInSmem = tabIn; best.ctasPerSM = ctas; best.smem = smem; best.oneD = oneD; }
    }
  if (bestScore < 0) { set_error ("wide engine: no launch shape fits in shared memory"); return 1; }
  return 0;
}

static int ensure_paths (mb_batch* b, int64_t need) {
  if (need <= b->pathsCapacity) return 0;
  const int64_t cap = std::max<int64_t> (need, 2 * b->pathsCapacity);
  int32_t* p = nullptr;
  MB_CUDA (cudaMalloc (&p, (size_t) cap * 4));
  if (b->dPaths) {
    MB_CUDA (cudaMemcpyAsync (p, b->dPaths, (size_t) b->pathsCapacity * 4, cudaMemcpyDeviceToDevice, b->stream));
    MB_CUDA (cudaStreamSynchronize (b->stream));
    cudaFree (b->dPaths);
  }
  b->dPaths = p;
  b->pathsCapacity = cap;
  return 0;
}

struct WBuf {
  void* p = nullptr;
  ~WBuf() { if (p) cudaFree (p); }
  int alloc (size_t bytes) { MB_CUDA (cudaMalloc (&p, bytes ? bytes : 8)); return 0; }
  template<class T> T* as() { return (T*) p; }
};

static std::vector<int64_t> cost_order (const mb_batch* b, const std::vector<int64_t>* subset) {
  std::vector<int64_t> order;
  if (subset) order = *subset;
  else { order.resize ((size_t) b->nPairs); for (int64_t k = 0; k < b->nPairs; ++k) order[k] = k; }
  auto cost = [&] (int64_t k) { return (double) (b->xOff[k + 1] - b->xOff[k] + 1) * (double) (b->yOff[k + 1] - b->yOff[k] + 1); };
  std::stable_sort (order.begin(), order.end(), [&] (int64_t a, int64_t c) { return cost (a) > cost (c); });
  return order;
}

// one launch of wide_kernel<OP> over `order`; bp / bpOff only for OP_MAX with traceback
template<int OP>
static int wide_launch (mb_machine* m, mb_batch* b, const std::vector<int64_t>& order, double* dResult, int32_t* dFlag,
                        unsigned char* dBp, const int64_t* dBpOff) {
  WHost* h = wh (m);
  bool oneD = true;
  int64_t maxLo = 0;
  for (int64_t k: order) { if (b->xOff[k + 1] != b->xOff[k]) oneD = false; maxLo = std::max (maxLo, b->yOff[k + 1] - b->yOff[k]); }
  WShape sh;
  if (choose_shape<OP> (h, oneD, sh)) return 1;
  const int nWarps = sh.W + 1;
  const int64_t nWork = (int64_t) order.size();
  const int64_t wantCtas = oneD ? (nWork + nWarps - 1) / nWarps : nWork;
  const int grid = (int) std::max<int64_t> (1, std::min<int64_t> (wantCtas, (int64_t) sh.ctasPerSM * h->numSMs));
  WBuf dOrder, dCounter, dBnd, dBndFG;
  if (dOrder.alloc (order.size() * 8) || dCounter.alloc (8)) return 1;
  MB_CUDA (cudaMemcpyAsync (dOrder.p, order.data(), order.size() * 8, cudaMemcpyHostToDevice, b->stream));
  MB_CUDA (cudaMemsetAsync (dCounter.p, 0, 8, b->stream));
  const int64_t bndRows = maxLo + 1;
  if (!oneD) {
    if (dBnd.alloc ((size_t) grid * 2 * bndRows * std::max (h->t.nLiveIn, 1) * 8) || dBndFG.alloc ((size_t) grid * 2 * bndRows * 8)) return 1;
  }
  WParams p {};
  p.t = h->t;
  p.blob = OP == OP_SUM ? h->dLin : h->dLog;
  p.tabInSmem = sh.tabInSmem; p.W = sh.W; p.R = sh.R; p.oneD = oneD ? 1 : 0;
  p.b = b->dev;
  p.order = dOrder.as<int64_t>(); p.nWork = nWork; p.counter = dCounter.as<unsigned long long>();
  p.result = dResult; p.flag = dFlag;
  p.bnd = dBnd.as<double>(); p.bndFG = dBndFG.as<int2>(); p.bndRows = bndRows;
  p.bp = dBp; p.bpOff = dBpOff;
  wide_kernel<OP><<<grid, nWarps * 32, sh.smem, b->stream>>> (p);
  MB_CUDA (cudaGetLastError());
  MB_CUDA (cudaStreamSynchronize (b->stream));      // the scratch buffers above die with this scope
  return 0;
}

int wide_forward (mb_machine* m, mb_batch* b, double* loglike) {
  b->lastRedo = 0;
  if (b->nPairs == 0) return 0;
  WHost* h = wh (m);
  const std::vector<int64_t> order = cost_order (b, nullptr);
  WBuf dRes, dFlag;
  if (dRes.alloc ((size_t) b->nPairs * 8) || dFlag.alloc ((size_t) b->nPairs * 4)) return 1;
  MB_CUDA (cudaMemsetAsync (dFlag.p, 0, (size_t) b->nPairs * 4, b->stream));
  if (timing_begin (b)) return 1;
  int64_t launches = 1;
  if (h->linearOk) { if (wide_launch<OP_SUM> (m, b, order, dRes.as<double>(), dFlag.as<int32_t>(), nullptr, nullptr)) return 1; }
  else { if (wide_launch<OP_LSE> (m, b, order, dRes.as<double>(), dFlag.as<int32_t>(), nullptr, nullptr)) return 1; }
  if (h->linearOk) {
    // pairs whose dynamic range the scaled sweep could not hold (or that came out as -inf) go through the log-domain sweep
    std::vector<int32_t> flag ((size_t) b->nPairs);
    MB_CUDA (cudaMemcpy (flag.data(), dFlag.p, flag.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> redo;
    for (int64_t k = 0; k < b->nPairs; ++k) if (flag[k]) redo.push_back (k);
    if (!redo.empty()) {
      if (wide_launch<OP_LSE> (m, b, cost_order (b, &redo), dRes.as<double>(), dFlag.as<int32_t>(), nullptr, nullptr)) return 1;
      ++launches;
    }
    b->lastRedo = (int64_t) redo.size();
  }
  if (timing_end (b, launches)) return 1;
  MB_CUDA (cudaMemcpy (loglike, dRes.p, (size_t) b->nPairs * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int wide_viterbi (mb_machine* m, mb_batch* b, double* score, int64_t* pathLen) {
  b->pathStart.clear();
  b->pathLen.clear();
  if (b->nPairs == 0) return 0;
  WHost* h = wh (m);
  const bool trace = pathLen != nullptr;
  const std::vector<int64_t> order = cost_order (b, nullptr);
  WBuf dRes;
  if (dRes.alloc ((size_t) b->nPairs * 8)) return 1;
  if (!trace) {
    if (timing_begin (b)) return 1;
    if (wide_launch<OP_MAX> (m, b, order, dRes.as<double>(), nullptr, nullptr, nullptr)) return 1;
    if (timing_end (b, 1)) return 1;
    MB_CUDA (cudaMemcpy (score, dRes.p, (size_t) b->nPairs * 8, cudaMemcpyDeviceToHost));
    return 0;
  }
  // chunks of pairs whose back-pointers ((Li+1)(Lo+1) S bytes or half-words) fit in free memory
  size_t freeB = 0, totalB = 0;
  MB_CUDA (cudaMemGetInfo (&freeB, &totalB));
  const double budget = 0.75 * (double) freeB;
  b->pathStart.assign ((size_t) b->nPairs, 0);
  b->pathLen.assign ((size_t) b->nPairs, 0);
  int64_t packed = 0, launches = 0;
  double ms = 0;
  for (size_t c0 = 0; c0 < order.size();) {
    std::vector<int64_t> chunk, bpOff;
    double bytes = 0;
    size_t c1 = c0;
    for (; c1 < order.size(); ++c1) {
      const int64_t k = order[c1];
      const double need = (double) (b->xOff[k + 1] - b->xOff[k] + 1) * (double) (b->yOff[k + 1] - b->yOff[k] + 1) * m->S * h->t.bpBytes;
      if (need > budget) { set_error ("pair " + std::to_string (k) + " needs more device memory for its back-pointers than is free (" + std::to_string (need) + " bytes)"); return 1; }
      if (!chunk.empty() && bytes + need > budget) break;
      chunk.push_back (k);
      bpOff.push_back ((int64_t) bytes);
      bytes += (need + 15) - fmod (need + 15, 16);
    }
    c0 = c1;
    WBuf dBp, dBpOff, dOrder, dLen, dOutOff;
    if (dBp.alloc ((size_t) bytes) || dBpOff.alloc (bpOff.size() * 8) || dOrder.alloc (chunk.size() * 8) || dLen.alloc (chunk.size() * 8)) return 1;
    MB_CUDA (cudaMemcpyAsync (dBpOff.p, bpOff.data(), bpOff.size() * 8, cudaMemcpyHostToDevice, b->stream));
    MB_CUDA (cudaMemcpyAsync (dOrder.p, chunk.data(), chunk.size() * 8, cudaMemcpyHostToDevice, b->stream));
    if (timing_begin (b)) return 1;
    if (wide_launch<OP_MAX> (m, b, chunk, dRes.as<double>(), nullptr, dBp.as<unsigned char>(), dBpOff.as<int64_t>())) return 1;
    const int64_t nWork = (int64_t) chunk.size();
    const unsigned tg = (unsigned) ((nWork + 31) / 32);
    wide_traceback_kernel<<<tg, 32, 0, b->stream>>> (m->dev, b->dev, dOrder.as<int64_t>(), nWork, dBp.as<unsigned char>(), dBpOff.as<int64_t>(), h->t.bpBytes,
                                                      dRes.as<double>(), dLen.as<int64_t>(), nullptr, nullptr);
    MB_CUDA (cudaGetLastError());
    std::vector<int64_t> len (chunk.size()), off (chunk.size());
    MB_CUDA (cudaMemcpyAsync (len.data(), dLen.p, len.size() * 8, cudaMemcpyDeviceToHost, b->stream));
    MB_CUDA (cudaStreamSynchronize (b->stream));
    for (size_t n = 0; n < len.size(); ++n) {
      off[n] = packed;
      b->pathStart[chunk[n]] = packed;
      b->pathLen[chunk[n]] = len[n];
      packed += len[n];
    }
    if (ensure_paths (b, packed)) return 1;
    if (dOutOff.alloc (off.size() * 8)) return 1;
    MB_CUDA (cudaMemcpyAsync (dOutOff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, b->stream));
    wide_traceback_kernel<<<tg, 32, 0, b->stream>>> (m->dev, b->dev, dOrder.as<int64_t>(), nWork, dBp.as<unsigned char>(), dBpOff.as<int64_t>(), h->t.bpBytes,
                                                      dRes.as<double>(), dLen.as<int64_t>(), b->dPaths, dOutOff.as<int64_t>());
    MB_CUDA (cudaGetLastError());
    launches += 3;
    if (timing_end (b, launches)) return 1;
    ms += b->lastMs;
  }
  b->lastMs = ms;
  b->lastLaunches = launches;
  MB_CUDA (cudaMemcpy (score, dRes.p, (size_t) b->nPairs * 8, cudaMemcpyDeviceToHost));
  // paths were packed in chunk order; mb_viterbi_paths copies them out by pathStart
  for (int64_t k = 0; k < b->nPairs; ++k) pathLen[k] = b->pathLen[k];
  return 0;
}

}  // namespace mb
