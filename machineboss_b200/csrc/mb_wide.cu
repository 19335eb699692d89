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
#define W_SPREAD 600             // exponent spread that hands a pair to the log-domain sweep
#define W_RSYNC 1                // round flags: group barrier after the round (end of a dependency level / phase)
#define W_RREDUCE 2              //              some list of the round is split over several lanes

// Transition tables ("rounds").  The lists of one dependency-free group (one silent level, or one
// token context of one kind) are dealt to the G lanes of a cell group; a round is G lists side by
// side, padded to the longest (weight 0 / -inf), so every lane runs the same number of rows; a long
// list is cut into 2^k contiguous pieces on adjacent lanes and combined with shuffles.
struct WEntry { double w; uint32_t srcOff; uint32_t pad; };              // srcOff = source state * 8
struct WRound { uint32_t rowStart; uint16_t nRows; uint16_t flags; };
struct WLane { uint32_t a; uint32_t b; };      // a: dst (16; 0xffff = nothing to write) | log2 pieces (3) | round flags (2);  b: idx0 (16) | rows of the round (16)

struct WideTables {       // byte offsets into one blob (entries first: the two blobs differ only in their weights)
  uint32_t oEnt, oRound, oLane, oPhM, oPhD, oPhI, oLive, bytes;
  uint32_t silR0, silR1;
  int32_t S, nIn, nOut, hasMatch, nLiveIn, bpBytes, G;
};

struct WParams {
  WideTables t;
  const char* blob;
  int32_t W, R, oneD;      // W: column warps per CTA (each 32/G columns)
  DevBatch b;
  const int64_t* order;
  int64_t nWork;
  unsigned long long* counter;
  double* result;
  int32_t* flag;
  double* bnd;          // per CTA: 2 buffers of bndRows * nLiveIn doubles
  int2* bndFG;          // per CTA: 2 buffers of bndRows (F, G)
  int64_t bndRows;
  unsigned char* bp;    // back-pointers, work item n at bpOff[n] (bytes), layout [o][i][s]
  const int64_t* bpOff;
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
  const WEntry* ent; const WRound* round; const WLane* lane; const uint32_t* phM; const uint32_t* phD; const uint32_t* phI; const uint16_t* live;
};

// rounds [r0, r1) of one phase, reading the cell `from` (a neighbour, or the cell itself for silent lists).
// Rows of consecutive rounds are consecutive in the entry table, so one pointer runs through the
// phase with the next entry always in flight; the tables end with a spare row and round.
template<int OP, int G>
__device__ __forceinline__ void run_rounds (const WTab& T, uint32_t r0, uint32_t r1, const double* from, double* cur, uint16_t* bpS,
                                            double f, unsigned kindBits, int gl, unsigned gmask) {
  if (r0 == r1) return;
  const WLane* lp = T.lane + (size_t) r0 * G + gl;
  uint2 li = *reinterpret_cast<const uint2*> (lp);
  const WEntry* e = T.ent + (size_t) T.round[r0].rowStart * G + gl;
  uint4 v = *reinterpret_cast<const uint4*> (e);
  for (uint32_t r = r0; r < r1; ++r) {
    lp += G;
    const uint2 nli = *reinterpret_cast<const uint2*> (lp);      // next round's lane word
    const unsigned dst = li.x & 0xffffu, segW = 1u << ((li.x >> 16) & 7u), flags = li.x >> 19, nRows = li.y >> 16, idx0 = li.y & 0xffffu;
    if (OP != OP_MAX) {
      const double c0 = dst != 0xffffu ? cur[dst] : 0.;
      double acc = OP == OP_SUM ? 0. : w_ninf();
#pragma unroll 2
      for (unsigned k = 0; k < nRows; ++k) {
        const uint4 u = v;
        e += G;
        v = *reinterpret_cast<const uint4*> (e);
        const double w = __hiloint2double ((int) u.y, (int) u.x);
        const double x = *reinterpret_cast<const double*> (reinterpret_cast<const char*> (from) + u.z);
        acc = OP == OP_SUM ? fma (w, x, acc) : w_lse (acc, x + w);
      }
      if (flags & W_RREDUCE)
        for (unsigned off = G / 2; off; off >>= 1) {
          const double o2 = __shfl_down_sync (gmask, acc, off, G);
          if (off < segW) acc = OP == OP_SUM ? acc + o2 : w_lse (acc, o2);
        }
      if (dst != 0xffffu) cur[dst] = OP == OP_SUM ? fma (f, acc, c0) : w_lse (c0, acc);
    } else {
      const double c0 = dst != 0xffffu ? cur[dst] : 0.;
      double best = w_ninf();
      unsigned bp = 0xffffu;
#pragma unroll 2
      for (unsigned k = 0; k < nRows; ++k) {
        const uint4 u = v;
        e += G;
        v = *reinterpret_cast<const uint4*> (e);
        const double c = *reinterpret_cast<const double*> (reinterpret_cast<const char*> (from) + u.z) + __hiloint2double ((int) u.y, (int) u.x);
        if (best < c) { best = c; bp = kindBits | (idx0 + k); }      // strict: the first maximum wins (dpmatrix.defs.h:171-174)
      }
      if (flags & W_RREDUCE)
        for (unsigned off = G / 2; off; off >>= 1) {      // the tree pairs lanes out of list order: a tie goes to the earlier candidate
          const double ob = __shfl_down_sync (gmask, best, off, G);
          const unsigned obp = __shfl_down_sync (gmask, bp, off, G);
          if (off < segW && (best < ob || (best == ob && obp < bp))) { best = ob; bp = obp; }
        }
      if (dst != 0xffffu && c0 < best) { cur[dst] = best; bpS[dst] = (uint16_t) bp; }      // earlier phases keep ties
    }
    if (flags & W_RSYNC) __syncwarp (gmask);
    li = nli;
  }
}

// One cell, computed by the G lanes of a group.  up/left/diag: neighbour cells (null when outside
// the matrix or empty).  Linear domain: fU/fL/fD scale the neighbour sums into this cell's frame.
template<int OP, int G>
__device__ __forceinline__ void compute_cell (const WTab& T, const WideTables& t, double* cur, uint16_t* bpS,
                                              const double* up, const double* left, const double* diag,
                                              double fU, double fL, double fD, int a, int c, bool origin, int gl, unsigned gmask) {
  const int S = t.S;
  for (int d = gl; d < S; d += G) { cur[d] = OP == OP_SUM ? 0. : w_ninf(); if (OP == OP_MAX) bpS[d] = 0xffff; }
  __syncwarp (gmask);
  if (origin && gl == 0) cur[0] = OP == OP_SUM ? 1. : 0.;
  __syncwarp (gmask);
  const unsigned kb = t.bpBytes == 1 ? 6 : 14;
  // reference candidate order: match, delete, insert, silent (dpmatrix.defs.h:90-103)
  if (diag) { const int k = (a - 1) * t.nOut + (c - 1); run_rounds<OP, G> (T, T.phM[k], T.phM[k + 1], diag, cur, bpS, fD, (unsigned) T_MATCH << kb, gl, gmask); }
  if (left) run_rounds<OP, G> (T, T.phD[a - 1], T.phD[a], left, cur, bpS, fL, (unsigned) T_DELETE << kb, gl, gmask);
  if (up) run_rounds<OP, G> (T, T.phI[c - 1], T.phI[c], up, cur, bpS, fU, (unsigned) T_INSERT << kb, gl, gmask);
  run_rounds<OP, G> (T, t.silR0, t.silR1, cur, cur, bpS, 1., (unsigned) T_SILENT << kb, gl, gmask);
}

// exponent bookkeeping of a finished linear-domain cell: returns G (W_SENT if empty); sets bad on
// overflow, underflow in progress, or an exponent spread beyond W_SPREAD
template<int G>
__device__ __forceinline__ int cell_frame (const double* cur, int S, int F, int gl, unsigned gmask, bool& bad) {
  int mx = 0, mn = 0x7fffffff;
  bool dust = false;
  for (int d = gl; d < S; d += G) {
    const double v = cur[d];
    const int hi = __double2hiint (v) & 0x7fffffff;
    mx = max (mx, hi);
    if (hi >= 0x00100000) mn = min (mn, hi);
    else if (v != 0.) dust = true;      // a denormal: something is underflowing
  }
  mx = __reduce_max_sync (gmask, mx);
  mn = __reduce_min_sync (gmask, mn);
  if (__any_sync (gmask, dust)) bad = true;
  if (mx < 0x00100000) return W_SENT;     // an empty cell
  const int emx = mx >> 20, emn = mn >> 20;
  if (emx == 0x7ff || emx - emn > W_SPREAD) bad = true;
  return F + emx - 1023;
}

struct EnvD {
  const int64_t* start; const int64_t* end;
  __device__ __forceinline__ bool contains (int64_t i, int64_t o) const { return !start || (i >= start[o] && i < end[o]); }
};

template<int OP, int G, bool TABS>
__global__ void __launch_bounds__(544) wide_kernel (const __grid_constant__ WParams p) {
  extern __shared__ __align__(16) char smem[];
  constexpr int CPW = 32 / G;      // cells (matrix columns, or pairs in the 1-D sweep) per warp
  const WideTables& t = p.t;
  const int S = t.S, R = p.R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nWarps = blockDim.x >> 5;
  const int grp = lane / G, gl = lane % G;
  const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (grp * G);
  // ---- shared memory: [tables] [ring: column slots x R cells of S doubles] [FG] [bp stage] [work slot]
  const int nSlots = nWarps * CPW;      // column slots; in the 2-D sweep slot CPW-1 of warp 0 is the virtual column left of the strip
  size_t at = 0;
  const char* tab = p.blob;
  if (TABS) {
    for (uint32_t n = threadIdx.x * 16; n < t.bytes; n += blockDim.x * 16) *reinterpret_cast<uint4*> (smem + n) = *reinterpret_cast<const uint4*> (p.blob + n);
    at = (t.bytes + 15) & ~(size_t) 15;
  }
  double* ring = reinterpret_cast<double*> (smem + at); at += (size_t) nSlots * R * S * 8;
  int2* fg = reinterpret_cast<int2*> (smem + at); at += (size_t) nSlots * R * 8;
  uint16_t* bpStage = reinterpret_cast<uint16_t*> (smem + at); at += OP == OP_MAX ? (size_t) nSlots * S * 2 : 0;
  at = (at + 7) & ~(size_t) 7;
  volatile long long* workSlot = reinterpret_cast<volatile long long*> (smem + at);
  WTab T;
  if (TABS) T = { reinterpret_cast<const WEntry*> (smem + t.oEnt), reinterpret_cast<const WRound*> (smem + t.oRound), reinterpret_cast<const WLane*> (smem + t.oLane),
                  reinterpret_cast<const uint32_t*> (smem + t.oPhM), reinterpret_cast<const uint32_t*> (smem + t.oPhD), reinterpret_cast<const uint32_t*> (smem + t.oPhI),
                  reinterpret_cast<const uint16_t*> (smem + t.oLive) };
  else T = { reinterpret_cast<const WEntry*> (tab + t.oEnt), reinterpret_cast<const WRound*> (tab + t.oRound), reinterpret_cast<const WLane*> (tab + t.oLane),
             reinterpret_cast<const uint32_t*> (tab + t.oPhM), reinterpret_cast<const uint32_t*> (tab + t.oPhD), reinterpret_cast<const uint32_t*> (tab + t.oPhI),
             reinterpret_cast<const uint16_t*> (tab + t.oLive) };
  __syncthreads();
  const int mySlot = warp * CPW + grp;
  double* myRing = ring + (size_t) mySlot * R * S;
  int2* myFG = fg + mySlot * R;
  uint16_t* bpS = bpStage + (OP == OP_MAX ? (size_t) mySlot * S : 0);
  const double LN2 = 0.693147180559945309417232121458;

  if (p.oneD) {
    // ---- batches without input sequences: every lane group sweeps its own pair, no CTA barriers
    for (;;) {
      long long wk = 0;
      if (gl == 0) wk = (long long) atomicAdd (p.counter, 1ULL);
      wk = __shfl_sync (gmask, wk, 0, G);
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
        if (inside) compute_cell<OP, G> (T, t, cur, bpS, up, nullptr, nullptr, fU, 0., 0., 0, o ? y[o - 1] : 0, o == 0, gl, gmask);
        else { for (int d = gl; d < S; d += G) cur[d] = OP == OP_SUM ? 0. : w_ninf(); __syncwarp (gmask); }
        if (OP == OP_SUM) { Fprev = F; Gprev = inside ? cell_frame<G> (cur, S, F, gl, gmask, bad) : W_SENT; }
        else if (bp && inside) {
          unsigned char* row = bp + (size_t) o * S * t.bpBytes;
          if (t.bpBytes == 1) for (int d = gl; d < S; d += G) row[d] = (unsigned char) bpS[d];
          else for (int d = gl; d < S; d += G) reinterpret_cast<uint16_t*> (row)[d] = bpS[d];
        }
        if (o == Lo) { res = cur[S - 1]; Fres = F; }
        __syncwarp (gmask);
        slot ^= 1;
      }
      if (gl == 0) {
        if (OP == OP_SUM) { p.result[k] = res > 0. ? log (res) + Fres * LN2 : w_ninf(); p.flag[k] = bad || !(res > 0.) || !(res < 1e300); }
        else p.result[k] = res;
      }
    }
    return;
  }

  // ---- two-dimensional sweep: warp 0 streams the left boundary in, warps 1.. own CPW columns each
  const int WC = (nWarps - 1) * CPW;      // columns per strip
  double* bndBase = p.bnd + (size_t) blockIdx.x * 2 * p.bndRows * t.nLiveIn;
  int2* bndFGBase = p.bndFG + (size_t) blockIdx.x * 2 * p.bndRows;
  double* virtRing = ring + (size_t) (CPW - 1) * R * S;      // the column slot just left of warp 1's first column
  int2* virtFG = fg + (CPW - 1) * R;
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
    const int64_t nStrips = (Li + WC) / WC;     // ceil ((Li + 1) / WC)
    for (int64_t strip = 0; strip < nStrips; ++strip) {
      const int64_t i0 = strip * WC;
      const int nCols = (int) min ((int64_t) WC, Li + 1 - i0);
      const int col = (warp - 1) * CPW + grp;
      const bool active = warp >= 1 && col < nCols;
      const int64_t i = i0 + col;
      const int a = (active && i > 0) ? x[i - 1] : 0;
      const bool writesBnd = active && col == nCols - 1 && strip + 1 < nStrips;
      const double* bndIn = bndBase + (size_t) ((strip & 1) ^ 1) * p.bndRows * t.nLiveIn;
      double* bndOut = bndBase + (size_t) (strip & 1) * p.bndRows * t.nLiveIn;
      const int2* fgIn = bndFGBase + (size_t) ((strip & 1) ^ 1) * p.bndRows;
      int2* fgOut = bndFGBase + (size_t) (strip & 1) * p.bndRows;
      if (warp == 0 && strip > 0) {      // boundary row 0 -> the slot step 0 reads
        double* dst = virtRing + (size_t) (R - 1) * S;
        for (int q = lane; q < t.nLiveIn; q += 32) dst[T.live[q]] = bndIn[q];
        if (lane == 0 && OP == OP_SUM) virtFG[R - 1] = fgIn[0];
      }
      __syncthreads();
      const int64_t nSteps = Lo + nCols;
      int slot = 0;      // t % R
      for (int64_t ts = 0; ts < nSteps; ++ts) {
        const int prev = slot == 0 ? R - 1 : slot - 1;            // (t-1) % R
        const int prev2 = prev == 0 ? R - 1 : prev - 1;           // (t-2) % R (only read when R == 3)
        if (warp == 0) {
          if (strip > 0 && ts + 1 <= Lo) {      // the virtual column left of the strip "computes" row t+1 at step t
            double* dst = virtRing + (size_t) slot * S;
            const double* srcRow = bndIn + (size_t) (ts + 1) * t.nLiveIn;
            for (int q = lane; q < t.nLiveIn; q += 32) dst[T.live[q]] = srcRow[q];
            if (lane == 0 && OP == OP_SUM) virtFG[slot] = fgIn[ts + 1];
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
              const int Gm = max (gU.y, max (gL.y, gD.y));
              F = Gm == W_SENT ? 0 : Gm;
              if (gU.y == W_SENT) up = nullptr; else { fU = w_pow2 (gU.x - F); if (F - gU.y > W_SPREAD) bad = true; }
              if (gL.y == W_SENT) left = nullptr; else { fL = w_pow2 (gL.x - F); if (F - gL.y > W_SPREAD) bad = true; }
              if (gD.y == W_SENT) diag = nullptr; else { fD = w_pow2 (gD.x - F); if (F - gD.y > W_SPREAD) bad = true; }
            }
            if (inside) compute_cell<OP, G> (T, t, cur, bpS, up, left, diag, fU, fL, fD, a, o ? y[o - 1] : 0, i == 0 && o == 0, gl, gmask);
            else { for (int d = gl; d < S; d += G) cur[d] = OP == OP_SUM ? 0. : w_ninf(); __syncwarp (gmask); }
            if (OP == OP_SUM) {
              const int Gc = inside ? cell_frame<G> (cur, S, F, gl, gmask, bad) : W_SENT;
              if (gl == 0) myFG[slot] = make_int2 (F, Gc);
              if (writesBnd && gl == 0) fgOut[o] = make_int2 (F, Gc);
            } else if (bp && inside) {
              unsigned char* row = bp + ((size_t) o * (Li + 1) + i) * S * t.bpBytes;
              if (t.bpBytes == 1) for (int d = gl; d < S; d += G) row[d] = (unsigned char) bpS[d];
              else for (int d = gl; d < S; d += G) reinterpret_cast<uint16_t*> (row)[d] = bpS[d];
            }
            if (writesBnd) { double* dstRow = bndOut + (size_t) o * t.nLiveIn; for (int q = gl; q < t.nLiveIn; q += G) dstRow[q] = cur[T.live[q]]; }
            if (i == Li && o == Lo && gl == 0) {
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
    if (OP == OP_SUM && bad && gl == 0) p.flag[k] = 1;      // flag[] is cleared by the host before the launch
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
  std::vector<int64_t> entPerm;            // table entry n is hInc entry entPerm[n] (-1: padding)
  char* dLin = nullptr;
  char* dLog = nullptr;
  bool linearOk = false;                   // every finite log-weight within +-30 ln 2
  int maxList = 0;
  int numSMs = 148;
  double estInstrPerCell = 0;
};

static WHost* wh (const mb_machine* m) { return static_cast<WHost*> (m->wide); }

template<class T> static uint32_t put_vec (std::vector<char>& blob, const std::vector<T>& v) {
  const size_t at = (blob.size() + 15) & ~(size_t) 15;
  blob.resize (at + std::max<size_t> (v.size(), 1) * sizeof (T), 0);
  if (!v.empty()) memcpy (blob.data() + at, v.data(), v.size() * sizeof (T));
  return (uint32_t) at;
}

bool wide_supported (const mb_machine* m, std::string* why) {
  if (m->S > 16000) { if (why) *why = "wide engine: more than 16000 states"; return false; }
  if (m->T > 100000000) { if (why) *why = "wide engine: too many transitions"; return false; }
  // one column (a ring of 3 cells + the back-pointer stage) must fit in shared memory
  if ((size_t) m->S * (3 * 8 + 2) + 1024 > 227 * 1024) { if (why) *why = "wide engine: a cell does not fit in shared memory"; return false; }
  return true;
}

// ---- table construction -------------------------------------------------------------------------
struct WList { int dst; int64_t p0, p1; };      // one destination's token-selected source list: hInc entries [p0, p1)

struct WBuilder {
  int G = 32;
  std::vector<WEntry> ent;
  std::vector<int64_t> perm;
  std::vector<WRound> rounds;
  std::vector<WLane> lanes;
  double cost = 0;      // estimated warp instructions spent in the rounds built so far
};

static int pow2ceil (int v) { int n = 1; while (n < v) n <<= 1; return n; }
static int ilog2 (int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

struct WItem { int list, n, rows; };

// lists of one dependency-free group -> rounds; piece cap C; returns the estimated cost, appends to B unless dry
static double pack_group (const std::vector<WList>& lists, const mb_machine* m, int C, WBuilder& B, bool dry) {
  const int G = B.G;
  std::vector<WItem> items;
  for (size_t j = 0; j < lists.size(); ++j) {
    const int L = (int) (lists[j].p1 - lists[j].p0);
    const int n = std::min (G, pow2ceil ((L + C - 1) / C));
    items.push_back (WItem { (int) j, n, (L + n - 1) / n });
  }
  std::stable_sort (items.begin(), items.end(), [] (const WItem& a, const WItem& b) { return a.rows != b.rows ? a.rows > b.rows : a.n > b.n; });
  double cost = 0;
  size_t q = 0;
  while (q < items.size()) {
    // one round: place items at aligned lane offsets until one does not fit
    std::vector<std::pair<int, WItem>> placed;
    int fill = 0, rows = 0;
    bool split = false;
    while (q < items.size()) {
      const int at = (fill + items[q].n - 1) / items[q].n * items[q].n;
      if (at + items[q].n > G) break;
      placed.push_back (std::make_pair (at, items[q]));
      fill = at + items[q].n;
      rows = std::max (rows, items[q].rows);
      split = split || items[q].n > 1;
      ++q;
    }
    cost += 18 + 5.5 * rows + (split ? 4.0 * ilog2 (G) : 0);
    if (dry) continue;
    WRound r; r.rowStart = (uint32_t) (B.ent.size() / G); r.nRows = (uint16_t) rows; r.flags = (uint16_t) ((split ? W_RREDUCE : 0) | (q == items.size() ? W_RSYNC : 0));
    B.rounds.push_back (r);
    const size_t l0 = B.lanes.size(), e0 = B.ent.size();
    WLane idle; idle.a = 0xffffu | ((uint32_t) r.flags << 19); idle.b = (uint32_t) rows << 16;
    B.lanes.resize (l0 + G, idle);
    WEntry padE; padE.w = 0; padE.srcOff = 0; padE.pad = 0;
    B.ent.resize (e0 + (size_t) rows * G, padE);
    B.perm.resize (e0 + (size_t) rows * G, -1);
    for (auto& pl: placed) {
      const WList& ls = lists[pl.second.list];
      const int L = (int) (ls.p1 - ls.p0);
      for (int j = 0; j < pl.second.n; ++j) {
        WLane& ln = B.lanes[l0 + pl.first + j];
        ln.a = (j == 0 ? (uint32_t) ls.dst : 0xffffu) | ((uint32_t) ilog2 (pl.second.n) << 16) | ((uint32_t) r.flags << 19);
        ln.b = (uint32_t) (j * pl.second.rows) | ((uint32_t) rows << 16);
        for (int k = 0; k < pl.second.rows; ++k) {
          const int n = j * pl.second.rows + k;
          if (n >= L) break;
          const size_t e = e0 + (size_t) k * G + pl.first + j;
          B.ent[e].srcOff = (uint32_t) m->hInc.other[ls.p0 + n] * 8u;
          B.perm[e] = ls.p0 + n;
        }
      }
    }
  }
  return cost;
}

static double add_group (const std::vector<WList>& lists, const mb_machine* m, WBuilder& B) {
  if (lists.empty()) return 0;
  int maxL = 1;
  for (auto& l: lists) maxL = std::max<int> (maxL, (int) (l.p1 - l.p0));
  std::vector<int> caps;
  for (int c = 1; c < maxL; c = c < 4 ? c + 1 : c + c / 2) caps.push_back (c);
  caps.push_back (maxL);
  int bestC = maxL;
  double best = 1e300;
  for (int c: caps) { const double v = pack_group (lists, m, c, B, true); if (v < best) { best = v; bestC = c; } }
  B.cost += pack_group (lists, m, bestC, B, false);
  return best;
}

// all tables for lane-group width G; the cost estimate is per cell (a warp works on 32/G cells at once)
static void build_tables (const mb_machine* m, int G, WBuilder& B, std::vector<uint32_t>& phM, std::vector<uint32_t>& phD, std::vector<uint32_t>& phI,
                          uint32_t& silR0, uint32_t& silR1, std::vector<uint16_t>& live, bool& hasMatch, int& maxList, double& perCell) {
  const int S = m->S, nIn = m->nIn, nOut = m->nOut, nIn1 = nIn + 1, nOut1 = nOut + 1;
  const HostCsr& inc = m->hInc;
  B = WBuilder();
  B.G = G;
  maxList = 0;
  std::vector<char> isLive ((size_t) S, 0);
  auto lists_for = [&] (int a, int c, int level, bool input) {
    std::vector<WList> ls;
    for (int d = 0; d < S; ++d) {
      if (level >= 0 && d == 0) continue;      // state 0's only possible silent source is its own self-loop, which contributes nothing
      const int64_t key = ((int64_t) d * nIn1 + a) * nOut1 + c;
      if (inc.off[key] == inc.off[key + 1]) continue;
      ls.push_back (WList { d, inc.off[key], inc.off[key + 1] });
      maxList = std::max<int> (maxList, (int) (inc.off[key + 1] - inc.off[key]));
      if (input) for (int64_t q = inc.off[key]; q < inc.off[key + 1]; ++q) isLive[inc.other[q]] = 1;
    }
    return ls;
  };
  double cM = 0, cD = 0, cI = 0, cS = 0;
  hasMatch = false;
  phM.clear(); phD.clear(); phI.clear();
  for (int a = 1; a <= nIn; ++a) for (int c = 1; c <= nOut; ++c) {
    phM.push_back ((uint32_t) B.rounds.size());
    const std::vector<WList> ls = lists_for (a, c, -1, true);
    if (!ls.empty()) hasMatch = true;
    cM += add_group (ls, m, B);
  }
  phM.push_back ((uint32_t) B.rounds.size());
  for (int a = 1; a <= nIn; ++a) { phD.push_back ((uint32_t) B.rounds.size()); cD += add_group (lists_for (a, 0, -1, true), m, B); }
  phD.push_back ((uint32_t) B.rounds.size());
  for (int c = 1; c <= nOut; ++c) { phI.push_back ((uint32_t) B.rounds.size()); cI += add_group (lists_for (0, c, -1, false), m, B); }
  phI.push_back ((uint32_t) B.rounds.size());
  silR0 = (uint32_t) B.rounds.size();
  const int nLevels = (int) m->fwdLevelOff.size() - 1;
  for (int l = 1; l < nLevels; ++l) {      // level 0 has no silent sources
    std::vector<WList> ls;
    for (int n = m->fwdLevelOff[l]; n < m->fwdLevelOff[l + 1]; ++n) {
      const int d = m->fwdLevelStates[n];
      if (d == 0) continue;
      const int64_t key = (int64_t) d * nIn1 * nOut1;
      if (inc.off[key] == inc.off[key + 1]) continue;
      ls.push_back (WList { d, inc.off[key], inc.off[key + 1] });
      maxList = std::max<int> (maxList, (int) (inc.off[key + 1] - inc.off[key]));
    }
    cS += add_group (ls, m, B);
  }
  silR1 = (uint32_t) B.rounds.size();
  live.clear();
  for (int s = 0; s < S; ++s) if (isLive[s]) live.push_back ((uint16_t) s);
  const double perWarp = cS + (nIn && nOut ? cM / ((double) nIn * nOut) : 0) + (nIn ? cD / nIn : 0) + (nOut ? cI / nOut : 0)
    + 9.0 * ((S + G - 1) / G) + 120;      // + zero pass, frame / back-pointer pass, step overhead
  perCell = perWarp / (32 / G);
}

static void wide_fill_weights (const mb_machine* m, WHost* h) {
  WEntry* lin = reinterpret_cast<WEntry*> (h->blobLin.data() + h->t.oEnt);
  WEntry* lg = reinterpret_cast<WEntry*> (h->blobLog.data() + h->t.oEnt);
  bool ok = true;
  const double lim = 30. * 0.6931471805599453;
  for (size_t n = 0; n < h->entPerm.size(); ++n) {
    if (h->entPerm[n] < 0) { lin[n].w = 0.; lg[n].w = -INFINITY; continue; }
    const double lw = m->hInc.lw[h->entPerm[n]];
    lg[n].w = lw;
    lin[n].w = exp (lw);
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
  const size_t eBytes = std::max<size_t> (h->entPerm.size(), 1) * sizeof (WEntry);
  MB_CUDA (cudaMemcpy (h->dLin + h->t.oEnt, h->blobLin.data() + h->t.oEnt, eBytes, cudaMemcpyHostToDevice));
  MB_CUDA (cudaMemcpy (h->dLog + h->t.oEnt, h->blobLog.data() + h->t.oEnt, eBytes, cudaMemcpyHostToDevice));
  return 0;
}

int wide_prepare (mb_machine* m) {
  WHost* h = new WHost;
  m->wide = h;
  WBuilder B, best;
  std::vector<uint32_t> phM, phD, phI, bM, bD, bI;
  std::vector<uint16_t> live, bLive;
  uint32_t s0 = 0, s1 = 0, b0 = 0, b1 = 0;
  bool hasMatch = false;
  int maxList = 0;
  double bestCost = 1e300;
  int forceG = 0;
  if (const char* e = getenv ("MB_WIDE_G")) forceG = atoi (e);
  for (int G: { 32, 16, 8 }) {
    if (forceG && G != forceG) continue;
    double perCell = 0;
    build_tables (m, G, B, phM, phD, phI, s0, s1, live, hasMatch, maxList, perCell);
    // the sweep is latency-bound below ~16 warps per SM: weigh the instruction estimate by the warps that fit
    const size_t tabBytes = B.ent.size() * sizeof (WEntry) + B.lanes.size() * sizeof (WLane) + B.rounds.size() * sizeof (WRound) + 4096;
    const size_t perWarp = (size_t) (32 / G) * ((hasMatch ? 3 : 2) * ((size_t) m->S * 8 + 8) + (size_t) m->S * 2);
    const size_t kSmem = 227 * 1024;
    const size_t room = tabBytes + 2 * perWarp <= kSmem ? kSmem - tabBytes : kSmem;      // tables in shared memory when two warps still fit
    const int warps = (int) std::min<size_t> (17, room / perWarp);
    if (warps < 2 && !(forceG && G == forceG)) continue;
    perCell *= std::max (1.0, 16.0 / std::max (warps, 1)) * (room == kSmem && tabBytes + 2 * perWarp > kSmem ? 1.5 : 1.0);
    if (perCell < bestCost) { bestCost = perCell; best = B; bM = phM; bD = phD; bI = phI; bLive = live; b0 = s0; b1 = s1; }
  }
  if (bestCost >= 1e300) { set_error ("wide engine: a cell does not fit in shared memory"); return 1; }
  if (maxList > 16383) { set_error ("wide engine: a transition list has more than 16383 entries"); return 1; }
  {      // spare row and round: the sweep keeps one entry and one lane word in flight
    WEntry padE; padE.w = 0; padE.srcOff = 0; padE.pad = 0;
    WLane padL; padL.a = 0xffffu; padL.b = 0;
    best.ent.resize (best.ent.size() + best.G, padE);
    best.perm.resize (best.perm.size() + best.G, -1);
    best.lanes.resize (best.lanes.size() + best.G, padL);
    WRound padR; padR.rowStart = 0; padR.nRows = 0; padR.flags = 0;
    best.rounds.push_back (padR);
  }
  if (best.ent.size() / best.G > 0xfffffff0u) { set_error ("wide engine: tables too large"); return 1; }
  WideTables& t = h->t;
  std::vector<char>& blob = h->blobLin;
  t.oEnt = put_vec (blob, best.ent); t.oRound = put_vec (blob, best.rounds); t.oLane = put_vec (blob, best.lanes);
  t.oPhM = put_vec (blob, bM); t.oPhD = put_vec (blob, bD); t.oPhI = put_vec (blob, bI); t.oLive = put_vec (blob, bLive);
  blob.resize ((blob.size() + 15) & ~(size_t) 15, 0);
  t.bytes = (uint32_t) blob.size();
  t.silR0 = b0; t.silR1 = b1;
  t.S = m->S; t.nIn = m->nIn; t.nOut = m->nOut; t.hasMatch = hasMatch ? 1 : 0;
  t.nLiveIn = (int32_t) bLive.size(); t.bpBytes = maxList <= 63 ? 1 : 2; t.G = best.G;
  h->maxList = maxList;
  h->estInstrPerCell = bestCost;
  h->entPerm = best.perm;
  h->blobLog = h->blobLin;
  wide_fill_weights (m, h);
  MB_CUDA (cudaSetDevice (m->device));
  MB_CUDA (cudaMalloc (&h->dLin, t.bytes));
  MB_CUDA (cudaMalloc (&h->dLog, t.bytes));
  MB_CUDA (cudaMemcpy (h->dLin, h->blobLin.data(), t.bytes, cudaMemcpyHostToDevice));
  MB_CUDA (cudaMemcpy (h->dLog, h->blobLog.data(), t.bytes, cudaMemcpyHostToDevice));
  MB_CUDA (cudaDeviceGetAttribute (&h->numSMs, cudaDevAttrMultiProcessorCount, m->device));
  if (getenv ("MB_WIDE_VERBOSE"))
    fprintf (stderr, "wide engine: S=%d G=%d rounds=%zu rows=%zu tables=%u bytes, est. %.0f warp-instructions per cell, live-in %d, bp %d bytes\n",
             m->S, best.G, best.rounds.size(), best.ent.size() / best.G, t.bytes, bestCost, t.nLiveIn, t.bpBytes);
  return 0;
}

// launch shape: column warps per CTA (W), ring depth, where the tables live, CTAs per SM
struct WShape { int W = 0, R = 2, tabInSmem = 0, ctasPerSM = 0; size_t smem = 0; bool oneD = false; const void* fn = nullptr; };

template<int OP> static const void* kernel_for (int G, bool tabs) {
  if (G == 32) return tabs ? (const void*) wide_kernel<OP, 32, true> : (const void*) wide_kernel<OP, 32, false>;
  if (G == 16) return tabs ? (const void*) wide_kernel<OP, 16, true> : (const void*) wide_kernel<OP, 16, false>;
  return tabs ? (const void*) wide_kernel<OP, 8, true> : (const void*) wide_kernel<OP, 8, false>;
}

template<int OP>
static int choose_shape (const WHost* h, bool oneD, WShape& best) {
  const WideTables& t = h->t;
  const int R = t.hasMatch ? 3 : 2, CPW = 32 / t.G;
  const size_t kMaxSmem = 227 * 1024;
  double bestScore = -1;
  static const int cand[] = { 16, 12, 8, 6, 4, 3, 2, 1, 0 };
  int forceW = 0;
  if (const char* e = getenv ("MB_WIDE_W")) forceW = atoi (e);
  for (int tabIn = 1; tabIn >= 0; --tabIn) {
    const void* fn = kernel_for<OP> (t.G, tabIn != 0);
    MB_CUDA (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxSmem));
    for (int W: cand) {
      if (forceW && W != forceW) continue;
      if (W == 0 && !oneD) continue;      // the 2-D sweep needs the loader warp and at least one column warp
      const int nWarps = W + 1, nSlots = nWarps * CPW;
      const size_t smem = (tabIn ? ((size_t) t.bytes + 15) & ~(size_t) 15 : 0) + (size_t) nSlots * R * t.S * 8 + (size_t) nSlots * R * 8
        + (OP == OP_MAX ? (size_t) nSlots * t.S * 2 : 0) + 32;
      if (smem > kMaxSmem) continue;
      int ctas = 0;
      MB_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&ctas, fn, nWarps * 32, smem));
      if (ctas < 1) continue;
      const double score = (double) (oneD ? nWarps : W) * ctas * (tabIn ? 1.0 : 0.4);
      if (score > bestScore) { bestScore = score; best.W = W; best.R = R; best.tabInSmem = tabIn; best.ctasPerSM = ctas; best.smem = smem; best.oneD = oneD; best.fn = fn; }
    }
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
  const int nWarps = sh.W + 1, CPW = 32 / h->t.G;
  const int64_t nWork = (int64_t) order.size();
  const int64_t wantCtas = oneD ? (nWork + nWarps * CPW - 1) / (nWarps * CPW) : nWork;
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
  p.W = sh.W; p.R = sh.R; p.oneD = oneD ? 1 : 0;
  p.b = b->dev;
  p.order = dOrder.as<int64_t>(); p.nWork = nWork; p.counter = dCounter.as<unsigned long long>();
  p.result = dResult; p.flag = dFlag;
  p.bnd = dBnd.as<double>(); p.bndFG = dBndFG.as<int2>(); p.bndRows = bndRows;
  p.bp = dBp; p.bpOff = dBpOff;
  void* args[] = { &p };
  MB_CUDA (cudaLaunchKernel (sh.fn, dim3 ((unsigned) grid), dim3 ((unsigned) (nWarps * 32)), args, sh.smem, b->stream));
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
