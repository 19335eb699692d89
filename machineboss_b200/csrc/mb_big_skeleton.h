// mb_big_skeleton.h -- CUDA source of the strip kernel for mid-size machines (tens to hundreds of states:
// composed transducers such as prot2dna => dnapsw), compiled at run time with
// NVRTC for sm_100a after the generated prelude (defines + the machine's cell function, see mb_big.cu).
//
// Mapping.  One warp per sequence pair; the matrix is cut into vertical strips of 32 input positions and
// LANE j OWNS COLUMN j of the strip; rows are swept in a skew (at step t lane j computes output row t-j).
// The whole cell -- every transition group of the machine, in state order, which is a topological order
// of the silent groups -- is straight-line code executed by one thread, so there is no barrier, no flag and
// no gather index anywhere: silent weights are constant-bank operands, emission weights are read from
// shared-memory tables by token, the states of the cell above (the sources of insert groups: "live-up"
// states) come from the lane's own column of a shared-memory array v[state][lane], and the few states
// the cell to the left contributes (sources of delete / match groups) arrive with one shuffle each.  What
// the cell leaves behind is the live-up states (written back in place) and the left-going ones
// (registers); everything else lives and dies in registers (the compiler spills what does not fit).
//
// Arithmetic: scaled linear domain with a power-of-two frame per LANE, renormalised every MB_BIG_RESCALE
// steps, exactly as the score-only sweeps of mb_jit_skeleton.h (MB_LANE_FRAMES): the left neighbour's values
// enter through 2^(its frame - mine); strip-boundary rows carry their frame; a spread above 2^700, a
// neighbour or boundary frame more than 2^900 away or a zero result flags the pair for the log-domain
// sweep (wide engine).
//
// MODE 0: that Forward sweep.  MODE 1 / 2: Viterbi (viterbi.cpp:18-47) with / without back-pointers: the same
// sweep in the log domain, FP64 add + strict '<' over each state's groups in the reference's candidate order
// (bit-exact scores, first maximum wins); no frames.  A cell's pointers (ceil(log2(#groups)) bits per state,
// MB_NPW 32-bit words) are stored in sweep order, bp[((step * nStrips + strip) * MB_NPW + word) * 32 + lane] with step = row +
// lane: every store of the warp is one 128-byte line (indexed by row the skewed lanes would scatter over 32 lines).
#ifndef MB_BIG_SKELETON_H
#define MB_BIG_SKELETON_H

static const char* const kBigSkeleton = R"MBSRC(
#define MB_FULL 0xffffffffu
#define MB_BIG_RESCALE 16
#define MB_BROW (MB_NLL + 1)      // doubles per strip-boundary row: the left-going states + the frame

struct MBBigArgs {
  const uint8_t* x; const int64_t* xOff;
  const uint8_t* y; const int64_t* yOff;
  const int64_t* order; int64_t nWork; unsigned long long* counter;
  double* bnd; int64_t bndStride;      // per warp: two buffers of (maxLo + 1) boundary rows
  double* result; int32_t* flag;
  const double* emit;
  unsigned* bp; const int64_t* bpOff;      // MODE 1: back-pointer words of work item n at bp + bpOff[n]
  double resLog;                           // linear sweep: log of the scale the end state carries (normalised weights)
};

__device__ __forceinline__ double mb_pow2 (int d) { return __hiloint2double ((1023 + d) << 20, 0); }

template<int MODE>
__device__ __forceinline__ void mb_big_run (const MBBigArgs& A) {
  constexpr bool LIN = MODE == 0;
  const double ZERO = LIN ? 0.0 : __longlong_as_double (0xfff0000000000000LL);
  // The sums keep FOLDED live-up values (mb_big.cu: one per class of proportional insert groups instead of one per source
  // state) and their own emission table; the max-plus sweeps keep the sources themselves (folding would re-associate the
  // additions and lose the bit-exact scores).
  constexpr int NLU = LIN ? MB_NLU_LIN : MB_NLU;
  constexpr int NEMIT = LIN ? MB_NEMIT_LIN : MB_NEMIT;
  extern __shared__ double mb_smem[];
  double* E = mb_smem;
  for (int q = threadIdx.x; q < NEMIT; q += blockDim.x) E[q] = A.emit[q];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  // shared memory: emission tables | per warp: live-up states v[q][lane] | per warp: MB_BIG_RESCALE staged boundary rows
  double* up = mb_smem + ((NEMIT + 1) & ~1) + warp * (NLU * 32) + lane;
  double* sIn = mb_smem + ((NEMIT + 1) & ~1) + nWarps * (NLU * 32) + warp * (MB_BIG_RESCALE * MB_NLL);
  const int64_t wslot = (int64_t) blockIdx.x * nWarps + warp;
  double* bndA = A.bnd + wslot * A.bndStride;
  double* bndB = bndA + (A.bndStride >> 1);

  for (;;) {
    unsigned long long w = 0;
    if (lane == 0) w = atomicAdd (A.counter, 1ULL);
    w = __shfl_sync (MB_FULL, w, 0);
    if ((int64_t) w >= A.nWork) break;
    const int64_t k = A.order[w];
    const int64_t x0 = A.xOff[k], y0 = A.yOff[k];
    const int Li = (int) (A.xOff[k + 1] - x0), Lo = (int) (A.yOff[k + 1] - y0);
    const uint8_t* x = A.x + x0;
    const uint8_t* y = A.y + y0;
    const int nStrips = (Li + 32) / 32;
    int suspect = 0;
    unsigned* bp = MODE == 1 ? A.bp + A.bpOff[w] + lane : (unsigned*) 0;

    for (int strip = 0; strip < nStrips; ++strip) {
      const int col = strip * 32 + lane;
      const bool inCol = col <= Li;
      const int a = (col >= 1 && inCol) ? x[col - 1] - 1 : 0;
      for (int q = 0; q < NLU; ++q) up[q * 32] = ZERO;
      double Lown[MB_NLL], Lprev[MB_NLL];      // my last cell's left-going states; what I received a step ago (the diagonal cell)
#pragma unroll
      for (int j = 0; j < MB_NLL; ++j) { Lown[j] = ZERO; Lprev[j] = ZERO; }
      const double* bin = (strip & 1) ? bndB : bndA;
      double* bout = (strip & 1) ? bndA : bndB;
      const bool hasIn = strip > 0, hasOut = strip + 1 < nStrips;
      int ecur = (LIN && hasIn) ? (int) __ldcg (bin + MB_NLL) : 0;      // frame: true value = stored value * 2^ecur
      double gl = 1.0;                                         // 2^(left neighbour's frame - mine)
      if (!hasIn) {
        __syncwarp();
        for (int q = lane; q < MB_BIG_RESCALE * MB_NLL; q += 32) sIn[q] = ZERO;
        __syncwarp();
      }
      double stageNext[MB_NLL], stageNextE = (double) ecur;      // boundary row t + lane of the next block (lanes < MB_BIG_RESCALE)
#pragma unroll
      for (int j = 0; j < MB_NLL; ++j)
        stageNext[j] = (hasIn && lane < MB_BIG_RESCALE && lane <= Lo) ? __ldcg (bin + (int64_t) lane * MB_BROW + j) : ZERO;
      if (LIN && hasIn && lane < MB_BIG_RESCALE && lane <= Lo) stageNextE = __ldcg (bin + (int64_t) lane * MB_BROW + MB_NLL);
      int tokNext = 0;      // output token of my row at the next step (row 0 has none)
      const int nSteps = Lo + 32;
      double res = ZERO;

      for (int t = 0; t < nSteps; ++t) {
        if ((t & (MB_BIG_RESCALE - 1)) == 0) {
          bool nz = false;
          if (LIN && t > 0) {      // renormalise my values to [1, 2)
            int mh = 0;
            unsigned ml = 0xffffffffu;
            for (int q = 0; q < NLU; ++q) { const int h = __double2hiint (up[q * 32]); mh = max (mh, h); ml = min (ml, (unsigned) (h - 1)); }
#pragma unroll
            for (int j = 0; j < MB_NLL; ++j) {
              const int h = __double2hiint (Lown[j]), g = __double2hiint (Lprev[j]);
              mh = max (mh, max (h, g)); ml = min (ml, min ((unsigned) (h - 1), (unsigned) (g - 1)));
            }
            nz = mh >= 0x00100000;
            if (nz) {
              const int ex = mh >> 20;
              const int shift = min (ex - 1023, 1000);
              if (shift != 0) {
                const double f = mb_pow2 (-shift);
                for (int q = 0; q < NLU; ++q) up[q * 32] *= f;
#pragma unroll
                for (int j = 0; j < MB_NLL; ++j) { Lown[j] *= f; Lprev[j] *= f; }
                ecur += shift;
              }
              if (ml != 0xffffffffu && ex - (int) ((ml + 1u) >> 20) > 700) suspect |= 1;
            }
          }
          // A lane that holds nothing yet takes the frame of what is about to reach it: the nearest lane to its
          // left that holds something, or -- for the lanes left of all of those -- the first staged boundary row
          // that is not all zero (a strip starts with every lane empty, and its first rows are often empty too:
          // cells no path reaches).  Then every lane learns its left neighbour's frame.
          if (LIN) {
            bool rowNz = false;
            if (hasIn && lane < MB_BIG_RESCALE) {
#pragma unroll
              for (int j = 0; j < MB_NLL; ++j) rowNz |= stageNext[j] != 0.0;
            }
            const unsigned rowMask = __ballot_sync (MB_FULL, rowNz), nzMask = __ballot_sync (MB_FULL, nz);
            const double eRow = __shfl_sync (MB_FULL, stageNextE, rowMask ? __ffs (rowMask) - 1 : 0);
            if (lane == 0 && !nz && hasIn) ecur = (int) eRow;
            const unsigned below = nzMask & ((1u << lane) - 1u);
            const int eFrom = __shfl_sync (MB_FULL, ecur, below ? 31 - __clz (below) : 0);
            if (!nz && lane > 0) ecur = eFrom;
            const int eL = __shfl_up_sync (MB_FULL, ecur, 1);
            const bool leftNz = __shfl_up_sync (MB_FULL, (int) nz, 1) != 0;
            int d = lane ? eL - ecur : 0;
            if (leftNz && (d < -900 || d > 900)) suspect |= 2;
            d = max (min (d, 1000), -1022);
            gl = mb_pow2 (d);
          }
          if (hasIn) {      // stage rows t .. t+MB_BIG_RESCALE-1 of the previous strip's last column, in lane 0's frame
            const int e0 = __shfl_sync (MB_FULL, ecur, 0);
            __syncwarp();
            if (lane < MB_BIG_RESCALE) {
              int d = (int) stageNextE - e0;
              const bool far = d < -900 || d > 900;
              bool any = false;
              d = max (min (d, 1000), -1023);
              const double f = mb_pow2 (d);
#pragma unroll
              for (int j = 0; j < MB_NLL; ++j) { any |= stageNext[j] != 0.0; sIn[lane * MB_NLL + j] = LIN ? stageNext[j] * f : stageNext[j]; }
              if (LIN && far && any) suspect |= 4;
              const int rowN = t + MB_BIG_RESCALE + lane;
              const double* src = bin + (int64_t) rowN * MB_BROW;
#pragma unroll
              for (int j = 0; j < MB_NLL; ++j) stageNext[j] = rowN <= Lo ? __ldcg (src + j) : ZERO;
              if (LIN) stageNextE = rowN <= Lo ? __ldcg (src + MB_NLL) : (double) e0;
            }
            __syncwarp();
          }
        }
        const int r = t - lane;
        const int tokb = tokNext;
        tokNext = (r + 1 >= 1 && r + 1 <= Lo) ? y[r] - 1 : 0;      // consumed next step
        double Lin[MB_NLL];
#pragma unroll
        for (int j = 0; j < MB_NLL; ++j) {
          const double fromUp = __shfl_up_sync (MB_FULL, Lown[j], 1);
          const double fromLane = LIN ? fromUp * gl : fromUp;
          Lin[j] = lane ? fromLane : sIn[(t & (MB_BIG_RESCALE - 1)) * MB_NLL + j];
        }
        if (r >= 0 && r <= Lo && inCol) {
          if (LIN) mb_big_cell (up, Lin, Lprev, Lown, a, tokb, r == 0 && col == 0, E, res);
          else {
            unsigned pw[MB_NPW];
            mb_big_cell_vit (up, Lin, Lprev, Lown, a, tokb, r == 0 && col == 0, E, res, pw);
            if (MODE == 1) {
              unsigned* dst = bp + ((int64_t) t * nStrips + strip) * (MB_NPW * 32);      // by STEP, not by row: the lanes of a step are at different rows
#pragma unroll
              for (int q = 0; q < MB_NPW; ++q) dst[q * 32] = pw[q];
            }
          }
          if (hasOut && lane == 31) {
#pragma unroll
            for (int j = 0; j < MB_NLL; ++j) bout[(int64_t) r * MB_BROW + j] = Lown[j];
            if (LIN) bout[(int64_t) r * MB_BROW + MB_NLL] = (double) ecur;
          }
          if (r == Lo && col == Li) {
            if (LIN) A.result[k] = res > 0.0 ? log (res) + (double) ecur * 0.6931471805599453094 + A.resLog : __longlong_as_double (0xfff0000000000000LL);
            else A.result[k] = res;
          }
        }
#pragma unroll
        for (int j = 0; j < MB_NLL; ++j) Lprev[j] = Lin[j];
      }
      suspect = __reduce_or_sync (MB_FULL, (unsigned) suspect);      // why: 1 spread, 2 neighbour frame, 4 boundary frame
      __syncwarp();
    }
    if (LIN && lane == 0) A.flag[k] = suspect;
  }
}

extern "C" __global__ void __launch_bounds__(MB_BIG_THREADS_LIN, 1) mb_k_big_forward (const __grid_constant__ MBBigArgs A) { mb_big_run<0> (A); }
extern "C" __global__ void __launch_bounds__(MB_BIG_THREADS, 1) mb_k_big_viterbi (const __grid_constant__ MBBigArgs A) { mb_big_run<1> (A); }
extern "C" __global__ void __launch_bounds__(MB_BIG_THREADS, 1) mb_k_big_viterbi_score (const __grid_constant__ MBBigArgs A) { mb_big_run<2> (A); }
)MBSRC";

#endif
