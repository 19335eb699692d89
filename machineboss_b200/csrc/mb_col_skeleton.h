// mb_col_skeleton.h -- CUDA source of the column engine's strip kernel, compiled at run time with NVRTC for sm_100a
// after the generated prelude (defines + the machine's cell, see mb_col.cu).
//
// What it is for.  A generator without input (a profile HMM, or one composed with an error model: SURVEY.md
// section 8 config 5) scores a read in a matrix of ONE column per read position and thousands of states per cell;
// but the states of such a machine repeat with a period (the nodes of the profile), and transitions only run
// inside a period or to the next one.  mb_col.cu finds that period and re-reads the machine as a two-dimensional
// recurrence: COLUMN k = period k of the machine, ROW o = read position, a cell of P + (a few) states with the
// transition weights depending on the column.  That is the shape the strip engines sweep at full speed.
//
// Mapping.  Strips of 32 * MB_COL_C columns, lane j owns MB_COL_C adjacent columns of the strip (their cells at one row
// are computed one after the other in the same step: the second reads the first's fresh values from registers), rows
// swept in a skew (at step t lane j is at row t - j), as the big engine (mb_big_skeleton.h), with three differences:
//   * the weights belong to the column: one table per strip, [slot][column of the lane][lane], staged in shared memory
//     for the whole CTA; the silent slots of a lane's columns live in registers for the strip (MB_COL_DECLW) when there
//     are few enough;
//   * all warps of a CTA work on the SAME strip (of different reads) so that they share that table: a CTA claims
//     nWarps * R reads at a time and takes them through the strips together;
//   * the first strip's left boundary is not empty: it holds, per row, the values of the machine's prefix states
//     that feed the periods ("carried" states, copied from column to column), written by the prefix kernel; and the
//     last column's left-going states (the "accumulators" that collect what the periods send to the machine's
//     suffix states) are written back for the suffix kernel.  One boundary buffer per read, updated in place: the
//     rows a strip reads are always ahead of the rows its last lane writes.
//
// Arithmetic: MODE 0, scaled linear domain with a power-of-two frame per lane, renormalised every MB_COL_RESCALE
// steps (as MB_LANE_FRAMES of mb_jit_skeleton.h: the left neighbour's values enter through 2^(its frame - mine),
// boundary rows carry their frame; a spread above 2^700 or a frame more than 2^900 away flags the read for the
// lane engine's log-domain sweep).  MODE 2: max-plus in the log domain (Viterbi scores; FP64 add + compare, exact).
// MODE 1: the same with every cell's pointers stored -- per cell state the index of the winning candidate among the
// state's groups, packed into MB_BPBYTES bytes -- for col_traceback_kernel (mb_col.cu).
#ifndef MB_COL_SKELETON_H
#define MB_COL_SKELETON_H

static const char* const kColSkeleton = R"MBSRC(
#define MB_FULL 0xffffffffu
#define MB_COL_RESCALE 16
#define MB_BROW (MB_NLL + 1)      // doubles per boundary row: the left-going states + the frame

struct MBColArgs {
  const uint8_t* y; const int64_t* yOff;
  const int64_t* order; int64_t nWork; unsigned long long* counter;
  double* bnd; const int64_t* bndOff;      // work item n: (Lo + 1) rows of MB_BROW doubles at bnd + bndOff[n]
  const double* tab;                       // [strip][slot][column of the lane][lane] weights
  int32_t* flag;
  int nStrips, K, R, pad;
  // MODE 1: work item n's pointers at bp + bpOff[n], in SWEEP ORDER: one block per (strip, step) holding what the warp's lanes
  // produced in that step -- [(strip * (Lo + 32) + step) * 32 + lane] groups of MB_COL_C * MB_BPBYTES bytes -- so that a step's
  // store is one contiguous run (row-major rows would scatter the skewed lanes over 32 sectors: measured 2 x the kernel time)
  unsigned char* bp; const int64_t* bpOff;
};

__device__ __forceinline__ double mb_pow2 (int d) { return __hiloint2double ((1023 + d) << 20, 0); }
__device__ __forceinline__ void mb_lds_lane0 (double& v, const int lane, const unsigned addr) {
  asm volatile ("{ .reg .pred p; setp.eq.s32 p, %1, 0; @p ld.shared.f64 %0, [%2]; }" : "+d"(v) : "r"(lane), "r"(addr));
}

template<int MODE>
__device__ __forceinline__ void mb_col_run (const MBColArgs& A) {
  constexpr bool LIN = MODE == 0;
  const double ZERO = LIN ? 0.0 : __longlong_as_double (0xfff0000000000000LL);
  extern __shared__ double mb_smem[];
  __shared__ unsigned long long sBase;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  double* sTab = mb_smem;
  double* sIn = mb_smem + MB_NSLOTS * 32 * MB_COL_C + warp * (MB_COL_RESCALE * MB_NLL);
  const double* W = sTab + lane;
  const unsigned sInAddr = (unsigned) __cvta_generic_to_shared (sIn);

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sBase = atomicAdd (A.counter, (unsigned long long) (nWarps * A.R));
    __syncthreads();
    const int64_t base = (int64_t) sBase;
    if (base >= A.nWork) break;

    for (int strip = 0; strip < A.nStrips; ++strip) {
      __syncthreads();
      {
        const double* src = A.tab + (int64_t) strip * (MB_NSLOTS * 32 * MB_COL_C);
        for (int q = threadIdx.x; q < MB_NSLOTS * 32 * MB_COL_C; q += blockDim.x) sTab[q] = __ldg (src + q);
      }
      __syncthreads();
      MB_COL_DECLW      // the silent weights of my column, in registers for the strip
      const int col = (strip * 32 + lane) * MB_COL_C;      // my first column; A.K is a multiple of MB_COL_C (padded with empty columns)
      const bool inCol = col < A.K;
      const int outLane = (strip + 1 < A.nStrips) ? 31 : ((A.K / MB_COL_C - 1) & 31);

      for (int rr = 0; rr < A.R; ++rr) {
        const int64_t w = base + (int64_t) rr * nWarps + warp;
        if (w >= A.nWork) break;
        const int64_t k = A.order[w];
        const int64_t y0 = A.yOff[k];
        const int Lo = (int) (A.yOff[k + 1] - y0);
        const uint8_t* y = A.y + y0;
        double* bnd = A.bnd + A.bndOff[w];
        unsigned char* bp = MODE == 1 ? A.bp + A.bpOff[w] + ((int64_t) strip * (Lo + 32) * 32 + lane) * (MB_COL_C * MB_BPBYTES) : (unsigned char*) 0;
        int suspect = 0;

        double U[MB_NUREG > 0 ? MB_NUREG : 1];      // my columns' last cells: the sources of the groups that consume a token and stay in the column
        double Lown[MB_NLL];                        // my last column's left-going states
        double Lprev[MB_NDREG > 0 ? MB_NDREG : 1];  // sources of diagonal groups, a row ago: [0, MB_NLD) what the left lane sent, then my own columns but the last
#pragma unroll
        for (int q = 0; q < (MB_NUREG > 0 ? MB_NUREG : 1); ++q) U[q] = ZERO;
#pragma unroll
        for (int j = 0; j < MB_NLL; ++j) Lown[j] = ZERO;
#pragma unroll
        for (int j = 0; j < (MB_NDREG > 0 ? MB_NDREG : 1); ++j) Lprev[j] = ZERO;
        int ecur = LIN ? (int) __ldcg (bnd + MB_NLL) : 0;      // frame: true value = stored value * 2^ecur
        double gl = 1.0;                                        // 2^(left neighbour's frame - mine)
        double stageNext[MB_NLL], stageNextE = (double) ecur;   // boundary row t + lane of the next block (lanes < MB_COL_RESCALE)
#pragma unroll
        for (int j = 0; j < MB_NLL; ++j)
          stageNext[j] = (lane < MB_COL_RESCALE && lane <= Lo) ? __ldcg (bnd + (int64_t) lane * MB_BROW + j) : ZERO;
        if (LIN && lane < MB_COL_RESCALE && lane <= Lo) stageNextE = __ldcg (bnd + (int64_t) lane * MB_BROW + MB_NLL);
        int tokNext = 0;      // output token of my row at the next step (row 0 has none)
        const int nSteps = Lo + 32;

        for (int t = 0; t < nSteps; ++t) {
          if ((t & (MB_COL_RESCALE - 1)) == 0) {
            bool nz = false;
            if (LIN && t > 0) {      // renormalise my values to [1, 2)
              int mh = 0;
              unsigned ml = 0xffffffffu;
#pragma unroll
              for (int q = 0; q < MB_NUREG; ++q) { const int h = __double2hiint (U[q]); mh = max (mh, h); ml = min (ml, (unsigned) (h - 1)); }
#pragma unroll
              for (int j = 0; j < MB_NLL; ++j) { const int h = __double2hiint (Lown[j]); mh = max (mh, h); ml = min (ml, (unsigned) (h - 1)); }
#pragma unroll
              for (int j = 0; j < MB_NDREG; ++j) { const int h = __double2hiint (Lprev[j]); mh = max (mh, h); ml = min (ml, (unsigned) (h - 1)); }
              nz = mh >= 0x00100000;
              if (nz) {
                const int ex = mh >> 20;
                const int shift = min (ex - 1023, 1000);
                if (shift != 0) {
                  const double f = mb_pow2 (-shift);
#pragma unroll
                  for (int q = 0; q < MB_NUREG; ++q) U[q] *= f;
#pragma unroll
                  for (int j = 0; j < MB_NLL; ++j) Lown[j] *= f;
#pragma unroll
                  for (int j = 0; j < MB_NDREG; ++j) Lprev[j] *= f;
                  ecur += shift;
                }
                if (ml != 0xffffffffu && ex - (int) ((ml + 1u) >> 20) > 700) suspect |= 1;
              }
            }
            // A lane that holds nothing yet takes the frame of what is about to reach it: the nearest lane to its left
            // that holds something, or -- left of all of those -- the first staged boundary row that is not all zero.
            if (LIN) {
              bool rowNz = false;
              if (lane < MB_COL_RESCALE) {
#pragma unroll
                for (int j = 0; j < MB_NLL; ++j) rowNz |= stageNext[j] != 0.0;
              }
              const unsigned rowMask = __ballot_sync (MB_FULL, rowNz), nzMask = __ballot_sync (MB_FULL, nz);
              const double eRow = __shfl_sync (MB_FULL, stageNextE, rowMask ? __ffs (rowMask) - 1 : 0);
              if (lane == 0 && !nz) ecur = (int) eRow;
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
            {      // stage rows t .. t+MB_COL_RESCALE-1 of the boundary, in lane 0's frame
              const int e0 = __shfl_sync (MB_FULL, ecur, 0);
              __syncwarp();
              if (lane < MB_COL_RESCALE) {
                int d = (int) stageNextE - e0;
                const bool far = d < -900 || d > 900;
                bool any = false;
                d = max (min (d, 1000), -1023);
                const double f = mb_pow2 (d);
#pragma unroll
                for (int j = 0; j < MB_NLL; ++j) { any |= stageNext[j] != 0.0; sIn[lane * MB_NLL + j] = LIN ? stageNext[j] * f : stageNext[j]; }
                if (LIN && far && any) suspect |= 4;
                const int rowN = t + MB_COL_RESCALE + lane;
                const double* src = bnd + (int64_t) rowN * MB_BROW;
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
            double v = __shfl_up_sync (MB_FULL, Lown[j], 1);
            if (LIN) v *= gl;
            mb_lds_lane0 (v, lane, sInAddr + (unsigned) (((t & (MB_COL_RESCALE - 1)) * MB_NLL + j) * 8));      // lane 0 takes the staged boundary row: a predicated load, no select
            Lin[j] = v;
          }
          if (r >= 0 && r <= Lo && inCol) {
            if (LIN) { MB_COL_CELL_LIN }
            else if (MODE == 2) { MB_COL_CELL_MAX }
            else {
              unsigned pk[MB_COL_C * MB_NPW];
              MB_COL_CELL_MAXP
              unsigned char* dstp = bp + (int64_t) t * (32 * MB_COL_C * MB_BPBYTES);
              if (MB_COL_C * MB_BPBYTES == 4 && MB_BPBYTES < 4) {      // the lane's columns in one 32-bit store
                unsigned word = 0u;
#pragma unroll
                for (int c = 0; c < MB_COL_C; ++c) word |= pk[c] << (8 * MB_BPBYTES * c);
                *(unsigned*) dstp = word;
              } else
#pragma unroll
              for (int c = 0; c < MB_COL_C; ++c) {
                if (MB_BPBYTES == 1) dstp[c] = (unsigned char) pk[c];
                else if (MB_BPBYTES == 2) ((unsigned short*) dstp)[c] = (unsigned short) pk[c];
                else {
#pragma unroll
                  for (int q = 0; q < MB_NPW; ++q) ((unsigned*) dstp)[c * MB_NPW + q] = pk[c * MB_NPW + q];
                }
              }
            }
            if (lane == outLane) {
              double* dst = bnd + (int64_t) r * MB_BROW;
#pragma unroll
              for (int j = 0; j < MB_NLL; ++j) dst[j] = Lown[j];
              if (LIN) dst[MB_NLL] = (double) ecur;
            }
          }
          MB_COL_KEEPDIAG      // Lprev <- the left-going states just received that diagonal groups read
        }
        suspect = __reduce_or_sync (MB_FULL, (unsigned) suspect);      // why: 1 spread, 2 neighbour frame, 4 boundary frame
        if (LIN && lane == 0 && suspect) atomicOr (A.flag + k, suspect);
        __syncwarp();
      }
    }
  }
}

extern "C" __global__ void __launch_bounds__(MB_COL_THREADS, MB_COL_MINBLOCKS) mb_k_col_sum (const __grid_constant__ MBColArgs A) { mb_col_run<0> (A); }
extern "C" __global__ void __launch_bounds__(MB_COL_THREADS, MB_COL_MINBLOCKS) mb_k_col_max (const __grid_constant__ MBColArgs A) { mb_col_run<2> (A); }
extern "C" __global__ void __launch_bounds__(MB_COL_THREADS, MB_COL_MINBLOCKS) mb_k_col_maxp (const __grid_constant__ MBColArgs A) { mb_col_run<1> (A); }
)MBSRC";

#endif
