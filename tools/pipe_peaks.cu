// tools/pipe_peaks.cu -- microbenchmark of the per-SM issue rates that bound the DP kernels
// (MEASURED_PEAKS.json has only HBM and bf16 numbers).  For each op class it
// runs 8 independent dependency chains per thread, 1024 threads per CTA, ONE CTA per SM, and
// reports thread-level ops per clock per SM (from clock64) and ops/s chip-wide (from CUDA events).
// (Round 1 launched two CTAs per SM and assumed both resident; at ~40 registers per thread only one is, the two
// ran back to back, and the per-clock column came out twice too high at half the clock.  The Gop/s column, taken
// from CUDA events over the whole launch, was right and is what the rooflines use.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

enum Op { DADD, DFMA, DMAX, DSETSEL, F2F_DOWN, F2F_UP, EX2, LG2, FFMA, LSE_MIXED, LSE_F32, ICMP64SEL, DSETP_PTR, NOPS };
static const char* names[NOPS] = { "dadd", "dfma", "dmax", "dsetp_sel", "f2f_f64_to_f32", "f2f_f32_to_f64", "mufu_ex2", "mufu_lg2", "ffma",
                                   "lse_f64_mixed", "lse_f32", "icmp64_sel", "dadd_dsetp_sel_ptr" };

template<int OP>
__global__ void __launch_bounds__(1024, 1) k (double* out, long long* cyc, double seed) {
  double d[CHAINS];
  float f[CHAINS];
  unsigned ptr = 0;
  for (int c = 0; c < CHAINS; ++c) { d[c] = seed + c * 1e-3 + threadIdx.x * 1e-6; f[c] = (float) d[c]; }
  const double w = seed * 0.999;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == DADD) d[c] = d[c] + w;
      if (OP == DFMA) d[c] = fma (d[c], w, w);
      if (OP == DMAX) d[c] = fmax (d[c] * 1.0, d[(c + 1) % CHAINS]);
      if (OP == DSETSEL) d[c] = d[c] < d[(c + 1) % CHAINS] ? w : d[c];
      if (OP == F2F_DOWN) { f[c] = (float) d[c]; d[c] = __hiloint2double (__float_as_int (f[c]), __double2loint (d[c])); }
      if (OP == F2F_UP) { d[c] = (double) f[c]; f[c] = __int_as_float (__double2hiint (d[c])); }
      if (OP == EX2) f[c] = exp2f (f[c]);
      if (OP == LG2) f[c] = __log2f (f[c]);
      if (OP == FFMA) f[c] = fmaf (f[c], 0.999f, 0.5f);
      if (OP == LSE_MIXED) {   // the device log-sum-exp: FP64 max/diff, FP32 softplus via MUFU
        const double a = d[c], b = d[(c + 1) % CHAINS] + w;
        const double mx = fmax (a, b);
        const float x = (float) fabs (a - b);
        const float sp = x < 10.f ? __log2f (1.f + exp2f (-1.4426950408889634f * x)) * 0.6931471805599453f : 0.f;
        d[c] = mx + (double) sp;
      }
      if (OP == ICMP64SEL) {   // the same select with the compare done on the bit patterns (two integer compares)
        const unsigned long long a = (unsigned long long) __double_as_longlong (d[c]), b = (unsigned long long) __double_as_longlong (d[(c + 1) % CHAINS]);
        d[c] = a > b ? w : d[c];
      }
      if (OP == DSETP_PTR) {   // one Viterbi candidate: add, compare, select, pointer bit under the predicate
        const double t = d[(c + 1) % CHAINS] + w;
        asm ("{ .reg .pred p; setp.lt.f64 p, %0, %2; selp.f64 %0, %2, %0, p; @p or.b32 %1, %1, %3; }" : "+d"(d[c]), "+r"(ptr) : "d"(t), "r"(1u << c));
      }
      if (OP == LSE_F32) {
        const float a = f[c], b = f[(c + 1) % CHAINS] + 0.25f;
        const float mx = fmaxf (a, b);
        const float x = fabsf (a - b);
        f[c] = mx + (x < 10.f ? __log2f (1.f + exp2f (-1.4426950408889634f * x)) * 0.6931471805599453f : 0.f);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += d[c] + f[c];
  s += ptr;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int OP> void run (int sms, double* out, long long* cyc) {
  const int grid = sms;
  cudaEvent_t e0, e1;
  cudaEventCreate (&e0); cudaEventCreate (&e1);
  k<OP><<<grid, 1024>>> (out, cyc, 1.0000001);
  cudaDeviceSynchronize();
  cudaEventRecord (e0);
  k<OP><<<grid, 1024>>> (out, cyc, 1.0000001);
  cudaEventRecord (e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime (&ms, e0, e1);
  long long h[4096];
  cudaMemcpy (h, cyc, sizeof (long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double) h[i];
  mean /= grid;
  const double opsPerSM = 1024.0 * (double) ITERS * CHAINS;     // one resident CTA per SM
  printf ("  {\"op\": \"%s\", \"ops_per_clk_per_sm\": %.2f, \"gops_per_s\": %.1f, \"ms\": %.3f, \"sm_mhz_effective\": %.0f}",
          names[OP], opsPerSM / mean, opsPerSM * sms / (ms * 1e6), ms, mean / (ms * 1e3));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties (&p, 0);
  double* out; long long* cyc;
  cudaMalloc (&out, sizeof (double) * 1024 * 4096);
  cudaMalloc (&cyc, sizeof (long long) * 4096);
  printf ("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, p.multiProcessorCount);
  run<DADD> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<DFMA> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<DMAX> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<DSETSEL> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<F2F_DOWN> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<F2F_UP> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<EX2> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<LG2> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<FFMA> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<LSE_MIXED> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<LSE_F32> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<ICMP64SEL> (p.multiProcessorCount, out, cyc); printf (",\n");
  run<DSETP_PTR> (p.multiProcessorCount, out, cyc); printf ("\n]}\n");
  return 0;
}
