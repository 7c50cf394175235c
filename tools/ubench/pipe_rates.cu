// Issue-rate microbenchmarks for the instructions the softmax warps of the attention kernel are made of
// (MUFU.EX2, F2FP pack, FMNMX3, FFMA) -- warp-instructions per clock per SM sub-partition on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void rate_kernel(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed * (threadIdx.x + k + 1) * 1e-3f;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k])); }
      if (OP == 1) { a[k] = fmaf(a[k], 1.0001f, 0.5f); }
      if (OP == 2) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(a[(k + 1) & 7]), "f"(seed)); }
      if (OP == 3) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[k]), "f"(a[(k + 1) & 7])); acc ^= r; }
      if (OP == 5) { unsigned r = __float_as_uint(a[k]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[k] = __uint_as_float(r); }
      if (OP == 6 && (k & 1) == 0) {
        unsigned long long x = ((unsigned long long)__float_as_uint(a[k + 1]) << 32) | __float_as_uint(a[k]), m = 0x3f8000013f800001ull, c = 0x3f0000003f000000ull;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(m), "l"(c));
        a[k] = __uint_as_float((unsigned)x); a[k + 1] = __uint_as_float((unsigned)(x >> 32));
      }
      if (OP == 7) { unsigned r = __float_as_uint(a[k]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r)); a[k] = __uint_as_float(r); }
      if (OP == 4) {   // the softmax inner loop shape: FFMA -> EX2 -> FADD (+ pack per pair)
        float p; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(a[k], 0.18f, -seed)));
        a[k] += p;
        if (k & 1) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(p), "f"(a[k - 1])); acc ^= r; }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
}

template <int OP>
void run(const char* name, int warps_per_smsp, float insts_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int threads = 128 * warps_per_smsp, iters = 4096;
  float* out; cudaMalloc(&out, sizeof(float) * sms * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<OP><<<sms, threads>>>(out, iters, 1.0f);
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); rate_kernel<OP><<<sms, threads>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  const double clks = best * 1e-3 * khz * 1e3;
  const double per_smsp = (double)iters * insts_per_iter * warps_per_smsp;
  printf("%-28s warps/SMSP=%d  %.3f ms  clk/warp-inst/SMSP = %.2f  (nominal clock %d MHz)\n", name, warps_per_smsp, best,
         clks / per_smsp, khz / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<0>("MUFU.EX2", w, 8);
    run<1>("FFMA", w, 8);
    run<2>("FMNMX3", w, 8);
    run<3>("F2FP.PACK", w, 8);
    run<4>("FFMA+EX2+FADD(+pack/2)", w, 8);
    run<5>("MUFU.EX2.F16x2", w, 8);
    run<6>("FFMA2 (f32x2)", w, 4);
  }
  return 0;
}
