// accuracy of ex2.approx.f16x2 over the softmax input range
#include <cstdio>
#include <cmath>
#include <cuda_fp16.h>
__global__ void k(float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  float x = -16.0f * i / n;           // [-16, 0]
  __half2 h = __floats2half2_rn(x, x); unsigned r = *(unsigned*)&h;
  asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
  __half2 o = *(__half2*)&r; out[2*i] = __low2float(o); out[2*i+1] = x;
}
int main() {
  const int n = 1 << 16; float* d; cudaMalloc(&d, 8 * n); k<<<n / 256, 256>>>(d, n);
  float* h = new float[2 * n]; cudaMemcpy(h, d, 8 * n, cudaMemcpyDeviceToHost);
  double worst_in = 0, worst_tot = 0, sq = 0; 
  for (int i = 0; i < n; ++i) {
    double x = h[2*i+1]; float xr = __half2float(__float2half((float)x));
    double e_fn = fabs(h[2*i] / exp2((double)xr) - 1.0);   // error of the function given the rounded input
    double e_tot = fabs(h[2*i] / exp2(x) - 1.0);           // incl. input rounding
    if (x > -14 && e_fn > worst_in) worst_in = e_fn; if (x > -8 && e_tot > worst_tot) worst_tot = e_tot; if (x > -8) sq += e_tot*e_tot;
  }
  printf("ex2.f16x2: max rel err given fp16 input %.3e; incl. input rounding (x>-8) max %.3e rms %.3e (fp16 rounding alone: max 4.9e-4)\n", worst_in, worst_tot, sqrt(sq/(n/2)));
}
