#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i])); }
      else if (MODE == 1) { __nv_bfloat162 b = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]); acc ^= *reinterpret_cast<unsigned*>(&b); a[i] += 1.0f; }
      else if (MODE == 2) { a[i] = fmaf(a[i], 1.0001f, 0.5f); }
      else if (MODE == 3) { float t = a[i]; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(t)); a[i] = fmaf(t, 0.5f, a[(i+1)&7]); a[(i+3)&7] = fmaf(a[(i+3)&7], 1.0001f, 0.25f); a[(i+5)&7] = fmaf(a[(i+5)&7], 1.0001f, 0.25f);}
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 4);
  int iters = 4096; int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm; int blocks = 148 * (32 * warps_per_sm / threads);
  k<MODE><<<blocks, threads>>>(out, 16, 0.1f); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 0.1f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 8;
  printf("%-28s warps/SM %2d: %.3f ms  %.2f ops/clk/SM (at 1.9 GHz)\n", name, warps_per_sm, ms, ops / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<0>("ex2.approx", w); run<1>("cvt f32x2->bf16x2 (+fadd)", w); run<2>("ffma", w); run<3>("ex2 + 3 ffma", w);}
  return 0;
}
