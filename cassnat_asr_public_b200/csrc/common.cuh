// Shared device/host helpers for the cassnat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <float.h>

#include "../../include/cassnat_b200.h"

#define CN_LOGZERO (-1e10f)          // reference: src/utils/ctc_prefix.py:12
#define CN_F32_MIN (-FLT_MAX)        // reference: np.finfo(float32).min, modules/attention.py:19

#include <stdio.h>
#define CN_CHECK_LAUNCH()                                                                    \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      fprintf(stderr, "[cassnat_b200] %s:%d CUDA error: %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return CN_ERR_LAUNCH;                                                                  \
    }                                                                                        \
  } while (0)

static inline int cn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch: a kernel launched through cn_launch_pdl may be scheduled while its predecessor in
// the stream is still running -- its CTAs take the SMs the predecessor's tail leaves idle, run their prologue
// (barrier init, TMEM allocation, descriptor prefetch) and then block in cn_pdl_wait() until the predecessor has
// completed and its writes are visible.  EVERY global-memory access of such a kernel comes after cn_pdl_wait();
// cn_pdl_launch() right after it lets the successor do the same (look-ahead of one kernel).  Without the launch
// attribute both instructions are no-ops.  OFF by default (CN_PDL=1 in the environment turns the attribute on): measured
// on the cfg3 decode step, replayed as a CUDA graph, it costs 0.2 ms of 5.3 (5.48-5.53 vs 5.26-5.36 ms/step, two A/B
// pairs on one B200) -- the step's kernels are persistent one-CTA-per-SM grids that cannot co-reside with their
// successor anyway, and the programmatic edges make the graph's kernel-to-kernel hand-over slower, not faster.
#include <stdlib.h>
#include <utility>
__device__ __forceinline__ void cn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cn_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static inline bool cn_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CN_PDL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t cn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = cn_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ float cn_ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 cn_ld_stream_f32x4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ float cn_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float cn_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int cn_warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (value, index) pair ordered like torch.argmax / a stable descending sort:
// larger value first, equal values -> lower index first.
__device__ __forceinline__ bool cn_better(float va, int ia, float vb, int ib) {
  return (va > vb) || (va == vb && ia < ib);
}

// out[c] = sum_blk partial[blk][c] in block order; 32 columns x 8 row groups per CTA.
// blockIdx.y selects one of up to two (partial, out) pairs that share nblk and N (LayerNorm backward: dgamma and dbeta).
static __global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ partial, int64_t nblk, int64_t N,
                                                            float* __restrict__ out, int accumulate,
                                                            const float* __restrict__ partial2 = nullptr,
                                                            float* __restrict__ out2 = nullptr) {
  if (blockIdx.y == 1) { partial = partial2; out = out2; }
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < N) {
    const int64_t per = (nblk + 7) / 8;
    const int64_t b0 = gy * per, b1 = (b0 + per < nblk) ? b0 + per : nblk;
    for (int64_t b = b0; b < b1; ++b) s += partial[b * N + c];
  }
  sh[gy][cx] = s;
  __syncthreads();
  if (gy == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][cx];
    out[c] = accumulate ? out[c] + t : t;
  }
}

