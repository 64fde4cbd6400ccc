// Conformer convolution module core (modules/conformer_related.py:29-41), between the two
// pointwise convolutions (which are GEMMs): GLU over channels -> depthwise conv along the token
// axis -> GroupNorm(1, C) over the whole [C, L] slab of each utterance -> activation.
// Note the reference normalises over the PADDED length L of the batch (it passes no mask), so the
// statistics include padding rows; this kernel reproduces that.
#include "common.cuh"

namespace {
constexpr int TL = 32;     // token rows per CTA
constexpr int CC = 256;    // channels per pass (one per thread)

__device__ __forceinline__ float ld_as_f32(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_f32(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// pass 1: z = dwconv(glu(y)) + per-CTA partial (sum, sum of squares) in double
template <typename TIn>
__global__ void __launch_bounds__(CC) glu_dwconv_kernel(const TIn* __restrict__ y, int64_t ldy, int64_t sby,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        int L, int C, int ksz, float* __restrict__ z,
                                                        double* __restrict__ partial) {
  extern __shared__ float g[];                      // [(TL + ksz - 1)][CC]
  const int b = blockIdx.y, l0 = blockIdx.x * TL, pad = (ksz - 1) / 2;
  const int rows = TL + ksz - 1;
  float s1 = 0.f, s2 = 0.f;
  for (int c0 = 0; c0 < C; c0 += CC) {
    const int c = c0 + threadIdx.x;
    for (int r0 = 0; r0 < rows; r0 += 8) {          // 16 independent loads in flight per thread
      float a[8], gt[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int l = l0 - pad + r0 + k;
        const bool ok = c < C && r0 + k < rows && l >= 0 && l < L;
        const TIn* row = y + (int64_t)b * sby + (int64_t)(ok ? l : 0) * ldy;
        a[k] = ok ? ld_as_f32(row + c) : 0.f;
        gt[k] = ok ? ld_as_f32(row + C + c) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (r0 + k < rows) g[(r0 + k) * CC + threadIdx.x] = a[k] * (1.f / (1.f + __expf(-gt[k])));   // F.glu
    }
    // each thread only reads back its own column -> no barrier needed
    if (c < C) {
      const float bc = bias ? bias[c] : 0.f;
      for (int i = 0; i < TL && l0 + i < L; ++i) {
        float acc = bc;
        for (int k = 0; k < ksz; ++k) acc = fmaf(w[c * ksz + k], g[(i + k) * CC + threadIdx.x], acc);
        z[((int64_t)b * L + (l0 + i)) * C + c] = acc;
        s1 += acc;
        s2 = fmaf(acc, acc, s2);
      }
    }
  }
  __shared__ double red[2][CC / 32];
  double d1 = s1, d2 = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d1; red[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < CC / 32; ++i) { t1 += red[0][i]; t2 += red[1][i]; }
    partial[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 0] = t1;
    partial[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 1] = t2;
  }
}

// pass 1, two adjacent channels per thread (C even): 4-byte bf16x2 / 8-byte float2 loads, float2 stores
__device__ __forceinline__ float2 ld2_as_f32(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2_as_f32(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename TIn>
__global__ void __launch_bounds__(CC) glu_dwconv2_kernel(const TIn* __restrict__ y, int64_t ldy, int64_t sby,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         int L, int C, int ksz, float* __restrict__ z,
                                                         double* __restrict__ partial) {
  extern __shared__ float g[];                      // [(TL + ksz - 1)][2 * CC]
  const int b = blockIdx.y, l0 = blockIdx.x * TL, pad = (ksz - 1) / 2;
  const int rows = TL + ksz - 1;
  float s1 = 0.f, s2 = 0.f;
  for (int c0 = 0; c0 < C; c0 += 2 * CC) {
    const int c = c0 + 2 * threadIdx.x;
    for (int r0 = 0; r0 < rows; r0 += 8) {
      float2 a[8], gt[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int l = l0 - pad + r0 + k;
        const bool ok = c < C && r0 + k < rows && l >= 0 && l < L;
        const TIn* row = y + (int64_t)b * sby + (int64_t)(ok ? l : 0) * ldy;
        a[k] = ok ? ld2_as_f32(row + c) : make_float2(0.f, 0.f);
        gt[k] = ok ? ld2_as_f32(row + C + c) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (r0 + k < rows) {
          g[(r0 + k) * 2 * CC + 2 * threadIdx.x] = a[k].x * (1.f / (1.f + __expf(-gt[k].x)));
          g[(r0 + k) * 2 * CC + 2 * threadIdx.x + 1] = a[k].y * (1.f / (1.f + __expf(-gt[k].y)));
        }
    }
    if (c < C) {
      const float b0 = bias ? bias[c] : 0.f, b1 = bias ? bias[c + 1] : 0.f;
      for (int i = 0; i < TL && l0 + i < L; ++i) {
        float a0 = b0, a1 = b1;
        for (int k = 0; k < ksz; ++k) {
          a0 = fmaf(w[c * ksz + k], g[(i + k) * 2 * CC + 2 * threadIdx.x], a0);
          a1 = fmaf(w[(c + 1) * ksz + k], g[(i + k) * 2 * CC + 2 * threadIdx.x + 1], a1);
        }
        *reinterpret_cast<float2*>(z + ((int64_t)b * L + (l0 + i)) * C + c) = make_float2(a0, a1);
        s1 += a0 + a1;
        s2 = fmaf(a0, a0, fmaf(a1, a1, s2));
      }
    }
  }
  __shared__ double red[2][CC / 32];
  double d1 = s1, d2 = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d1; red[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < CC / 32; ++i) { t1 += red[0][i]; t2 += red[1][i]; }
    partial[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 0] = t1;
    partial[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 1] = t2;
  }
}

// pass 1, wide variant: time tiles of 128 rows (a 31-tap kernel's halo is then 23 % instead of the 94 % of a 32-row
// tile), 128 channels per CTA, one channel per thread in the convolution with the KSZ taps in registers and eight
// outputs per register tile (KSZ + 7 shared-memory reads for 8 KSZ FMAs; the generic kernels re-read every tap weight
// and every input per FMA).  Measured on the encoder's conv module (k = 31, [64 x 500 x 512]): 664 -> see DESIGN.
constexpr int CW = 128, TO = 8, WTH = 256;             // channels per CTA, outputs per register tile, threads
template <int KSZ> struct WideTile { static constexpr int TLW = KSZ >= 15 ? 128 : 64; };      // rows per CTA
// MODE 0: conv output to z (fp32) + partial sums.  When only the normalised bf16 result is wanted the conv output
// never goes to memory: MODE 1 computes the partial sums only, MODE 2 recomputes GLU + convolution from y (L2-hot: the
// producer GEMM's output was just read by MODE 1) and writes act(GroupNorm(.)) as bf16 -- 65 + 65 + 33 MB of traffic per
// call at [64 x 500 x 512] instead of 65 + 65 (z) + 65 (z) + 33.
struct WideApply { const float* stats; const float* gamma; const float* beta; int act; __nv_bfloat16* out16; };
template <typename TIn, int KSZ, int MODE>
__global__ void __launch_bounds__(WTH) glu_dwconv_wide_kernel(const TIn* __restrict__ y, int64_t ldy, int64_t sby,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              int L, int C, float* __restrict__ z, double* __restrict__ partial,
                                                              WideApply ap) {
  extern __shared__ float g[];                      // [TLW + KSZ - 1][CW]
  constexpr int TLW = WideTile<KSZ>::TLW, ROWS = TLW + KSZ - 1, PAD = (KSZ - 1) / 2;
  const int b = blockIdx.y, l0 = blockIdx.x * TLW, c0 = blockIdx.z * CW;
  // ---- GLU of the tile (+ halo): pairs of adjacent channels, eight rows in flight per thread
  for (int i0 = threadIdx.x; i0 < ROWS * (CW / 2); i0 += WTH * 8) {
    float2 a[8], gt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * WTH, r = i / (CW / 2), c = c0 + 2 * (i % (CW / 2)), l = l0 - PAD + r;
      const bool ok = i < ROWS * (CW / 2) && c < C && l >= 0 && l < L;
      const TIn* row = y + (int64_t)b * sby + (int64_t)(ok ? l : 0) * ldy;
      a[k] = ok ? ld2_as_f32(row + c) : make_float2(0.f, 0.f);
      gt[k] = ok ? ld2_as_f32(row + C + c) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * WTH;
      if (i < ROWS * (CW / 2)) {
        const int r = i / (CW / 2), cp = 2 * (i % (CW / 2));
        *reinterpret_cast<float2*>(g + r * CW + cp) =
            make_float2(__fdividef(a[k].x, 1.f + __expf(-gt[k].x)), __fdividef(a[k].y, 1.f + __expf(-gt[k].y)));
      }
    }
  }
  __syncthreads();
  // ---- depthwise convolution along time: thread = (channel, half of the tile's rows)
  const int ch = threadIdx.x % CW, half = threadIdx.x / CW, c = c0 + ch;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    float wr[KSZ];
#pragma unroll
    for (int k = 0; k < KSZ; ++k) wr[k] = w[c * KSZ + k];
    const float bz = bias ? bias[c] : 0.f;
    const int nout = min(TLW, L - l0);
    float mul = 1.f, add = 0.f;                      // MODE 2: GroupNorm(1,C) affine of this channel
    if (MODE == 2) {
      const float mu = ap.stats[2 * b], rstd = ap.stats[2 * b + 1];
      mul = rstd * ap.gamma[c];
      add = fmaf(-mu, mul, ap.beta[c]);
    }
    for (int i0 = half * (TLW / 2); i0 < min(nout, (half + 1) * (TLW / 2)); i0 += TO) {
      float in[TO + KSZ - 1];
#pragma unroll
      for (int k = 0; k < TO + KSZ - 1; ++k) in[k] = g[(i0 + k) * CW + ch];
#pragma unroll
      for (int o = 0; o < TO; ++o) {
        float acc = bz;
#pragma unroll
        for (int k = 0; k < KSZ; ++k) acc = fmaf(wr[k], in[o + k], acc);
        if (i0 + o < nout) {
          if (MODE == 0) z[((int64_t)b * L + (l0 + i0 + o)) * C + c] = acc;
          if (MODE == 2) {
            float v = fmaf(acc, mul, add);
            if (ap.act == CN_ACT_RELU) v = fmaxf(v, 0.f);
            else if (ap.act == CN_ACT_SWISH) v = __fdividef(v, 1.f + __expf(-v));
            ap.out16[((int64_t)b * L + (l0 + i0 + o)) * C + c] = __float2bfloat16(v);
          } else {
            s1 += acc;
            s2 = fmaf(acc, acc, s2);
          }
        }
      }
    }
  }
  if (MODE == 2) return;
  __shared__ double red[2][WTH / 32];
  double d1 = s1, d2 = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d1; red[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < WTH / 32; ++i) { t1 += red[0][i]; t2 += red[1][i]; }
    const int64_t blk = ((int64_t)b * gridDim.x + blockIdx.x) * gridDim.z + blockIdx.z;
    partial[blk * 2 + 0] = t1;
    partial[blk * 2 + 1] = t2;
  }
}
template <typename TIn, int KSZ, int MODE>
int launch_wide(const TIn* y, int64_t ldy, int64_t sby, const float* w, const float* bias, int L, int C, int64_t B, float* z,
                double* partial, WideApply ap, cudaStream_t stream) {
  constexpr int TLW = WideTile<KSZ>::TLW;
  constexpr int smem = (TLW + KSZ - 1) * CW * (int)sizeof(float);
  if (cudaFuncSetAttribute(glu_dwconv_wide_kernel<TIn, KSZ, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  dim3 grid((unsigned)((L + TLW - 1) / TLW), (unsigned)B, (unsigned)((C + CW - 1) / CW));
  glu_dwconv_wide_kernel<TIn, KSZ, MODE><<<grid, WTH, smem, stream>>>(y, ldy, sby, w, bias, L, C, z, partial, ap);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}
template <typename TIn, int MODE>
int dispatch_wide(int ksz, const TIn* y, int64_t ldy, int64_t sby, const float* w, const float* bias, int L, int C, int64_t B,
                  float* z, double* partial, WideApply ap, cudaStream_t stream) {
  switch (ksz) {
    case 3: return launch_wide<TIn, 3, MODE>(y, ldy, sby, w, bias, L, C, B, z, partial, ap, stream);
    case 7: return launch_wide<TIn, 7, MODE>(y, ldy, sby, w, bias, L, C, B, z, partial, ap, stream);
    case 15: return launch_wide<TIn, 15, MODE>(y, ldy, sby, w, bias, L, C, B, z, partial, ap, stream);
    case 31: return launch_wide<TIn, 31, MODE>(y, ldy, sby, w, bias, L, C, B, z, partial, ap, stream);
    default: return CN_ERR_UNSUPPORTED;
  }
}

// pass 1b: per-utterance mean / rstd from the double partials (one warp per utterance); keeps the fp64 divide and
// square root out of the elementwise pass
__global__ void __launch_bounds__(32) gn_finalize_kernel(const double* __restrict__ partial, int nblk, int L, int C, float eps,
                                                         float* __restrict__ stats) {
  const int b = blockIdx.x;
  double t1 = 0, t2 = 0;
  for (int i = threadIdx.x; i < nblk; i += 32) { t1 += partial[((int64_t)b * nblk + i) * 2]; t2 += partial[((int64_t)b * nblk + i) * 2 + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o); }
  if (threadIdx.x == 0) {
    const double n = (double)L * C;
    const double mean = t1 / n;
    double var = t2 / n - mean * mean;
    if (var < 0) var = 0;
    stats[2 * b] = (float)mean;
    stats[2 * b + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// pass 2: GroupNorm(1,C) affine + activation, 4 channels per thread; fp32 in place, or bf16 into out_bf16
__global__ void __launch_bounds__(256) groupnorm_act_kernel(float* __restrict__ z, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int act, int L, int C, __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.y;
  const float mu = stats[2 * b], rstd = stats[2 * b + 1];
  const int total4 = (L * C) >> 2, c4 = C >> 2;
  float4* zb = reinterpret_cast<float4*>(z + (int64_t)b * L * C);
  uint2* ob = out_bf16 ? reinterpret_cast<uint2*>(out_bf16 + (int64_t)b * L * C) : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const int c = (i % c4) << 2;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
    float4 v = zb[i];
    v.x = (v.x - mu) * rstd * g.x + bt.x; v.y = (v.y - mu) * rstd * g.y + bt.y;
    v.z = (v.z - mu) * rstd * g.z + bt.z; v.w = (v.w - mu) * rstd * g.w + bt.w;
    if (act == CN_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    else if (act == CN_ACT_SWISH) {
      v.x = v.x / (1.f + __expf(-v.x)); v.y = v.y / (1.f + __expf(-v.y));
      v.z = v.z / (1.f + __expf(-v.z)); v.w = v.w / (1.f + __expf(-v.w));
    }
    if (ob) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
      ob[i] = pk;
    } else {
      zb[i] = v;
    }
  }
}
// scalar tail version for C % 4 != 0
__global__ void __launch_bounds__(256) groupnorm_act_scalar_kernel(float* __restrict__ z, const float* __restrict__ stats,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   int act, int L, int C, __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.y;
  const float mu = stats[2 * b], rstd = stats[2 * b + 1];
  const int total = L * C;
  float* zb = z + (int64_t)b * total;
  __nv_bfloat16* ob = out_bf16 ? out_bf16 + (int64_t)b * total : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    float v = (zb[i] - mu) * rstd * gamma[c] + beta[c];
    if (act == CN_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == CN_ACT_SWISH) v = v / (1.f + __expf(-v));
    if (ob) ob[i] = __float2bfloat16(v);
    else zb[i] = v;
  }
}
// ------------------------------------------------------------------------------------- backward
// pass B2: dn = dout * act'(n), dwh = dn * gamma;  per-CTA partials of sum(dwh), sum(dwh * what) (double)
// and of the channel sums dgamma_c = sum dn * what, dbeta_c = sum dn.  z (pre-norm conv output, from pass 1)
// is overwritten with what, dwh goes to `dwh`.
__global__ void __launch_bounds__(CC) gn_bwd_stats_kernel(float* __restrict__ z, const double* __restrict__ partial, int nblk,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, int act, int L, int C, const float* __restrict__ dout,
                                                          float* __restrict__ dwh, double* __restrict__ s_partial,
                                                          float* __restrict__ pg, float* __restrict__ pb) {
  const int b = blockIdx.y, l0 = blockIdx.x * TL;
  double t1 = 0, t2 = 0;
  for (int i = 0; i < nblk; ++i) { t1 += partial[((int64_t)b * nblk + i) * 2]; t2 += partial[((int64_t)b * nblk + i) * 2 + 1]; }
  const double n = (double)L * C;
  const double mean = t1 / n;
  double var = t2 / n - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps)), mu = (float)mean;
  float s1 = 0.f, s2 = 0.f;
  const int64_t blk = (int64_t)b * gridDim.x + blockIdx.x;
  for (int c0 = 0; c0 < C; c0 += CC) {
    const int c = c0 + threadIdx.x;
    if (c >= C) continue;
    const float gm = gamma[c], bt = beta[c];
    float ag = 0.f, ab = 0.f;
    for (int i = 0; i < TL && l0 + i < L; ++i) {
      const int64_t o = ((int64_t)b * L + (l0 + i)) * C + c;
      const float wh = (z[o] - mu) * rstd;
      const float v = wh * gm + bt;
      float d = dout[o];
      if (act == CN_ACT_RELU) d = v > 0.f ? d : 0.f;
      else if (act == CN_ACT_SWISH) { const float sg = 1.f / (1.f + __expf(-v)); d *= sg * (1.f + v * (1.f - sg)); }
      const float dh = d * gm;
      z[o] = wh;
      dwh[o] = dh;
      s1 += dh; s2 = fmaf(dh, wh, s2);
      ag = fmaf(d, wh, ag); ab += d;
    }
    pg[blk * C + c] = ag;
    pb[blk * C + c] = ab;
  }
  __shared__ double red[2][CC / 32];
  double d1 = s1, d2 = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d1; red[1][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a1 = 0, a2 = 0;
    for (int i = 0; i < CC / 32; ++i) { a1 += red[0][i]; a2 += red[1][i]; }
    s_partial[blk * 2 + 0] = a1;
    s_partial[blk * 2 + 1] = a2;
  }
}

// pass B3: dw = rstd * (dwh - mean(dwh) - what * mean(dwh * what)); depthwise-conv backward (weight, bias, input)
// and GLU backward into dy [B,L,2C].
template <typename TIn>
__global__ void __launch_bounds__(CC) dwconv_glu_bwd_kernel(const TIn* __restrict__ y, int64_t ldy, int64_t sby,
                                                            const float* __restrict__ w, const float* __restrict__ what,
                                                            const float* __restrict__ dwh, const double* __restrict__ partial,
                                                            const double* __restrict__ s_partial, int nblk, float eps,
                                                            int L, int C, int ksz, float* __restrict__ dy,
                                                            float* __restrict__ pw, float* __restrict__ pbias) {
  extern __shared__ float sm[];                     // u tile then dw tile: 2 x [(TL + ksz - 1)][CC]
  const int b = blockIdx.y, l0 = blockIdx.x * TL, pad = (ksz - 1) / 2;
  const int rows = TL + ksz - 1;
  float* ut = sm;
  float* dt = sm + (size_t)rows * CC;
  double t1 = 0, t2 = 0, q1 = 0, q2 = 0;
  for (int i = 0; i < nblk; ++i) {
    t1 += partial[((int64_t)b * nblk + i) * 2]; t2 += partial[((int64_t)b * nblk + i) * 2 + 1];
    q1 += s_partial[((int64_t)b * nblk + i) * 2]; q2 += s_partial[((int64_t)b * nblk + i) * 2 + 1];
  }
  const double n = (double)L * C;
  const double mean = t1 / n;
  double var = t2 / n - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float m1 = (float)(q1 / n), m2 = (float)(q2 / n);
  const int64_t blk = (int64_t)b * gridDim.x + blockIdx.x;
  for (int c0 = 0; c0 < C; c0 += CC) {
    const int c = c0 + threadIdx.x;
    if (c >= C) continue;                           // each thread only touches its own smem column
    for (int r = 0; r < rows; ++r) {
      const int l = l0 - pad + r;
      const bool ok = l >= 0 && l < L;
      float u = 0.f, d = 0.f;
      if (ok) {
        const TIn* row = y + (int64_t)b * sby + (int64_t)l * ldy;
        u = ld_as_f32(row + c) * (1.f / (1.f + __expf(-ld_as_f32(row + C + c))));
        const int64_t o = ((int64_t)b * L + l) * C + c;
        d = rstd * (dwh[o] - m1 - what[o] * m2);
      }
      ut[r * CC + threadIdx.x] = u;
      dt[r * CC + threadIdx.x] = d;
    }
    const int own = (L - l0 < TL) ? L - l0 : TL;
    // weight / bias gradients over the CTA's own rows (tile row of token l0+i is i+pad)
    float bsum = 0.f;
    for (int i = 0; i < own; ++i) bsum += dt[(i + pad) * CC + threadIdx.x];
    pbias[blk * C + c] = bsum;
    for (int k = 0; k < ksz; ++k) {
      float acc = 0.f;
      for (int i = 0; i < own; ++i) acc = fmaf(dt[(i + pad) * CC + threadIdx.x], ut[(i + k) * CC + threadIdx.x], acc);
      pw[blk * (int64_t)C * ksz + (int64_t)c * ksz + k] = acc;
    }
    // input gradient: du[l] = sum_k dw[l - k + pad] * w[c,k]; then GLU backward
    for (int i = 0; i < own; ++i) {
      float du = 0.f;
      for (int k = 0; k < ksz; ++k) du = fmaf(dt[(i + 2 * pad - k) * CC + threadIdx.x], w[c * ksz + k], du);
      const int l = l0 + i;
      const TIn* row = y + (int64_t)b * sby + (int64_t)l * ldy;
      const float a = ld_as_f32(row + c);
      const float sg = 1.f / (1.f + __expf(-ld_as_f32(row + C + c)));
      float* drow = dy + ((int64_t)b * L + l) * 2 * C;
      drow[c] = du * sg;
      drow[C + c] = du * a * sg * (1.f - sg);
    }
  }
}
// dynamic shared memory ceilings for the largest supported kernel size (ksz <= 63); the attribute is a
// per-function maximum, so it is always set to the ceiling, never to the size of one call
bool set_smem_limits() {
  const int fwd = (TL + 62) * CC * (int)sizeof(float), bwd = 2 * fwd;
  const int fwd2 = (TL + 30) * 2 * CC * (int)sizeof(float);            // pair kernel: ksz <= 31
  return cudaFuncSetAttribute(glu_dwconv2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2) == cudaSuccess &&
         cudaFuncSetAttribute(glu_dwconv2_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2) == cudaSuccess &&
         cudaFuncSetAttribute(glu_dwconv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd) == cudaSuccess &&
         cudaFuncSetAttribute(glu_dwconv_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd) == cudaSuccess &&
         cudaFuncSetAttribute(dwconv_glu_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd) == cudaSuccess &&
         cudaFuncSetAttribute(dwconv_glu_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd) == cudaSuccess;
}
}  // namespace

extern "C" int64_t cn_conv_module_bwd_workspace_bytes(int64_t B, int64_t L, int64_t C, int64_t ksz) {
  if (B < 0 || L < 0 || C < 0 || ksz < 1) return -1;
  const int64_t nblk = B * ((L + TL - 1) / TL);
  return 2 * B * L * C * (int64_t)sizeof(float)          // what, dwh
         + nblk * 4 * (int64_t)sizeof(double)             // forward stats + backward stats partials
         + nblk * C * (2 + 1 + ksz) * (int64_t)sizeof(float) + 64;   // dgamma, dbeta, dbias, dweight partials
}

/* Backward of cn_glu_dwconv_groupnorm_act (recomputes the forward intermediates from y). */
extern "C" int cn_glu_dwconv_groupnorm_act_bwd(const void* y, int y_is_bf16, int64_t ldy, int64_t sby, const float* dw_weight,
                                               const float* dw_bias, const float* gamma, const float* beta, float eps,
                                               int act, int64_t B, int64_t L, int64_t C, int64_t ksz, const float* dout,
                                               float* dy, float* d_dw_weight, float* d_dw_bias, float* dgamma,
                                               float* dbeta, void* workspace, int64_t workspace_bytes, cn_stream_t stream_) {
  if (!y || !dw_weight || !gamma || !beta || !dout || !dy || !d_dw_weight || !d_dw_bias || !dgamma || !dbeta || !workspace ||
      B < 0 || L <= 0 || C <= 0)
    return CN_ERR_BAD_ARG;
  if (ksz < 1 || (ksz & 1) == 0 || ksz > 63 || ldy < 2 * C || B > 65535) return CN_ERR_UNSUPPORTED;
  if (workspace_bytes < cn_conv_module_bwd_workspace_bytes(B, L, C, ksz)) return CN_ERR_WORKSPACE;
  if (B == 0) return CN_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nblk = (int)((L + TL - 1) / TL);
  const int64_t nb = (int64_t)B * nblk;
  char* p = (char*)workspace;
  double* partial = (double*)p;            p += nb * 2 * sizeof(double);
  double* s_partial = (double*)p;          p += nb * 2 * sizeof(double);
  float* what = (float*)p;                 p += B * L * C * sizeof(float);
  float* dwh = (float*)p;                  p += B * L * C * sizeof(float);
  float* pg = (float*)p;                   p += nb * C * sizeof(float);
  float* pb = (float*)p;                   p += nb * C * sizeof(float);
  float* pbias = (float*)p;                p += nb * C * sizeof(float);
  float* pw = (float*)p;
  const size_t smem1 = (size_t)(TL + ksz - 1) * CC * sizeof(float);
  const size_t smem3 = 2 * smem1;
  if (!set_smem_limits()) return CN_ERR_LAUNCH;
  dim3 grid(nblk, (unsigned)B);
  if (y_is_bf16)
    glu_dwconv_kernel<__nv_bfloat16><<<grid, CC, smem1, stream>>>((const __nv_bfloat16*)y, ldy, sby, dw_weight, dw_bias, (int)L,
                                                                  (int)C, (int)ksz, what, partial);
  else
    glu_dwconv_kernel<float><<<grid, CC, smem1, stream>>>((const float*)y, ldy, sby, dw_weight, dw_bias, (int)L, (int)C,
                                                          (int)ksz, what, partial);
  CN_CHECK_LAUNCH();
  gn_bwd_stats_kernel<<<grid, CC, 0, stream>>>(what, partial, nblk, gamma, beta, eps, act, (int)L, (int)C, dout, dwh, s_partial,
                                               pg, pb);
  CN_CHECK_LAUNCH();
  if (y_is_bf16)
    dwconv_glu_bwd_kernel<__nv_bfloat16><<<grid, CC, smem3, stream>>>((const __nv_bfloat16*)y, ldy, sby, dw_weight, what, dwh,
                                                                      partial, s_partial, nblk, eps, (int)L, (int)C, (int)ksz,
                                                                      dy, pw, pbias);
  else
    dwconv_glu_bwd_kernel<float><<<grid, CC, smem3, stream>>>((const float*)y, ldy, sby, dw_weight, what, dwh, partial,
                                                              s_partial, nblk, eps, (int)L, (int)C, (int)ksz, dy, pw, pbias);
  CN_CHECK_LAUNCH();
  colsum_reduce_kernel<<<(unsigned)((C + 31) / 32), 256, 0, stream>>>(pg, nb, C, dgamma, 0);
  colsum_reduce_kernel<<<(unsigned)((C + 31) / 32), 256, 0, stream>>>(pb, nb, C, dbeta, 0);
  colsum_reduce_kernel<<<(unsigned)((C + 31) / 32), 256, 0, stream>>>(pbias, nb, C, d_dw_bias, 0);
  colsum_reduce_kernel<<<(unsigned)((C * ksz + 31) / 32), 256, 0, stream>>>(pw, nb, C * ksz, d_dw_weight, 0);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

constexpr int WS_PARTS = 4;
extern "C" int64_t cn_conv_module_workspace_bytes(int64_t B, int64_t L) {
  // partial sums: WS_PARTS (sum, sum of squares) pairs per 32-row block (the wide kernels keep one per 64- or 128-row
  // tile and 128-channel block: fits up to 1024 channels), then mean / rstd per utterance
  return B * ((L + TL - 1) / TL) * WS_PARTS * 2 * (int64_t)sizeof(double) + B * 2 * (int64_t)sizeof(float) + 64;
}

extern "C" int cn_glu_dwconv_groupnorm_act(const void* y, int y_is_bf16, int64_t ldy, int64_t sby,
                                           const float* dw_weight, const float* dw_bias, const float* gamma,
                                           const float* beta, float eps, int act, int64_t B, int64_t L, int64_t C,
                                           int64_t ksz, float* out, void* out_bf16, void* workspace,
                                           int64_t workspace_bytes, cn_stream_t stream_) {
  if (!y || !dw_weight || !gamma || !beta || !out || !workspace || B < 0 || L <= 0 || C <= 0) return CN_ERR_BAD_ARG;
  if (ksz < 1 || (ksz & 1) == 0 || ksz > 63 || ldy < 2 * C || B > 65535) return CN_ERR_UNSUPPORTED;
  if (workspace_bytes < cn_conv_module_workspace_bytes(B, L)) return CN_ERR_WORKSPACE;
  if (B == 0) return CN_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nblk = (int)((L + TL - 1) / TL);
  const size_t smem = (size_t)(TL + ksz - 1) * CC * sizeof(float);
  if (!set_smem_limits()) return CN_ERR_LAUNCH;
  const bool pair = (C & 1) == 0 && (ldy & 1) == 0 && (sby & 1) == 0 && ksz <= 31 &&
                    (reinterpret_cast<uintptr_t>(y) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
  // wide variant (128-row time tiles, taps in registers): its (time tile, channel block) partial sums must fit the
  // workspace, which is sized for one partial per 32 rows
  const int tlw = ksz >= 15 ? 128 : 64;
  const int nblk_w = (int)((L + tlw - 1) / tlw) * (int)((C + CW - 1) / CW);
  const bool wide = pair && (ksz == 3 || ksz == 7 || ksz == 15 || ksz == 31) && nblk_w <= nblk * WS_PARTS;
  // fast mode (bf16 in, only the bf16 result wanted): statistics pass + recompute-and-apply pass, no fp32 conv output
  const bool recompute = wide && y_is_bf16 && out_bf16 != nullptr && ksz <= 7;      // (k = 31: 151 us vs 136 us with the fp32 store)
  const WideApply none = {nullptr, nullptr, nullptr, 0, nullptr};
  if (recompute) {
    const int rc = dispatch_wide<__nv_bfloat16, 1>((int)ksz, (const __nv_bfloat16*)y, ldy, sby, dw_weight, dw_bias, (int)L, (int)C, B,
                                                   out, (double*)workspace, none, stream);
    if (rc != CN_OK) return rc;
  } else if (wide) {
    const int rc = y_is_bf16 ? dispatch_wide<__nv_bfloat16, 0>((int)ksz, (const __nv_bfloat16*)y, ldy, sby, dw_weight, dw_bias, (int)L,
                                                               (int)C, B, out, (double*)workspace, none, stream)
                             : dispatch_wide<float, 0>((int)ksz, (const float*)y, ldy, sby, dw_weight, dw_bias, (int)L, (int)C, B, out,
                                                       (double*)workspace, none, stream);
    if (rc != CN_OK) return rc;
  } else if (pair && y_is_bf16)
    glu_dwconv2_kernel<__nv_bfloat16><<<dim3(nblk, (unsigned)B), CC, 2 * smem, stream>>>(
        (const __nv_bfloat16*)y, ldy, sby, dw_weight, dw_bias, (int)L, (int)C, (int)ksz, out, (double*)workspace);
  else if (pair)
    glu_dwconv2_kernel<float><<<dim3(nblk, (unsigned)B), CC, 2 * smem, stream>>>((const float*)y, ldy, sby, dw_weight, dw_bias,
                                                                               (int)L, (int)C, (int)ksz, out, (double*)workspace);
  else if (y_is_bf16)
    glu_dwconv_kernel<__nv_bfloat16><<<dim3(nblk, (unsigned)B), CC, smem, stream>>>(
        (const __nv_bfloat16*)y, ldy, sby, dw_weight, dw_bias, (int)L, (int)C, (int)ksz, out, (double*)workspace);
  else
    glu_dwconv_kernel<float><<<dim3(nblk, (unsigned)B), CC, smem, stream>>>((const float*)y, ldy, sby, dw_weight, dw_bias,
                                                                          (int)L, (int)C, (int)ksz, out, (double*)workspace);
  CN_CHECK_LAUNCH();
  if (L * C > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  float* stats = reinterpret_cast<float*>((char*)workspace + (int64_t)B * nblk * WS_PARTS * 2 * sizeof(double));
  gn_finalize_kernel<<<(unsigned)B, 32, 0, stream>>>((const double*)workspace, wide ? nblk_w : nblk, (int)L, (int)C, eps, stats);
  CN_CHECK_LAUNCH();
  if (recompute) {
    const WideApply ap = {stats, gamma, beta, act, (__nv_bfloat16*)out_bf16};
    return dispatch_wide<__nv_bfloat16, 2>((int)ksz, (const __nv_bfloat16*)y, ldy, sby, dw_weight, dw_bias, (int)L, (int)C, B, out,
                                           nullptr, ap, stream);
  }
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0;
  const int64_t work = vec ? (L * C) / 4 : L * C;
  int gx = (int)((work + 255) / 256);
  if (gx > 148 * 2) gx = 148 * 2;
  if (vec)
    groupnorm_act_kernel<<<dim3(gx, (unsigned)B), 256, 0, stream>>>(out, stats, gamma, beta, act, (int)L, (int)C,
                                                                    (__nv_bfloat16*)out_bf16);
  else
    groupnorm_act_scalar_kernel<<<dim3(gx, (unsigned)B), 256, 0, stream>>>(out, stats, gamma, beta, act, (int)L, (int)C,
                                                                           (__nv_bfloat16*)out_bf16);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
