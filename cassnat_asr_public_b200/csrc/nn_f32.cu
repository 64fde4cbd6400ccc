// fp32 ("parity mode") building blocks of the extractor / SAD / MAD / output heads:
// FFMA GEMM with fused bias / activation / scale / residual, the reference's LayerNorm,
// masked multi-head attention with the reference's finfo.min semantics (+ relative-position
// variant) and the segment-sparse extractor attention that never materialises the trigger mask.
#include "common.cuh"

// ===================================================================== GEMM (SIMT fp32)
// C[M,N] = residual + alpha * act(A[M,K] * W[N,K]^T + bias); 128x128x16 tiles, 8x8 per thread.
namespace {
constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CN_ACT_SWISH) return v / (1.f + expf(-v));   // x * sigmoid(x)
  return v;
}

template <bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ G, int64_t ld, int64_t row0, int64_t nrows,
                                          int k0, int K, float (*S)[GBM + GPAD], int tid) {
  // tile: 128 rows x 16 k; thread -> (row = tid/4 + 64*i, k = (tid%4)*4 .. +3); stored transposed [k][row]
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (tid >> 2) + 64 * i;
    const int kq = (tid & 3) << 2;
    const int64_t gr = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < nrows) {
      const float* src = G + gr * ld + k0 + kq;
      if (VEC && k0 + kq + 3 < K) {
        v = *reinterpret_cast<const float4*>(src);
      } else {
        if (k0 + kq + 0 < K) v.x = src[0];
        if (k0 + kq + 1 < K) v.y = src[1];
        if (k0 + kq + 2 < K) v.z = src[2];
        if (k0 + kq + 3 < K) v.w = src[3];
      }
    }
    S[kq + 0][r] = v.x; S[kq + 1][r] = v.y; S[kq + 2][r] = v.z; S[kq + 3][r] = v.w;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ W, int64_t ldw,
                                                       const float* __restrict__ bias,
                                                       const float* __restrict__ residual, int64_t ldr,
                                                       int64_t res_rows, float alpha, int act,
                                                       float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int K) {
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Ws[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * GBM, n0 = (int64_t)blockIdx.x * GBN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + GBK - 1) / GBK;
  load_tile<VEC>(A, lda, m0, M, 0, K, As[0], tid);
  load_tile<VEC>(W, ldw, n0, N, 0, K, Ws[0], tid);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<VEC>(A, lda, m0, M, (kt + 1) * GBK, K, As[cur ^ 1], tid);
      load_tile<VEC>(W, ldw, n0, N, (kt + 1) * GBK, K, Ws[cur ^ 1], tid);
    }
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      // rows ty*4..+3 and 64+ty*4..+3 ; cols tx*4..+3 and 64+tx*4..+3 (keeps smem reads conflict-free)
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const int64_t rr = res_rows > 0 ? (m % res_rows) : m;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      v = apply_act(v, act);
      v *= alpha;
      if (residual) v = residual[rr * ldr + n] + v;
      C[m * ldc + n] = v;
    }
  }
}
}  // namespace

extern "C" int cn_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                           const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act, float* C,
                           int64_t ldc, int64_t M, int64_t N, int64_t K, cn_stream_t stream_) {
  if (!A || !W || !C || M < 0 || N < 0 || K <= 0 || lda < K || ldw < K || ldc < N) return CN_ERR_BAD_ARG;
  if (residual && ldr < N) return CN_ERR_BAD_ARG;
  if (K > 0x7ffffff0 || (N + GBN - 1) / GBN > 0x7fffffff || (M + GBM - 1) / GBM > 65535) return CN_ERR_UNSUPPORTED;
  if (M == 0 || N == 0) return CN_OK;
  dim3 grid((unsigned)((N + GBN - 1) / GBN), (unsigned)((M + GBM - 1) / GBM));
  const bool vec = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0 && lda % 4 == 0 &&
                   ldw % 4 == 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (vec)
    gemm_f32_kernel<true><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, residual, ldr, res_rows, alpha, act, C, ldc, M, N, (int)K);
  else
    gemm_f32_kernel<false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, residual, ldr, res_rows, alpha, act, C, ldc, M, N, (int)K);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// ===================================================================== LayerNorm
// modules/norm.py:15-18: a * (x - mean) / (std_unbiased + eps) + b ; one warp per row.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ldx,
                                                        const float* __restrict__ a, const float* __restrict__ b,
                                                        float eps, float* __restrict__ of32,
                                                        __nv_bfloat16* __restrict__ obf, int64_t ldo, int64_t ldob, int64_t lo_off) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* row = x + r * ldx;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += row[i];
  const float mean = cn_warp_sum(s) / (float)d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float c = row[i] - mean;
    q = fmaf(c, c, q);
  }
  const float var = cn_warp_sum(q) / (float)(d - 1);
  const float denom = sqrtf(var) + eps;
  for (int i = lane; i < d; i += 32) {
    const float y = a[i] * (row[i] - mean) / denom + b[i];
    if (of32) of32[r * ldo + i] = y;
    if (obf) {
      const __nv_bfloat16 yh = __float2bfloat16(y);
      obf[r * ldob + i] = yh;
      if (lo_off) obf[r * ldob + lo_off + i] = __float2bfloat16(y - __bfloat162float(yh));   // split (bf16x3) operand
    }
  }
}

// Fast path: the row lives in registers (NV float4 per lane), 128-bit loads and stores.
// RCP (bf16 / split-bf16 output of the tensor-core modes): one reciprocal per row instead of a division per element -- the sixteen
// IEEE divisions of a 512-wide row made the kernel issue-bound (ncu: 67 % issue slots busy at 39 % of the DRAM
// peak); the fp32-output path (parity mode) keeps the reference's x / (std + eps) element by element.
template <int NV, bool RCP>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ x, int64_t rows, int64_t ldx,
                                                            const float* __restrict__ a, const float* __restrict__ b,
                                                            float eps, float* __restrict__ of32,
                                                            __nv_bfloat16* __restrict__ obf, int64_t ldo, int64_t ldob, int64_t lo_off) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  cn_pdl_wait();
  if (r >= rows) return;        // (an exiting CTA counts as having released its dependents)
  const float4* row = reinterpret_cast<const float4*>(x + r * ldx);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = cn_ld_stream_f32x4(row + lane + 32 * i);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = cn_warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q); q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
  }
  const float denom = sqrtf(cn_warp_sum(q) / (float)(D - 1)) + eps;
  const float rinv = 1.f / denom;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(a + c)), be = __ldg(reinterpret_cast<const float4*>(b + c));
    float4 y;
    if (RCP) {
      y.x = fmaf(ga.x * v[i].x, rinv, be.x); y.y = fmaf(ga.y * v[i].y, rinv, be.y);
      y.z = fmaf(ga.z * v[i].z, rinv, be.z); y.w = fmaf(ga.w * v[i].w, rinv, be.w);
    } else {
      y.x = ga.x * v[i].x / denom + be.x; y.y = ga.y * v[i].y / denom + be.y;
      y.z = ga.z * v[i].z / denom + be.z; y.w = ga.w * v[i].w / denom + be.w;
    }
    if (of32) *reinterpret_cast<float4*>(of32 + r * ldo + c) = y;
    if (obf) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(obf + r * ldob + c) = pk;
      if (lo_off) {                                         // split (bf16x3) operand: rounding residue in the lo half
        const float2 f0 = __bfloat1622float2(lo), f1 = __bfloat1622float2(hi);
        __nv_bfloat162 r0 = __floats2bfloat162_rn(y.x - f0.x, y.y - f0.y), r1 = __floats2bfloat162_rn(y.z - f1.x, y.w - f1.y);
        pk.x = *reinterpret_cast<uint32_t*>(&r0);
        pk.y = *reinterpret_cast<uint32_t*>(&r1);
        *reinterpret_cast<uint2*>(obf + r * ldob + lo_off + c) = pk;
      }
    }
  }
}

static int layernorm_impl(const float* x, int64_t rows, int64_t d, int64_t ldx, const float* a, const float* b,
                          float eps, float* out_f32, void* out_bf16, int64_t ldo, int64_t ldob, int64_t lo_off,
                          cn_stream_t stream_) {
  if (!x || !a || !b || rows < 0 || d < 2 || ldx < d || (!out_f32 && !out_bf16)) return CN_ERR_BAD_ARG;
  if ((out_f32 && ldo < d) || (out_bf16 && ldob < d + lo_off)) return CN_ERR_BAD_ARG;
  if (rows == 0) return CN_OK;
  const int64_t blocks = (rows + 7) / 8;
  cudaStream_t stream = (cudaStream_t)stream_;
  __nv_bfloat16* obf = (__nv_bfloat16*)out_bf16;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                         reinterpret_cast<uintptr_t>(out_f32)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0 && (ldx & 3) == 0 && (ldo & 3) == 0 && (ldob & 3) == 0 && (lo_off & 3) == 0;
  const bool rcp = !out_f32;                         // tensor-core modes (bf16 copy, or the split bf16x3 operand whose
                                                     // own resolution is 2^-17); fp32 output keeps the exact form
#define CN_LN_VEC(NV_)                                                                                                  \
  do {                                                                                                                  \
    if (rcp) cn_launch_pdl(layernorm_vec_kernel<NV_, true>, dim3((unsigned)blocks), dim3(256), 0, stream, x, rows, ldx, a, b, eps, out_f32, obf, ldo, ldob, lo_off);  \
    else cn_launch_pdl(layernorm_vec_kernel<NV_, false>, dim3((unsigned)blocks), dim3(256), 0, stream, x, rows, ldx, a, b, eps, out_f32, obf, ldo, ldob, lo_off);    \
  } while (0)
  if (aligned && d == 128) CN_LN_VEC(1);
  else if (aligned && d == 256) CN_LN_VEC(2);
  else if (aligned && d == 512) CN_LN_VEC(4);
  else if (aligned && d == 1024) CN_LN_VEC(8);
  else layernorm_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, rows, (int)d, ldx, a, b, eps, out_f32, obf, ldo, ldob, lo_off);
#undef CN_LN_VEC
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_layernorm(const float* x, int64_t rows, int64_t d, int64_t ldx, const float* a, const float* b,
                            float eps, float* out_f32, void* out_bf16, int64_t ldo, cn_stream_t stream_) {
  return layernorm_impl(x, rows, d, ldx, a, b, eps, out_f32, out_bf16, ldo, ldo, 0, stream_);
}

// LayerNorm emitting the split bf16 operand [rows, 2d] = [hi | lo] of the bf16x3 GEMM (see cn_gemm_bf16x3)
extern "C" int cn_layernorm_bf16x2(const float* x, int64_t rows, int64_t d, int64_t ldx, const float* a, const float* b,
                                   float eps, float* out_f32, int64_t ldo_f32, void* out_split, int64_t ldo,
                                   cn_stream_t stream_) {
  if (!out_split) return CN_ERR_BAD_ARG;
  return layernorm_impl(x, rows, d, ldx, a, b, eps, out_f32, out_split, ldo_f32, ldo, d, stream_);
}

// ===================================================================== masked MHA (fp32)
// One CTA per (64-query tile, head, utterance); keys streamed in tiles of 64 with an online
// softmax.  Masked scores are finfo.min, NOT -inf: a fully masked row therefore ends up uniform
// over all Tk keys exactly like modules/attention.py:19-21.
namespace {
constexpr int AT = 64;   // tile edge (queries, keys) and head dim

struct AttnArgs {
  const float *Q, *K, *V;
  int64_t ldq, sbq, ldk, sbk, ldv, sbv;
  const uint8_t* mask; int64_t mb, mi;
  const float *rel_pos, *rel_u, *rel_v; int max_rel;
  int h, Tq, Tk;
  float* out; int64_t ldo, sbo;
};

template <bool REL>
__global__ void __launch_bounds__(256) attention_f32_kernel(AttnArgs p) {
  extern __shared__ __align__(16) float sm[];
  float(*Qt)[AT + 4] = reinterpret_cast<float(*)[AT + 4]>(sm);                       // [d][i]  (q or q+u)
  float(*Kt)[AT + 4] = reinterpret_cast<float(*)[AT + 4]>(sm + AT * (AT + 4));       // [d][j]
  float(*Vs)[AT + 4] = reinterpret_cast<float(*)[AT + 4]>(sm + 2 * AT * (AT + 4));   // [j][d]
  float(*Pt)[AT + 4] = reinterpret_cast<float(*)[AT + 4]>(sm + 3 * AT * (AT + 4));   // [j][i]
  float* qp = sm + 4 * AT * (AT + 4);                                                // REL: [i][2m+1]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * AT, hh = blockIdx.y, b = blockIdx.z;
  const int nrel = 2 * p.max_rel + 1;
  const float* Qb = p.Q + (int64_t)b * p.sbq + (int64_t)hh * AT;
  const float* Kb = p.K + (int64_t)b * p.sbk + (int64_t)hh * AT;
  const float* Vb = p.V + (int64_t)b * p.sbv + (int64_t)hh * AT;

  // ---- stage the query tile (transposed); REL: also (q+v).P for the 2m+1 clamped offsets
  for (int idx = tid; idx < AT * AT; idx += 256) {
    const int i = idx >> 6, d = idx & 63;
    float v = (q0 + i < p.Tq) ? Qb[(int64_t)(q0 + i) * p.ldq + d] : 0.f;
    if (REL) {
      Pt[d][i] = v + p.rel_v[hh * AT + d];       // scratch use of Pt: q + v
      v += p.rel_u[hh * AT + d];
    }
    Qt[d][i] = v;
  }
  __syncthreads();
  if (REL) {
    for (int idx = tid; idx < AT * nrel; idx += 256) {
      const int i = idx / nrel, r = idx - i * nrel;
      const float* pr = p.rel_pos + (int64_t)r * p.h * AT + hh * AT;
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < AT; ++d) s = fmaf(Pt[d][i], pr[d], s);
      qp[i * nrel + r] = s;
    }
    __syncthreads();
  }

  float o[4][4], m_run[4], l_run[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < p.Tk; k0 += AT) {
    for (int idx = tid; idx < AT * AT; idx += 256) {
      const int j = idx >> 6, d = idx & 63;
      const bool ok = k0 + j < p.Tk;
      Kt[d][j] = ok ? Kb[(int64_t)(k0 + j) * p.ldk + d] : 0.f;
      Vs[j][d] = ok ? Vb[(int64_t)(k0 + j) * p.ldv + d] : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < AT; ++d) {
      const float4 qa = *reinterpret_cast<const float4*>(&Qt[d][ty * 4]);
      const float4 kb = *reinterpret_cast<const float4*>(&Kt[d][tx * 4]);
      const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
    bool keep[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx * 4 + j;
        float v = s[i][j];
        if (REL) {
          int r = kj - qi;
          r = r < -p.max_rel ? -p.max_rel : (r > p.max_rel ? p.max_rel : r);
          v += qp[(ty * 4 + i) * nrel + r + p.max_rel];
        }
        v *= 0.125f;                                             // 1/sqrt(64), exact
        bool kp = true;
        if (p.mask && qi < p.Tq && kj < p.Tk) kp = p.mask[(int64_t)b * p.mb + (int64_t)qi * p.mi + kj] != 0;
        if (!kp) v = CN_F32_MIN;
        if (kj >= p.Tk) v = -INFINITY;
        keep[i][j] = kp;
        s[i][j] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = expf(m_run[i] - m_new);
      m_run[i] = m_new;
      l_run[i] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pv = expf(s[i][j] - m_new);
        l_run[i] += pv;                          // the normaliser counts masked entries (uniform rows)
        if (REL && !keep[i][j]) pv = 0.f;        // attention.py:138 re-zeroes masked probabilities
        Pt[tx * 4 + j][ty * 4 + i] = pv;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < AT; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(&Pt[j][ty * 4]);
      const float4 vb = *reinterpret_cast<const float4*>(&Vs[j][tx * 4]);
      const float pv[4] = {pa.x, pa.y, pa.z, pa.w}, vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float l = l_run[i];
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    const int qi = q0 + ty * 4 + i;
    if (qi < p.Tq) {
      float* dst = p.out + (int64_t)b * p.sbo + (int64_t)qi * p.ldo + hh * AT + tx * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c] = o[i][c] / l;
    }
  }
}
}  // namespace

extern "C" int cn_attention_f32(const float* Q, int64_t ldq, int64_t sbq, const float* K, int64_t ldk, int64_t sbk,
                                const float* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb,
                                int64_t mi, const float* rel_pos, const float* rel_bias_u,
                                const float* rel_bias_v, int64_t max_rel, int64_t B, int64_t h, int64_t Tq,
                                int64_t Tk, int64_t dk, float* out, int64_t ldo, int64_t sbo, cn_stream_t stream_) {
  if (!Q || !K || !V || !out || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (dk != AT || h > 65535 || B > 65535) return CN_ERR_UNSUPPORTED;
  const bool rel = rel_pos != nullptr;
  if (rel && (!rel_bias_u || !rel_bias_v || max_rel < 0 || max_rel > 64)) return CN_ERR_BAD_ARG;
  if (B == 0 || Tq == 0) return CN_OK;
  AttnArgs p;
  p.Q = Q; p.K = K; p.V = V; p.ldq = ldq; p.sbq = sbq; p.ldk = ldk; p.sbk = sbk; p.ldv = ldv; p.sbv = sbv;
  p.mask = mask; p.mb = mb; p.mi = mi; p.rel_pos = rel_pos; p.rel_u = rel_bias_u; p.rel_v = rel_bias_v;
  p.max_rel = (int)max_rel; p.h = (int)h; p.Tq = (int)Tq; p.Tk = (int)Tk; p.out = out; p.ldo = ldo; p.sbo = sbo;
  dim3 grid((unsigned)((Tq + AT - 1) / AT), (unsigned)h, (unsigned)B);
  size_t smem = (size_t)4 * AT * (AT + 4) * sizeof(float) + (rel ? (size_t)AT * (2 * max_rel + 1) * sizeof(float) : 0);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rel) {
    if (cudaFuncSetAttribute(attention_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CN_ERR_LAUNCH;
    attention_f32_kernel<true><<<grid, 256, smem, stream>>>(p);
  } else {
    if (cudaFuncSetAttribute(attention_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CN_ERR_LAUNCH;
    attention_f32_kernel<false><<<grid, 256, smem, stream>>>(p);
  }
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// ===================================================================== segment-sparse extractor attention
namespace {
__global__ void __launch_bounds__(256) column_mean_kernel(const float* __restrict__ V, int64_t ldv, int64_t sbv, int T,
                                                          int d, float* __restrict__ out) {
  // out[b, c] = mean_t V[b, t, c]; grid (ceil(d/256), B)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= d) return;
  const float* p = V + (int64_t)b * sbv + c;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += p[(int64_t)t * ldv];
  out[(int64_t)b * d + c] = s / (float)T;
}

struct SegArgs {
  const float *Q, *K, *V; int64_t ldq, ldk, sbk, ldv, sbv;
  const int32_t* seg; const int64_t* ylens; const uint8_t* src_mask; const int64_t* src_size;
  int left, right, h, U, T;
  float* out; int64_t ldo, sbo; const float* vmean;
  const int32_t* run_first; const int32_t* run_last;
};

// first / last frame of every segment id (run) of every utterance: one CTA per utterance
__global__ void __launch_bounds__(256) segment_bounds_kernel(const int32_t* __restrict__ seg, int T, int U,
                                                             int32_t* __restrict__ run_first,
                                                             int32_t* __restrict__ run_last) {
  const int b = blockIdx.x;
  const int32_t* s = seg + (int64_t)b * T;
  int32_t* f = run_first + (int64_t)b * U;
  int32_t* l = run_last + (int64_t)b * U;
  for (int u = threadIdx.x; u < U; u += blockDim.x) { f[u] = T; l[u] = -1; }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int v = s[t];
    if (v < 0 || v >= U) continue;
    if (t == 0 || s[t - 1] != v) f[v] = t;
    if (t == T - 1 || s[t + 1] != v) l[v] = t;
  }
}

// one warp per (token row, head); lanes own 2 of the 64 head dims
__global__ void __launch_bounds__(256) segment_attention_kernel(SegArgs p) {
  const int lane = threadIdx.x & 31, hh = threadIdx.x >> 5;
  const int u = blockIdx.x, b = blockIdx.y;
  if (hh >= p.h) return;
  const int T = p.T;
  const int32_t* seg = p.seg + (int64_t)b * T;
  const uint8_t* src = p.src_mask + (int64_t)b * T;
  const int ylens = (int)p.ylens[b];
  int fix_t = (int)p.src_size[b] - 1;
  if (fix_t < 0) fix_t += T;
  auto base = [&](int t) -> bool {
    if (t < 0 || t >= T) return false;
    return (seg[t] == u && src[t]) || (u == ylens && t == fix_t);
  };
  auto member = [&](int t) -> bool {
    bool m = base(t);
    if (p.right) m = m || base(t - 1);
    if (p.left) m = m || base(t + 1);
    return m && src[t];
  };
  // frames whose segment id equals u form one contiguous run (seg is non-decreasing); its bounds
  // were tabulated per utterance by segment_bounds_kernel (empty run: last < first)
  const int first = p.run_first[(int64_t)b * p.U + u], last = p.run_last[(int64_t)b * p.U + u];
  const float2 q = *reinterpret_cast<const float2*>(p.Q + (int64_t)u * p.ldq + hh * 64 + lane * 2);
  const float* Kb = p.K + (int64_t)b * p.sbk + hh * 64 + lane * 2;
  const float* Vb = p.V + (int64_t)b * p.sbv + hh * 64 + lane * 2;
  float m_run = -INFINITY, l_run = 0.f, o0 = 0.f, o1 = 0.f;
  int n_hit = 0;
  auto visit = [&](int t) {
    if (t < 0 || t >= T || !member(t)) return;
    const float2 k = *reinterpret_cast<const float2*>(Kb + (int64_t)t * p.ldk);
    float s = cn_warp_sum(fmaf(q.x, k.x, q.y * k.y)) * 0.125f;
    const float2 v = *reinterpret_cast<const float2*>(Vb + (int64_t)t * p.ldv);
    const float m_new = fmaxf(m_run, s);
    const float corr = expf(m_run - m_new), pv = expf(s - m_new);
    l_run = l_run * corr + pv;
    o0 = o0 * corr + pv * v.x;
    o1 = o1 * corr + pv * v.y;
    m_run = m_new;
    ++n_hit;
  };
  const int w0 = first - 1, w1 = last + 1;       // the run, dilated by one frame on both sides
  for (int t = w0; t <= w1; ++t) visit(t);
  if (u == ylens) {
    // eos fix-up cell and its dilation may lie outside the run's window
    for (int t = fix_t - 1; t <= fix_t + 1; ++t)
      if (t < w0 || t > w1) visit(t);
  }
  float* dst = p.out + (int64_t)b * p.sbo + (int64_t)u * p.ldo + hh * 64 + lane * 2;
  if (n_hit == 0) {
    // fully masked row: the reference's softmax over finfo.min scores is uniform over all T frames
    const float2 mv = *reinterpret_cast<const float2*>(p.vmean + ((int64_t)b * p.h + hh) * 64 + lane * 2);
    dst[0] = mv.x; dst[1] = mv.y;
  } else {
    dst[0] = o0 / l_run; dst[1] = o1 / l_run;
  }
}
}  // namespace

extern "C" int64_t cn_segment_attention_workspace_bytes(int64_t B, int64_t U, int64_t d) {
  return (((B * d + 63) / 64) * 64 + 2 * B * U) * 4;
}

extern "C" int cn_segment_attention_f32(const float* Q, int64_t ldq, const float* K, int64_t ldk, int64_t sbk,
                                        const float* V, int64_t ldv, int64_t sbv, const int32_t* seg,
                                        const int64_t* ylens, const uint8_t* src_mask, const int64_t* src_size,
                                        int left, int right, int64_t B, int64_t h, int64_t U, int64_t T, int64_t dk,
                                        float* out, int64_t ldo, int64_t sbo, void* workspace, int64_t workspace_bytes,
                                        cn_stream_t stream_) {
  if (!Q || !K || !V || !seg || !ylens || !src_mask || !src_size || !out || !workspace) return CN_ERR_BAD_ARG;
  if (workspace_bytes < cn_segment_attention_workspace_bytes(B, U, h * dk)) return CN_ERR_WORKSPACE;
  float* vmean_ws = (float*)workspace;
  int32_t* run_first = (int32_t*)(vmean_ws + ((B * h * dk + 63) / 64) * 64);
  int32_t* run_last = run_first + B * U;
  if (B < 0 || U < 0 || T <= 0 || h <= 0) return CN_ERR_BAD_ARG;
  if (dk != 64 || h > 8 || B > 65535) return CN_ERR_UNSUPPORTED;
  if (((ldq | ldk | ldv | ldo | sbk | sbv | sbo) & 1) != 0) return CN_ERR_UNSUPPORTED;
  if (B == 0 || U == 0) return CN_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int d = (int)(h * dk);
  column_mean_kernel<<<dim3((unsigned)((d + 255) / 256), (unsigned)B), 256, 0, stream>>>(V, ldv, sbv, (int)T, d, vmean_ws);
  CN_CHECK_LAUNCH();
  segment_bounds_kernel<<<(unsigned)B, 256, 0, stream>>>(seg, (int)T, (int)U, run_first, run_last);
  CN_CHECK_LAUNCH();
  SegArgs p;
  p.Q = Q; p.K = K; p.V = V; p.ldq = ldq; p.ldk = ldk; p.sbk = sbk; p.ldv = ldv; p.sbv = sbv;
  p.seg = seg; p.ylens = ylens; p.src_mask = src_mask; p.src_size = src_size;
  p.left = left > 0; p.right = right > 0; p.h = (int)h; p.U = (int)U; p.T = (int)T;
  p.out = out; p.ldo = ldo; p.sbo = sbo; p.vmean = vmean_ws; p.run_first = run_first; p.run_last = run_last;
  segment_attention_kernel<<<dim3((unsigned)U, (unsigned)B), (unsigned)(32 * h), 0, stream>>>(p);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
