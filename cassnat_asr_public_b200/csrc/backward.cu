// Backward-pass kernels of the training step: what loss.backward() of the reference
// (src/tasks/cassnat_task.py:291) runs through ATen autograd for the blocks on the path.
//   * batched strided fp32 GEMM (dX = dY W, dW = dY^T X, and the five contractions of attention backward)
//   * LayerNorm backward (unbiased std, eps on std: src/models/modules/norm.py:15-18)
//   * bias / activation backward with deterministic column sums
//   * softmax backward of masked attention with the reference's finfo.min semantics
//   * label-smoothing KL loss forward + gradient at the logits (src/utils/loss.py:21-35)
//   * transposing cast to bf16 (operands of the tensor-core backward GEMMs)
// All reductions are order-deterministic (no floating-point atomics).
#include "common.cuh"

namespace {

// ===================================================================== batched strided GEMM
// C[z][m,n] = alpha * sum_k A[z](m,k) * B[z](n,k) + beta * C[z][m,n];  z = (z1, z2) with separate strides.
constexpr int SBM = 64, SBN = 64, SBK = 16;

struct StridedGemm {
  const float* A; const float* B; float* C;
  int64_t a_z1, a_z2, a_m, a_k;
  int64_t b_z1, b_z2, b_n, b_k;
  int64_t c_z1, c_z2, c_m;
  int nz2, M, N, K;
  float alpha, beta;
};

__device__ __forceinline__ void strided_tile(const float* __restrict__ G, int64_t s_row, int64_t s_k, int row0, int nrows,
                                             int k0, int K, float (*S)[SBM + 1], int tid) {
#pragma unroll
  for (int i = 0; i < (SBM * SBK) / 256; ++i) {
    const int e = tid + 256 * i;
    int r, k;
    if (s_k == 1) { k = e % SBK; r = e / SBK; } else { r = e % SBM; k = e / SBM; }
    const int gr = row0 + r, gk = k0 + k;
    S[k][r] = (gr < nrows && gk < K) ? G[(int64_t)gr * s_row + (int64_t)gk * s_k] : 0.f;
  }
}

__global__ void __launch_bounds__(256) gemm_f32_strided_kernel(const StridedGemm g) {
  __shared__ float As[SBK][SBM + 1];
  __shared__ float Bs[SBK][SBN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z1 = blockIdx.z / g.nz2, z2 = blockIdx.z % g.nz2;
  const float* A = g.A + z1 * g.a_z1 + z2 * g.a_z2;
  const float* B = g.B + z1 * g.b_z1 + z2 * g.b_z2;
  float* C = g.C + z1 * g.c_z1 + z2 * g.c_z2;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < g.K; k0 += SBK) {
    strided_tile(A, g.a_m, g.a_k, m0, g.M, k0, g.K, As, tid);
    strided_tile(B, g.b_n, g.b_k, n0, g.N, k0, g.K, Bs, tid);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= g.N) continue;
      float* c = C + (int64_t)m * g.c_m + n;
      float v = g.alpha * acc[i][j];
      if (g.beta != 0.f) v += g.beta * *c;
      *c = v;
    }
  }
}

// ===================================================================== column sums (deterministic)
constexpr int CS_ROWS = 64;

// partial[blk][c] = sum over the block's rows; also the optional activation backward:
//   dpre[m,c] = alpha * dy[m,c] * act'(saved[m,c])     (ReLU: saved = output; Swish: saved = pre-activation)
__global__ void __launch_bounds__(256) act_bwd_colsum_kernel(const float* __restrict__ dy, int64_t lddy,
                                                             const float* __restrict__ saved, int64_t lds, int act,
                                                             float alpha, float* __restrict__ dpre, int64_t ldp,
                                                             int64_t M, int64_t N, float* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = (r0 + CS_ROWS < M) ? r0 + CS_ROWS : M;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    float v = alpha * dy[r * lddy + c];
    if (act == CN_ACT_RELU) {
      v = saved[r * lds + c] > 0.f ? v : 0.f;
    } else if (act == CN_ACT_SWISH) {
      const float z = saved[r * lds + c];
      const float sg = 1.f / (1.f + expf(-z));
      v *= sg * (1.f + z * (1.f - sg));
    }
    if (dpre) dpre[r * ldp + c] = v;
    s += v;
  }
  if (partial) partial[(int64_t)blockIdx.y * N + c] = s;
}

// The same with mixed element types: dy and saved may be bf16 (gradients / activations of bf16 GEMM outputs), dpre is
// written in bf16 -- the operand both backward GEMMs read -- so that the tensor-core backward of a linear layer needs
// no dtype copies around this kernel.
template <typename TD, typename TS>
__global__ void __launch_bounds__(256) act_bwd_colsum16_kernel(const TD* __restrict__ dy, int64_t lddy,
                                                               const TS* __restrict__ saved, int64_t lds, int act,
                                                               float alpha, __nv_bfloat16* __restrict__ dpre, int64_t ldp,
                                                               int64_t M, int64_t N, float* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = (r0 + CS_ROWS < M) ? r0 + CS_ROWS : M;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    float v = alpha * (float)dy[r * lddy + c];
    if (act == CN_ACT_RELU) {
      v = (float)saved[r * lds + c] > 0.f ? v : 0.f;
    } else if (act == CN_ACT_SWISH) {
      const float z = (float)saved[r * lds + c];
      const float sg = 1.f / (1.f + expf(-z));
      v *= sg * (1.f + z * (1.f - sg));
    }
    if (dpre) dpre[r * ldp + c] = __float2bfloat16(v);
    s += v;
  }
  if (partial) partial[(int64_t)blockIdx.y * N + c] = s;
}

// ===================================================================== LayerNorm backward
constexpr int LN_ROWS = 32;

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ dy, int64_t lddy,
                                                            const float* __restrict__ a, float eps, int64_t rows, int d,
                                                            float* __restrict__ dx, int64_t lddx,
                                                            float* __restrict__ pa, float* __restrict__ pb) {
  __shared__ float s_mean[LN_ROWS], s_r[LN_ROWS], s_c2[LN_ROWS], s_mg[LN_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * LN_ROWS;
  for (int i = warp; i < LN_ROWS; i += 8) {
    const int64_t r = r0 + i;
    if (r >= rows) continue;                      // warp-uniform
    const float* xr = x + r * ldx;
    const float* gr = dy + r * lddy;
    float sx = 0.f;
    for (int c = lane; c < d; c += 32) sx += xr[c];
    const float mean = cn_warp_sum(sx) / d;
    float ss = 0.f, sg = 0.f, sgc = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float cc = xr[c] - mean, g = gr[c] * a[c];
      ss += cc * cc; sg += g; sgc += g * cc;
    }
    ss = cn_warp_sum(ss); sg = cn_warp_sum(sg); sgc = cn_warp_sum(sgc);
    const float sd = sqrtf(ss / (d - 1));
    const float rr = 1.f / (sd + eps);
    if (lane == 0) {
      s_mean[i] = mean; s_r[i] = rr; s_mg[i] = sg / d;
      s_c2[i] = sd > 0.f ? sgc * rr * rr / ((d - 1) * sd) : 0.f;
    }
  }
  __syncthreads();
  const int nr = (int)((rows - r0 < LN_ROWS) ? rows - r0 : LN_ROWS);
  for (int c = threadIdx.x; c < d; c += 256) {
    const float ac = a[c];
    float da = 0.f, db = 0.f;
    for (int i = 0; i < nr; ++i) {
      const int64_t r = r0 + i;
      const float cc = x[r * ldx + c] - s_mean[i];
      const float g0 = dy[r * lddy + c];
      dx[r * lddx + c] = s_r[i] * (g0 * ac - s_mg[i]) - s_c2[i] * cc;
      da += g0 * cc * s_r[i];
      db += g0;
    }
    pa[(int64_t)blockIdx.x * d + c] = da;
    pb[(int64_t)blockIdx.x * d + c] = db;
  }
}

// ===================================================================== attention softmax backward
// In:  S  [Z,Tq,Tk] = scaled raw scores (before masking), dP [Z,Tq,Tk] = dO V^T.
// Out: S  <- P (probabilities exactly as the forward used them), dP <- dS = dL/d(scaled score).
// Plain variant: fully masked rows are uniform over all Tk keys and carry no score gradient
// (masked_fill makes them constants, attention.py:19-21).  zero_masked (relative variant,
// attention.py:138): masked probabilities are re-zeroed after the softmax.
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(float* __restrict__ S, float* __restrict__ dP,
                                                               const uint8_t* __restrict__ mask, int64_t mb, int64_t mi,
                                                               int h, int Tq, int Tk, int64_t rows_total, int zero_masked) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  const int i = (int)(row % Tq);
  const int64_t b = row / ((int64_t)Tq * h);
  float* s = S + row * Tk;
  float* g = dP ? dP + row * Tk : nullptr;                  // NULL: forward only (P in place)
  const uint8_t* mk = mask ? mask + b * mb + (int64_t)i * mi : nullptr;
  float mx = -INFINITY;
  int any = 0;
  for (int j = lane; j < Tk; j += 32) {
    const bool keep = !mk || mk[j];
    if (keep) { mx = fmaxf(mx, s[j]); any = 1; }
  }
  mx = cn_warp_max(mx);
  any = __any_sync(0xffffffffu, any);
  if (!any) {
    const float u = zero_masked ? 0.f : 1.f / Tk;
    for (int j = lane; j < Tk; j += 32) { s[j] = u; if (g) g[j] = 0.f; }
    return;
  }
  float sum = 0.f;
  for (int j = lane; j < Tk; j += 32) {
    const bool keep = !mk || mk[j];
    const float e = keep ? expf(s[j] - mx) : 0.f;
    s[j] = e;
    sum += e;
  }
  const float inv = 1.f / cn_warp_sum(sum);
  float dot = 0.f;
  for (int j = lane; j < Tk; j += 32) {
    const float p = s[j] * inv;
    s[j] = p;
    if (g) dot += p * g[j];
  }
  if (!g) return;
  dot = cn_warp_sum(dot);
  for (int j = lane; j < Tk; j += 32) g[j] = s[j] * (g[j] - dot);
}

// ===================================================================== dropout (training)
// nn.Dropout of the reference (modules/utils.py:32 residual branches, attention.py:22-23 attention probabilities,
// positionff.py:16 FFN hidden, embedding.py:31,60 positional encodings).  The keep mask is a pure function of
// (seed, site, element index) -- a counter-based hash, no state -- so the backward pass regenerates it instead of
// storing it, the seed can live in device memory (a CUDA-graph-replayed step advances it without re-capture), and a
// test can obtain the very mask the kernels used.
__device__ __forceinline__ uint32_t cn_fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ bool cn_dropout_keep(uint64_t seed, uint32_t site, uint64_t idx, uint32_t threshold) {
  uint32_t h = cn_fmix32((uint32_t)idx * 0x9E3779B1u ^ (uint32_t)seed);
  h = cn_fmix32(h ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u + site * 0xC2B2AE3Du + (uint32_t)(seed >> 32)));
  return h >= threshold;                                   // P(keep) = 1 - threshold / 2^32
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dropout_kernel(const TI* __restrict__ x, TO* __restrict__ out, int64_t n, uint32_t threshold,
                                                      float inv_keep, const int64_t* __restrict__ seed_dev, uint64_t seed,
                                                      uint32_t site, const float* __restrict__ res = nullptr) {
  const uint64_t sd = seed_dev ? (uint64_t)seed_dev[0] : seed;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = cn_dropout_keep(sd, site, (uint64_t)i, threshold) ? (float)x[i] * inv_keep : 0.f;
    if (res) v += res[i];                                   // residual + scale * dropout(branch) in one pass
    out[i] = (TO)v;
  }
}

// ===================================================================== label smoothing loss
// logp [R,V] (log-softmax rows), target i64 [R]; rows with target == pad are skipped.
// row_loss[r] = KL(smoothed one-hot || p) of row r (0 for pad rows); dlogits = (p - w) * keep * gscale
__global__ void __launch_bounds__(256) label_smoothing_kernel(const float* __restrict__ logp, int64_t ld, int64_t V,
                                                              const int64_t* __restrict__ target, int64_t pad,
                                                              float conf, float eps_v, float konst,
                                                              const float* __restrict__ gscale_dev, float gscale,
                                                              float* __restrict__ row_loss, float* __restrict__ dlogits,
                                                              int64_t ldg) {
  __shared__ float red[8];
  const int64_t r = blockIdx.x;
  const int64_t t = target[r];
  const bool keep = t != pad;
  const float* lp = logp + r * ld;
  const float gs = gscale_dev ? gscale * gscale_dev[0] : gscale;
  float s = 0.f;
  if (keep || dlogits) {
    for (int64_t v = threadIdx.x; v < V; v += 256) {
      const float l = lp[v];
      s += l;
      if (dlogits) {
        const float w = (v == t) ? conf : eps_v;
        dlogits[r * ldg + v] = keep ? (expf(l) - w) * gs : 0.f;
      }
    }
  }
  if (!row_loss) return;
  s = cn_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float lt = keep ? lp[t] : 0.f;
    row_loss[r] = keep ? konst - conf * lt - eps_v * (tot - lt) : 0.f;
  }
}

// tokens = #kept rows; loss = sum(row_loss) / tokens  (single CTA, fixed order)
__global__ void __launch_bounds__(256) label_smoothing_reduce_kernel(const float* __restrict__ row_loss,
                                                                     const int64_t* __restrict__ target, int64_t pad,
                                                                     int64_t R, float* __restrict__ loss,
                                                                     float* __restrict__ inv_tokens) {
  __shared__ double sh[256];
  __shared__ int cnt[256];
  double s = 0.0;
  int n = 0;
  for (int64_t r = threadIdx.x; r < R; r += 256) { s += row_loss[r]; n += target[r] != pad; }
  sh[threadIdx.x] = s; cnt[threadIdx.x] = n;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[threadIdx.x] += sh[threadIdx.x + o]; cnt[threadIdx.x] += cnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float it = 1.f / (float)cnt[0];
    loss[0] = (float)(sh[0] * it);
    inv_tokens[0] = it;
  }
}

// ===================================================================== transposing cast
// out bf16 [C, ldo] <- in^T, in [R, C] (f32 or bf16) row stride ldi.
template <typename TIn>
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const TIn* __restrict__ in, int64_t ldi, int64_t R, int64_t C,
                                                             __nv_bfloat16* __restrict__ out, int64_t ldo) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < R && c < C) ? (float)in[r * ldi + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < C && r < R) out[c * ldo + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
  }
}

// ===================================================================== relative-position terms
// S[z,i,j] += scale * Tb[z,i,clamp(j-i,-m,m)+m]     (Tb = (q+v) P^T, attention.py:119-133)
__global__ void __launch_bounds__(256) rel_bias_add_kernel(float* __restrict__ S, const float* __restrict__ Tb, int Tq, int Tk,
                                                           int m, float scale, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const int j = (int)(e % Tk);
  const int64_t row = e / Tk;
  const int i = (int)(row % Tq);
  int r = j - i;
  r = r < -m ? -m : (r > m ? m : r);
  S[e] += scale * Tb[row * (2 * m + 1) + r + m];
}

// G[z,i,r] = sum_{j : clamp(j-i,-m,m)+m == r} dS[z,i,j]   (gradient reaching the tabulated bias)
__global__ void __launch_bounds__(256) rel_bucket_kernel(const float* __restrict__ dS, float* __restrict__ G, int Tq, int Tk,
                                                         int m, int64_t rows_total) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  const int i = (int)(row % Tq);
  const float* d = dS + row * Tk;
  float* g = G + row * (2 * m + 1);
  float lo = 0.f, hi = 0.f;
  for (int j = lane; j < Tk; j += 32) {
    const int r = j - i;
    if (r <= -m) lo += d[j];
    else if (r >= m) hi += d[j];
    else g[r + m] = d[j];
  }
  lo = cn_warp_sum(lo); hi = cn_warp_sum(hi);
  if (lane == 0) {
    if (m == 0) g[0] = lo + hi;            // every offset clamps to the single row
    else { g[0] = lo; g[2 * m] = hi; }
  }
  // offsets whose key index falls outside [0,Tk): no contribution
  for (int r = -m + 1 + lane; r < m; r += 32) {
    const int j = i + r;
    if (j < 0 || j >= Tk) g[r + m] = 0.f;
  }
}

// y = act(x) elementwise, f32 -> f32 or bf16 (training keeps the pre-activation for backward)
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, int act, float* __restrict__ y32,
                                                      __nv_bfloat16* __restrict__ y16, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float v = x[i];
    if (act == CN_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == CN_ACT_SWISH) v = v / (1.f + expf(-v));
    if (y32) y32[i] = v;
    if (y16) y16[i] = __float2bfloat16(v);
  }
}

}  // namespace

// ============================================================================ C ABI
extern "C" int cn_rel_bias_add(float* S, const float* Tb, int64_t Z, int64_t Tq, int64_t Tk, int64_t max_rel, float scale,
                               cn_stream_t stream) {
  if (!S || !Tb || Z < 0 || Tq < 0 || Tk <= 0 || max_rel < 0) return CN_ERR_BAD_ARG;
  const int64_t total = Z * Tq * Tk;
  if (total == 0) return CN_OK;
  if ((total + 255) / 256 > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  rel_bias_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(S, Tb, (int)Tq, (int)Tk, (int)max_rel,
                                                                                         scale, total);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_rel_bucket(const float* dS, float* G, int64_t Z, int64_t Tq, int64_t Tk, int64_t max_rel,
                             cn_stream_t stream) {
  if (!dS || !G || Z < 0 || Tq < 0 || Tk <= 0 || max_rel < 0) return CN_ERR_BAD_ARG;
  const int64_t rows = Z * Tq;
  if (rows == 0) return CN_OK;
  if ((rows + 7) / 8 > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  rel_bucket_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dS, G, (int)Tq, (int)Tk, (int)max_rel, rows);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_act_fwd(const float* x, int act, float* y_f32, void* y_bf16, int64_t n, cn_stream_t stream) {
  if (!x || (!y_f32 && !y_bf16) || n < 0) return CN_ERR_BAD_ARG;
  if (n == 0) return CN_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  act_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, act, y_f32, (__nv_bfloat16*)y_bf16, n);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_gemm_f32_strided(const float* A, int64_t a_z1, int64_t a_z2, int64_t a_m, int64_t a_k,
                                   const float* B, int64_t b_z1, int64_t b_z2, int64_t b_n, int64_t b_k,
                                   float* C, int64_t c_z1, int64_t c_z2, int64_t c_m,
                                   int64_t nz1, int64_t nz2, int64_t M, int64_t N, int64_t K,
                                   float alpha, float beta, cn_stream_t stream) {
  if (!A || !B || !C || nz1 < 0 || nz2 < 0 || M < 0 || N < 0 || K < 0) return CN_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff || nz1 * nz2 > 65535) return CN_ERR_UNSUPPORTED;
  if (nz1 * nz2 == 0 || M == 0 || N == 0) return CN_OK;
  const int64_t gy = (M + SBM - 1) / SBM;
  if (gy > 65535) return CN_ERR_UNSUPPORTED;
  StridedGemm g{A, B, C, a_z1, a_z2, a_m, a_k, b_z1, b_z2, b_n, b_k, c_z1, c_z2, c_m,
                (int)nz2, (int)M, (int)N, (int)K, alpha, beta};
  dim3 grid((unsigned)((N + SBN - 1) / SBN), (unsigned)gy, (unsigned)(nz1 * nz2));
  gemm_f32_strided_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int64_t cn_colsum_workspace_bytes(int64_t M, int64_t N) {
  if (M < 0 || N < 0) return -1;
  return ((M + CS_ROWS - 1) / CS_ROWS) * N * (int64_t)sizeof(float) + 16;
}

extern "C" int cn_act_bwd_colsum(const float* dy, int64_t lddy, const float* saved, int64_t lds, int act, float alpha,
                                 float* dpre, int64_t ldp, int64_t M, int64_t N, float* colsum, int accumulate,
                                 void* workspace, int64_t workspace_bytes, cn_stream_t stream) {
  if (!dy || M < 0 || N < 0 || (act != CN_ACT_NONE && !saved) || (!dpre && !colsum)) return CN_ERR_BAD_ARG;
  if (M == 0 || N == 0) return CN_OK;
  const int64_t nblk = (M + CS_ROWS - 1) / CS_ROWS;
  if (nblk > 65535) return CN_ERR_UNSUPPORTED;
  if (colsum && (!workspace || workspace_bytes < cn_colsum_workspace_bytes(M, N))) return CN_ERR_WORKSPACE;
  float* partial = colsum ? (float*)workspace : nullptr;
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)nblk);
  act_bwd_colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, lddy, saved, lds, act, alpha, dpre, ldp, M, N, partial);
  CN_CHECK_LAUNCH();
  if (colsum) {
    colsum_reduce_kernel<<<(unsigned)((N + 31) / 32), 256, 0, (cudaStream_t)stream>>>(partial, nblk, N, colsum, accumulate);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}

extern "C" int cn_act_bwd_colsum_bf16(const void* dy, int dy_is_bf16, int64_t lddy, const void* saved, int saved_is_bf16,
                                      int64_t lds, int act, float alpha, void* dpre_bf16, int64_t ldp, int64_t M, int64_t N,
                                      float* colsum, int accumulate, void* workspace, int64_t workspace_bytes,
                                      cn_stream_t stream) {
  if (!dy || M < 0 || N < 0 || (act != CN_ACT_NONE && !saved) || (!dpre_bf16 && !colsum)) return CN_ERR_BAD_ARG;
  if (M == 0 || N == 0) return CN_OK;
  const int64_t nblk = (M + CS_ROWS - 1) / CS_ROWS;
  if (nblk > 65535) return CN_ERR_UNSUPPORTED;
  if (colsum && (!workspace || workspace_bytes < cn_colsum_workspace_bytes(M, N))) return CN_ERR_WORKSPACE;
  float* partial = colsum ? (float*)workspace : nullptr;
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)nblk);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* dp = (__nv_bfloat16*)dpre_bf16;
  if (dy_is_bf16 && saved_is_bf16)
    act_bwd_colsum16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)saved, lds, act, alpha, dp, ldp, M, N, partial);
  else if (dy_is_bf16)
    act_bwd_colsum16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, lddy, (const float*)saved, lds, act, alpha, dp, ldp, M, N, partial);
  else if (saved_is_bf16)
    act_bwd_colsum16_kernel<<<grid, 256, 0, st>>>((const float*)dy, lddy, (const __nv_bfloat16*)saved, lds, act, alpha, dp, ldp, M, N, partial);
  else
    act_bwd_colsum16_kernel<<<grid, 256, 0, st>>>((const float*)dy, lddy, (const float*)saved, lds, act, alpha, dp, ldp, M, N, partial);
  CN_CHECK_LAUNCH();
  if (colsum) {
    colsum_reduce_kernel<<<(unsigned)((N + 31) / 32), 256, 0, st>>>(partial, nblk, N, colsum, accumulate);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}

extern "C" int64_t cn_layernorm_bwd_workspace_bytes(int64_t rows, int64_t d) {
  if (rows < 0 || d < 0) return -1;
  return 2 * ((rows + LN_ROWS - 1) / LN_ROWS) * d * (int64_t)sizeof(float) + 16;
}

extern "C" int cn_layernorm_bwd(const float* x, int64_t ldx, const float* dy, int64_t lddy, const float* a, float eps,
                                int64_t rows, int64_t d, float* dx, int64_t lddx, float* da, float* db,
                                void* workspace, int64_t workspace_bytes, cn_stream_t stream) {
  if (!x || !dy || !a || !dx || !da || !db || rows < 0 || d < 2) return CN_ERR_BAD_ARG;
  if (d > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < cn_layernorm_bwd_workspace_bytes(rows, d)) return CN_ERR_WORKSPACE;
  const int64_t nblk = (rows + LN_ROWS - 1) / LN_ROWS;
  float* pa = (float*)workspace;
  float* pb = pa + nblk * d;
  if (rows > 0) {
    layernorm_bwd_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(x, ldx, dy, lddy, a, eps, rows, (int)d, dx, lddx, pa, pb);
    CN_CHECK_LAUNCH();
  }
  colsum_reduce_kernel<<<dim3((unsigned)((d + 31) / 32), 2), 256, 0, (cudaStream_t)stream>>>(pa, nblk, d, da, 0, pb, db);   // both sums, one launch
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_dropout(const void* x, int x_is_bf16, void* out, int out_is_bf16, int64_t n, float p, const int64_t* seed_dev,
                          int64_t seed, int64_t site, cn_stream_t stream) {
  if (!x || !out || n < 0 || !(p >= 0.f) || p >= 1.f) return CN_ERR_BAD_ARG;
  if (n == 0) return CN_OK;
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  const float inv_keep = 1.f / (1.f - p);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = (unsigned)blocks;
  if (x_is_bf16 && out_is_bf16)
    dropout_kernel<<<g, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, n, thr, inv_keep, seed_dev, (uint64_t)seed, (uint32_t)site);
  else if (x_is_bf16)
    dropout_kernel<<<g, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)out, n, thr, inv_keep, seed_dev, (uint64_t)seed, (uint32_t)site);
  else if (out_is_bf16)
    dropout_kernel<<<g, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)out, n, thr, inv_keep, seed_dev, (uint64_t)seed, (uint32_t)site);
  else
    dropout_kernel<<<g, 256, 0, st>>>((const float*)x, (float*)out, n, thr, inv_keep, seed_dev, (uint64_t)seed, (uint32_t)site);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_dropout_add(const void* x, int x_is_bf16, const float* residual, float* out, int64_t n, float p, float scale,
                              const int64_t* seed_dev, int64_t seed, int64_t site, cn_stream_t stream) {
  if (!x || !out || n < 0 || !(p >= 0.f) || p >= 1.f) return CN_ERR_BAD_ARG;
  if (n == 0) return CN_OK;
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  const float inv_keep = scale / (1.f - p);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_is_bf16)
    dropout_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, out, n, thr, inv_keep, seed_dev, (uint64_t)seed, (uint32_t)site, residual);
  else
    dropout_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float*)x, out, n, thr, inv_keep, seed_dev, (uint64_t)seed, (uint32_t)site, residual);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_attn_softmax_bwd(float* S, float* dP, const uint8_t* mask, int64_t mb, int64_t mi, int64_t B, int64_t h,
                                   int64_t Tq, int64_t Tk, int zero_masked, cn_stream_t stream) {
  if (!S || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  const int64_t rows = B * h * Tq;
  if (rows == 0) return CN_OK;
  if ((rows + 7) / 8 > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  attn_softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, dP, mask, mb, mi, (int)h, (int)Tq,
                                                                                         (int)Tk, rows, zero_masked);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_label_smoothing(const float* logp, int64_t ld, int64_t R, int64_t V, const int64_t* target, int64_t pad,
                                  float smoothing, float grad_scale, float* loss, float* dlogits, int64_t ldg,
                                  void* workspace, int64_t workspace_bytes, cn_stream_t stream) {
  if (!logp || !target || !loss || R <= 0 || V < 2 || smoothing < 0.f || smoothing >= 1.f) return CN_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < (R + 4) * (int64_t)sizeof(float)) return CN_ERR_WORKSPACE;
  if (R > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  float* row_loss = (float*)workspace;
  float* inv_tokens = row_loss + R;
  const float conf = 1.f - smoothing, eps_v = smoothing / (float)(V - 1);
  const float konst = (conf > 0.f ? conf * logf(conf) : 0.f) + (eps_v > 0.f ? (float)(V - 1) * eps_v * logf(eps_v) : 0.f);
  cudaStream_t st = (cudaStream_t)stream;
  // pass 1: row losses (no gradient yet: the 1/tokens factor comes out of the reduce)
  label_smoothing_kernel<<<(unsigned)R, 256, 0, st>>>(logp, ld, V, target, pad, conf, eps_v, konst, nullptr, 0.f, row_loss,
                                                      nullptr, 0);
  CN_CHECK_LAUNCH();
  label_smoothing_reduce_kernel<<<1, 256, 0, st>>>(row_loss, target, pad, R, loss, inv_tokens);
  CN_CHECK_LAUNCH();
  if (dlogits) {
    label_smoothing_kernel<<<(unsigned)R, 256, 0, st>>>(logp, ld, V, target, pad, conf, eps_v, konst, inv_tokens, grad_scale,
                                                        nullptr, dlogits, ldg);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}

extern "C" int cn_transpose_bf16(const void* in, int in_is_bf16, int64_t ldi, int64_t R, int64_t C, void* out, int64_t ldo,
                                 cn_stream_t stream) {
  if (!in || !out || R < 0 || C < 0 || ldi < C || ldo < R) return CN_ERR_BAD_ARG;
  if (R == 0 || C == 0) return CN_OK;
  const int64_t gy = (R + 31) / 32;
  if (gy > 65535) {
    // rows beyond the grid.y limit: walk in slabs
    for (int64_t r0 = 0; r0 < R; r0 += 65535 * 32) {
      const int64_t n = (R - r0 < 65535 * 32) ? R - r0 : 65535 * 32;
      const char* src = (const char*)in + r0 * ldi * (in_is_bf16 ? 2 : 4);
      int st = cn_transpose_bf16(src, in_is_bf16, ldi, n, C, (__nv_bfloat16*)out + r0, ldo, stream);
      if (st != CN_OK) return st;
    }
    return CN_OK;
  }
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)gy);
  if (in_is_bf16)
    transpose_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, ldi, R, C,
                                                                                 (__nv_bfloat16*)out, ldo);
  else
    transpose_bf16_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)in, ldi, R, C, (__nv_bfloat16*)out, ldo);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
