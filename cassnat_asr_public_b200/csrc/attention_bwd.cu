// Backward of masked multi-head attention for the decoder side of the training step (few query rows per utterance:
// the label tokens), fused into ONE launch per attention: recompute the probabilities, dP = dO V^T, softmax
// backward, dV = P~^T dO, dK = dS^T Q / sqrt(dk), dQ = dS K / sqrt(dk) -- the composition the reference's autograd
// runs for attention() of src/models/modules/attention.py:13-24 (incl. dropout on the probabilities, :22-23).
//
// One CTA per (utterance, head), all Tq <= 128 query rows resident in shared memory (Q, dO, dQ accumulator), the
// keys walked in tiles of 64 twice: pass 1 gives every row its softmax statistics, pass 2 recomputes the tile's
// probabilities, forms dS with D_i = dO_i . O_i (= sum_j P_ij dP_ij, also under dropout) and produces the tile's
// dK / dV rows and the dQ contribution.  FFMA on shared-memory tiles: at 41 x 250 per (utterance, head) the five
// products are 6 MFLOP -- what the composed path paid for was 5 strided-GEMM launches, a softmax launch, the
// dropout launches and the dtype copies around them per attention, 14 attentions per step.
// Masking semantics = cn_attn_softmax_bwd (plain variant): masked keys weigh exactly 0 in a row that keeps any key;
// a row without kept keys is uniform over all Tk keys and carries no score gradient.
#include "common.cuh"

namespace {

constexpr int KT = 64, DK = 64, PITCH = 68, TQ_MAX = 128, NTH = 256;      // PITCH: 16-byte rows, conflict-free float4 reads

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// the keep decision of cn_dropout (backward.cu) for element idx of dropout site `site`
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t site, uint64_t idx, uint32_t threshold) {
  uint32_t h = fmix32((uint32_t)idx * 0x9E3779B1u ^ (uint32_t)seed);
  h = fmix32(h ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u + site * 0xC2B2AE3Du + (uint32_t)(seed >> 32)));
  return h >= threshold;
}

struct BwdArgs {
  const void *q, *k, *v, *o, *go;      // [B,T,h*64] views (row pitch ld*, batch pitch sb*, elements)
  int64_t ldq, sbq, ldk, sbk, ldv, sbv, ldo, sbo, ldg, sbg;
  void *dq, *dk, *dv;                  // gradients, same element type, own pitches
  int64_t lddq, sbdq, lddk, sbdk, lddv, sbdv;
  const uint8_t* mask; int64_t mb, mi;
  int h, Tq, Tk; bool vec;
  float scale, inv_keep; uint32_t threshold, site; const int64_t* seed_dev; uint64_t seed;
};

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

__device__ __forceinline__ void unpack8(const uint4& raw, const __nv_bfloat16*, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 v = __bfloat1622float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
// NR rows x 64 features of a [T, h*64] view -> smem tile (row pitch PITCH), zero beyond nvalid rows.  vec: rows are
// 16-byte aligned -> 16-byte loads, all of a thread's loads issued before the first is used (a scalar loop pays the
// L2 latency once per element: 14 tile loads x 16 elements per thread made the kernel latency-bound)
template <int NR, typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* src, int64_t ld, int nvalid, bool vec) {
  constexpr int EPV = 16 / (int)sizeof(T);        // elements per 16-byte vector
  constexpr int VPR = DK / EPV;                   // vectors per row
  constexpr int NIT = (NR * VPR + NTH - 1) / NTH;
  if (vec) {
    uint4 raw[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * NTH, r = i / VPR, c = i % VPR;
      raw[it] = (i < NR * VPR && r < nvalid) ? *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + c * EPV)
                                             : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * NTH, r = i / VPR, c = i % VPR;
      if (i >= NR * VPR) continue;
      float* d = dst + r * PITCH + c * EPV;
      if (sizeof(T) == 2) {
        float f[8];
        unpack8(raw[it], (const __nv_bfloat16*)nullptr, f);
        *reinterpret_cast<float4*>(d) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        *reinterpret_cast<uint4*>(d) = raw[it];
      }
    }
  } else {
    for (int i = threadIdx.x; i < NR * DK; i += NTH) {
      const int r = i >> 6, d = i & 63;
      dst[r * PITCH + d] = r < nvalid ? ldf(src + (int64_t)r * ld + d) : 0.f;
    }
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// C[i][j] = sum_d A[i][d] * Bm[j][d]   (i < 16*RA rows of A, 64 rows of Bm): thread (ty, tx) owns rows ty + 16a,
// columns tx + 16c; the contraction is read four features at a time (RA + 4 vector loads per 16 RA FMAs)
template <int RA>
__device__ __forceinline__ void prod_nt(const float* A, const float* Bm, float (&acc)[RA][4], int ty, int tx) {
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 2
  for (int d = 0; d < DK; d += 4) {
    float4 av[RA], bv[4];
#pragma unroll
    for (int a = 0; a < RA; ++a) av[a] = ld4(A + (ty + 16 * a) * PITCH + d);
#pragma unroll
    for (int c = 0; c < 4; ++c) bv[c] = ld4(Bm + (tx + 16 * c) * PITCH + d);
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[a][c] = fmaf(av[a].x, bv[c].x, acc[a][c]);
        acc[a][c] = fmaf(av[a].y, bv[c].y, acc[a][c]);
        acc[a][c] = fmaf(av[a].z, bv[c].z, acc[a][c]);
        acc[a][c] = fmaf(av[a].w, bv[c].w, acc[a][c]);
      }
  }
}

template <typename T, int RA>          // RA = padded query rows / 16
__global__ void __launch_bounds__(NTH) attn_bwd_small_kernel(BwdArgs p) {
  constexpr int TQP = 16 * RA;
  extern __shared__ float sm[];
  float* Qs = sm;                       // [TQP][PITCH]
  float* Gs = Qs + TQP * PITCH;         // dO
  float* dQs = Gs + TQP * PITCH;        // accumulator
  float* Ks = dQs + TQP * PITCH;        // [KT][PITCH]
  float* Vs = Ks + KT * PITCH;
  float* Ps = Vs + KT * PITCH;          // [TQP][PITCH]  scores -> dropped probabilities of the tile
  float* Ss = Ps + TQP * PITCH;         // [TQP][PITCH]  dP -> dS of the tile
  __shared__ float row_m[TQ_MAX], row_l[TQ_MAX], row_D[TQ_MAX];
  __shared__ int row_any[TQ_MAX];

  const int b = blockIdx.x / p.h, hh = blockIdx.x % p.h;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ty = tid >> 4, tx = tid & 15;
  const int Tq = p.Tq, Tk = p.Tk;
  const T* q = (const T*)p.q + (int64_t)b * p.sbq + hh * DK;
  const T* k = (const T*)p.k + (int64_t)b * p.sbk + hh * DK;
  const T* v = (const T*)p.v + (int64_t)b * p.sbv + hh * DK;
  const T* o = (const T*)p.o + (int64_t)b * p.sbo + hh * DK;
  const T* go = (const T*)p.go + (int64_t)b * p.sbg + hh * DK;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * p.mb : nullptr;
  const uint64_t seed = p.seed_dev ? (uint64_t)p.seed_dev[0] : p.seed;
  const bool drop = p.threshold != 0;

  load_tile<TQP>(Qs, q, p.ldq, Tq, p.vec);
  load_tile<TQP>(Gs, go, p.ldg, Tq, p.vec);
  for (int i = tid; i < TQP * PITCH; i += NTH) dQs[i] = 0.f;
  for (int i = tid; i < TQ_MAX; i += NTH) { row_m[i] = -INFINITY; row_l[i] = 0.f; row_any[i] = 0; row_D[i] = 0.f; }
  __syncthreads();
  // D_i = dO_i . O_i
  for (int i = warp; i < Tq; i += NTH / 32) {
    float s = Gs[i * PITCH + lane] * ldf(o + (int64_t)i * p.ldo + lane) +
              Gs[i * PITCH + lane + 32] * ldf(o + (int64_t)i * p.ldo + lane + 32);
    s = cn_warp_sum(s);
    if (lane == 0) row_D[i] = s;
  }

  const int n_tiles = (Tk + KT - 1) / KT;
  // ---- pass 1: softmax statistics of every query row (running maximum / sum over the kept keys)
  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    __syncthreads();
    load_tile<KT>(Ks, k + (int64_t)k0 * p.ldk, p.ldk, nk, p.vec);
    __syncthreads();
    float acc[RA][4];
    prod_nt<RA>(Qs, Ks, acc, ty, tx);
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) Ps[(ty + 16 * a) * PITCH + tx + 16 * c] = acc[a][c] * p.scale;
    __syncthreads();
    for (int i = warp; i < Tq; i += NTH / 32) {
      const uint8_t* mr = mk ? mk + (int64_t)i * p.mi : nullptr;
      float mx = -INFINITY;
      int any = 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = lane + 32 * c;
        if (j < nk && (!mr || mr[k0 + j])) { mx = fmaxf(mx, Ps[i * PITCH + j]); any = 1; }
      }
      mx = cn_warp_max(mx);
      any = __any_sync(0xffffffffu, any);
      if (!any) continue;
      const float m_old = row_m[i], m_new = fmaxf(m_old, mx);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = lane + 32 * c;
        if (j < nk && (!mr || mr[k0 + j])) s += expf(Ps[i * PITCH + j] - m_new);
      }
      s = cn_warp_sum(s);
      if (lane == 0) {
        row_l[i] = row_l[i] * (m_old == -INFINITY ? 0.f : expf(m_old - m_new)) + s;
        row_m[i] = m_new;
        row_any[i] = 1;
      }
    }
  }
  __syncthreads();

  // ---- pass 2: per key tile dV, dK rows and the dQ contribution
  const float inv_tk = 1.f / (float)Tk;
  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    __syncthreads();
    load_tile<KT>(Ks, k + (int64_t)k0 * p.ldk, p.ldk, nk, p.vec);
    load_tile<KT>(Vs, v + (int64_t)k0 * p.ldv, p.ldv, nk, p.vec);
    __syncthreads();
    {
      float acc[RA][4];
      prod_nt<RA>(Qs, Ks, acc, ty, tx);          // scores
      float acg[RA][4];
      prod_nt<RA>(Gs, Vs, acg, ty, tx);          // dP~ = dO V^T
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        const int i = ty + 16 * a;
        const uint8_t* mr = mk ? mk + (int64_t)i * p.mi : nullptr;
        const bool row_ok = i < Tq;
        const int any = row_ok ? row_any[i] : 0;
        const float m = row_m[i], inv_l = any ? 1.f / row_l[i] : 0.f, D = row_D[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = tx + 16 * c;
          float pr = 0.f, ds = 0.f;
          if (row_ok && j < nk) {
            const bool kept = !mr || mr[k0 + j];
            float dp = acg[a][c];
            float keepf = 1.f;
            if (drop) {
              const uint64_t idx = (((uint64_t)(b * p.h + hh) * Tq + i) * Tk) + (uint64_t)(k0 + j);
              keepf = dropout_keep(seed, p.site, idx, p.threshold) ? p.inv_keep : 0.f;
            }
            if (any) {
              pr = kept ? expf(acc[a][c] * p.scale - m) * inv_l : 0.f;
              ds = pr * (dp * keepf - D);
            } else {
              pr = inv_tk;                       // fully masked row: uniform, constant -> no score gradient
            }
            pr *= keepf;
          }
          Ps[i * PITCH + j] = pr;
          Ss[i * PITCH + j] = ds * p.scale;
        }
      }
    }
    __syncthreads();
    // dV[j][d] = sum_i P~[i][j] dO[i][d];  dK[j][d] = sum_i dS[i][j] Q[i][d]      (keys 4ty..4ty+3, features 4tx..4tx+3)
    {
      float av[4][4], ak[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) av[c][e] = ak[c][e] = 0.f;
#pragma unroll 2
      for (int i = 0; i < Tq; ++i) {
        const float4 p4 = ld4(Ps + i * PITCH + 4 * ty), s4 = ld4(Ss + i * PITCH + 4 * ty);
        const float4 g4 = ld4(Gs + i * PITCH + 4 * tx), q4 = ld4(Qs + i * PITCH + 4 * tx);
        const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            av[c][e] = fmaf(pv[c], gv[e], av[c][e]);
            ak[c][e] = fmaf(sv[c], qv[e], ak[c][e]);
          }
      }
      T* dvp = (T*)p.dv + (int64_t)b * p.sbdv + hh * DK;
      T* dkp = (T*)p.dk + (int64_t)b * p.sbdk + hh * DK;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = 4 * ty + c;
        if (j >= nk) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          stf(dvp + (int64_t)(k0 + j) * p.lddv + 4 * tx + e, av[c][e]);
          stf(dkp + (int64_t)(k0 + j) * p.lddk + 4 * tx + e, ak[c][e]);
        }
      }
    }
    // dQ[i][d] += sum_j dS[i][j] K[j][d]      (rows ty + 16a, features 4tx..4tx+3; four keys per step)
    {
      float aq[RA][4];
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) aq[a][e] = 0.f;
#pragma unroll 2
      for (int j = 0; j < KT; j += 4) {
        float4 sv[RA], kv[4];
#pragma unroll
        for (int a = 0; a < RA; ++a) sv[a] = ld4(Ss + (ty + 16 * a) * PITCH + j);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) kv[jj] = ld4(Ks + (j + jj) * PITCH + 4 * tx);
#pragma unroll
        for (int a = 0; a < RA; ++a) {
          const float s_[4] = {sv[a].x, sv[a].y, sv[a].z, sv[a].w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            aq[a][0] = fmaf(s_[jj], kv[jj].x, aq[a][0]);
            aq[a][1] = fmaf(s_[jj], kv[jj].y, aq[a][1]);
            aq[a][2] = fmaf(s_[jj], kv[jj].z, aq[a][2]);
            aq[a][3] = fmaf(s_[jj], kv[jj].w, aq[a][3]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) dQs[(ty + 16 * a) * PITCH + 4 * tx + e] += aq[a][e];      // (own elements only)
    }
  }
  __syncthreads();
  T* dqp = (T*)p.dq + (int64_t)b * p.sbdq + hh * DK;
  for (int i = tid; i < Tq * DK; i += NTH) {
    const int r = i >> 6, d = i & 63;
    stf(dqp + (int64_t)r * p.lddq + d, dQs[r * PITCH + d]);
  }
}

// ---------------------------------------------------------------- forward with dropout on the probabilities
// The training forward of the same attentions (attention.py:22-23: p_attn = dropout(softmax(scores))): one launch,
// online softmax over the key tiles, the keep mask regenerated from (seed, site, element index) -- the index being
// that of the [B,h,Tq,Tk] probability tensor, so that the backward above sees the same mask.
struct FwdArgs {
  const void *q, *k, *v; void* out;
  int64_t ldq, sbq, ldk, sbk, ldv, sbv, ldo, sbo;
  const uint8_t* mask; int64_t mb, mi;
  int h, Tq, Tk; bool vec;
  float scale, inv_keep; uint32_t threshold, site; const int64_t* seed_dev; uint64_t seed;
};

template <typename T, int RA>
__global__ void __launch_bounds__(NTH) attn_fwd_small_kernel(FwdArgs p) {
  constexpr int TQP = 16 * RA;
  extern __shared__ float sm[];
  float* Qs = sm;                       // [TQP][PITCH]
  float* Ks = Qs + TQP * PITCH;         // [KT][PITCH]
  float* Vs = Ks + KT * PITCH;
  float* Ps = Vs + KT * PITCH;          // [TQP][PITCH]
  __shared__ int row_any[TQ_MAX];

  const int b = blockIdx.x / p.h, hh = blockIdx.x % p.h;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ty = tid >> 4, tx = tid & 15;
  const int Tq = p.Tq, Tk = p.Tk;
  const T* q = (const T*)p.q + (int64_t)b * p.sbq + hh * DK;
  const T* k = (const T*)p.k + (int64_t)b * p.sbk + hh * DK;
  const T* v = (const T*)p.v + (int64_t)b * p.sbv + hh * DK;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * p.mb : nullptr;
  const uint64_t seed = p.seed_dev ? (uint64_t)p.seed_dev[0] : p.seed;
  const bool drop = p.threshold != 0;

  load_tile<TQP>(Qs, q, p.ldq, Tq, p.vec);
  // rows without any kept key are uniform over all Tk keys (masked_fill(finfo.min) + softmax): found up front
  for (int i = warp; i < TQP; i += NTH / 32) {
    int any = 0;
    if (i < Tq) {
      if (!mk) any = 1;
      else {
        const uint8_t* mr = mk + (int64_t)i * p.mi;
        for (int j = lane; j < Tk; j += 32) any |= mr[j] != 0;
      }
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) row_any[i] = any;
  }
  float o[RA][4], m_run[RA], l_run[RA];
#pragma unroll
  for (int a = 0; a < RA; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) o[a][e] = 0.f;
  }
  const float inv_tk = 1.f / (float)Tk;
  const int n_tiles = (Tk + KT - 1) / KT;
  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    __syncthreads();
    load_tile<KT>(Ks, k + (int64_t)k0 * p.ldk, p.ldk, nk, p.vec);
    load_tile<KT>(Vs, v + (int64_t)k0 * p.ldv, p.ldv, nk, p.vec);
    __syncthreads();
    float acc[RA][4];
    prod_nt<RA>(Qs, Ks, acc, ty, tx);
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int i = ty + 16 * a;
      const uint8_t* mr = mk ? mk + (int64_t)i * p.mi : nullptr;
      const bool row_ok = i < Tq;
      const int any = row_any[i];
      bool kept[4];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = tx + 16 * c;
        kept[c] = row_ok && j < nk && (!mr || mr[k0 + j]);
        acc[a][c] *= p.scale;
        if (kept[c]) mx = fmaxf(mx, acc[a][c]);
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));      // the 16 tx lanes of a row
      const float m_new = fmaxf(m_run[a], mx);
      const float corr = (m_run[a] == -INFINITY) ? 0.f : expf(m_run[a] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = tx + 16 * c;
        float pr = kept[c] ? expf(acc[a][c] - m_new) : 0.f;
        sum += pr;
        if (!any) pr = (row_ok && j < nk) ? inv_tk : 0.f;
        if (drop && pr != 0.f) {
          const uint64_t idx = (((uint64_t)(b * p.h + hh) * Tq + i) * Tk) + (uint64_t)(k0 + j);
          pr *= dropout_keep(seed, p.site, idx, p.threshold) ? p.inv_keep : 0.f;
        }
        Ps[i * PITCH + j] = pr;
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      if (mx > -INFINITY) {
        l_run[a] = l_run[a] * corr + sum;
        m_run[a] = m_new;
#pragma unroll
        for (int e = 0; e < 4; ++e) o[a][e] *= corr;
      }
    }
    __syncthreads();
    // O[i][d] += sum_j P~[i][j] V[j][d]      (rows ty + 16a, features 4tx..4tx+3; four keys per step)
#pragma unroll 2
    for (int j = 0; j < KT; j += 4) {
      float4 pv[RA], vv[4];
#pragma unroll
      for (int a = 0; a < RA; ++a) pv[a] = ld4(Ps + (ty + 16 * a) * PITCH + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) vv[jj] = ld4(Vs + (j + jj) * PITCH + 4 * tx);
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        const float p_[4] = {pv[a].x, pv[a].y, pv[a].z, pv[a].w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          o[a][0] = fmaf(p_[jj], vv[jj].x, o[a][0]);
          o[a][1] = fmaf(p_[jj], vv[jj].y, o[a][1]);
          o[a][2] = fmaf(p_[jj], vv[jj].z, o[a][2]);
          o[a][3] = fmaf(p_[jj], vv[jj].w, o[a][3]);
        }
      }
    }
  }
  T* op = (T*)p.out + (int64_t)b * p.sbo + hh * DK;
#pragma unroll
  for (int a = 0; a < RA; ++a) {
    const int i = ty + 16 * a;
    if (i >= Tq) continue;
    const float inv = row_any[i] ? 1.f / l_run[a] : 1.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) stf(op + (int64_t)i * p.ldo + 4 * tx + e, o[a][e] * inv);
  }
}

template <typename T, int RA>
int launch_fwd(const FwdArgs& p, int64_t B, cudaStream_t st) {
  constexpr int TQP = 16 * RA;
  const size_t smem = (size_t)(2 * TQP + 2 * KT) * PITCH * sizeof(float);
  if (cudaFuncSetAttribute(attn_fwd_small_kernel<T, RA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  attn_fwd_small_kernel<T, RA><<<(unsigned)(B * p.h), NTH, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

template <typename T>
int dispatch_fwd(const FwdArgs& p, int64_t B, cudaStream_t st) {
  switch ((p.Tq + 15) / 16) {
    case 1: return launch_fwd<T, 1>(p, B, st);
    case 2: return launch_fwd<T, 2>(p, B, st);
    case 3: return launch_fwd<T, 3>(p, B, st);
    case 4: return launch_fwd<T, 4>(p, B, st);
    case 5: return launch_fwd<T, 5>(p, B, st);
    case 6: return launch_fwd<T, 6>(p, B, st);
    case 7: return launch_fwd<T, 7>(p, B, st);
    case 8: return launch_fwd<T, 8>(p, B, st);
    default: return CN_ERR_UNSUPPORTED;
  }
}

// ================================================================ bf16: the same two kernels on mma.sync
// For bf16 operands the five products run on the tensor cores (mma.sync.m16n8k16, fp32 accumulate; tcgen05's
// 128-row tiles do not fit 41-row problems): Q, dO, K, V stay bf16 in shared memory (row pitch 72 elements:
// conflict-free ldmatrix), P~ and dS are rounded to bf16 for the second-level products exactly as the forward
// attention kernel rounds P.  Warp w owns the w-th 8-column slice of every product.
constexpr int PB = 72;               // bf16 row pitch (144 B)
constexpr int PF = 68;               // f32 scratch row pitch

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// fragment addresses (bytes, shared space) for lane l
__device__ __forceinline__ uint32_t a_rowmajor(uint32_t base, int m0, int k0, int l) {          // X[m][k]
  return base + 2u * (uint32_t)((m0 + (l & 15)) * PB + k0 + ((l >> 4) << 3));
}
__device__ __forceinline__ uint32_t a_transposed(uint32_t base, int m0, int k0, int l) {        // X[k][m], use .trans
  return base + 2u * (uint32_t)((k0 + (l & 7) + ((l >> 4) << 3)) * PB + m0 + (((l >> 3) & 1) << 3));
}
__device__ __forceinline__ uint32_t b_nk(uint32_t base, int n0, int k0, int l) {                // X[n][k]
  return base + 2u * (uint32_t)((n0 + (l & 7)) * PB + k0 + (((l >> 3) & 1) << 3));
}
__device__ __forceinline__ uint32_t b_kn(uint32_t base, int n0, int k0, int l) {                // X[k][n], use .trans
  return base + 2u * (uint32_t)((k0 + (l & 15)) * PB + n0);
}

// NR rows x 64 bf16 of a [T, h*64] view -> smem (pitch PB), zero beyond nvalid rows; 16-byte copies
template <int NR>
__device__ __forceinline__ void load_tile_b(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int nvalid, bool vec) {
  constexpr int NIT = (NR * 8 + NTH - 1) / NTH;
  if (vec) {
    uint4 raw[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * NTH, r = i >> 3, c = i & 7;
      raw[it] = (i < NR * 8 && r < nvalid) ? *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + c * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * NTH, r = i >> 3, c = i & 7;
      if (i < NR * 8) *reinterpret_cast<uint4*>(dst + r * PB + c * 8) = raw[it];
    }
  } else {
    for (int i = threadIdx.x; i < NR * DK; i += NTH) {
      const int r = i >> 6, d = i & 63;
      dst[r * PB + d] = r < nvalid ? src[(int64_t)r * ld + d] : __float2bfloat16(0.f);
    }
  }
}

// the same as an asynchronous copy (cp.async, one commit group per call): the next key tile is fetched while the
// current one is being used.  Rows beyond nvalid are zero-filled with ordinary stores.
template <int NR>
__device__ __forceinline__ void load_tile_b_async(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int nvalid, bool vec) {
  if (vec) {
    for (int i = threadIdx.x; i < NR * 8; i += NTH) {
      const int r = i >> 3, c = i & 7;
      __nv_bfloat16* d = dst + r * PB + c * 8;
      if (r < nvalid) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d)),
                     "l"(src + (int64_t)r * ld + c * 8) : "memory");
      } else {
        *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  } else {
    load_tile_b<NR>(dst, src, ld, nvalid, false);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// scores of one key tile: acc[a] = fragment (rows 16a + g, +8; columns 8*warp + 2t, +1) of X Y^T, X [TQP][64], Y [64][64]
template <int RA>
__device__ __forceinline__ void mma_nt(uint32_t X, uint32_t Y, float (&acc)[RA][4], int warp, int lane) {
  uint32_t bf[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x2(bf[ks], b_nk(Y, 8 * warp, 16 * ks, lane));
#pragma unroll
  for (int a = 0; a < RA; ++a) {
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[a][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[4];
      ldsm_x4(af, a_rowmajor(X, 16 * a, 16 * ks, lane));
      mma16816(acc[a], af, bf[ks]);
    }
  }
}

// cn_dropout's keep decision for an element index below 2^32 (the high word contributes a per-launch constant c2)
__device__ __forceinline__ float keep_factor32(uint32_t seed_lo, uint32_t c2, uint32_t idx, uint32_t threshold, float inv_keep) {
  uint32_t h = fmix32(idx * 0x9E3779B1u ^ seed_lo);
  h = fmix32(h ^ c2);
  return h >= threshold ? inv_keep : 0.f;
}

constexpr int MW = 64;               // mask words per row (Tk <= 2048)
// keep-mask of the whole problem as bits in shared memory, built once: reading mask bytes from global memory inside
// the per-tile loops costs an L2 round trip per row and tile (measured: 3 us per key tile, 10x the arithmetic)
__device__ __forceinline__ void build_mask_bits(uint32_t* Mw, const uint8_t* mk, int64_t mi, int Tq, int Tk, int warp, int lane) {
  const int n_rows = mk ? (mi ? Tq : 1) : 0, n_words = (Tk + 31) / 32;
  for (int i = warp; i < n_rows; i += NTH / 32) {
    const uint8_t* mr = mk + (int64_t)i * mi;
    for (int w0 = 0; w0 < n_words; w0 += 8) {
      uint8_t byte[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = (w0 + u) * 32 + lane;
        byte[u] = j < Tk ? mr[j] : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t bits = __ballot_sync(0xffffffffu, byte[u] != 0);
        if (lane == 0 && w0 + u < n_words) Mw[i * MW + w0 + u] = bits;
      }
    }
  }
}
__device__ __forceinline__ bool mask_kept(const uint32_t* Mw, bool has_mask, bool per_row, int i, int key) {
  return !has_mask || ((Mw[(per_row ? i : 0) * MW + (key >> 5)] >> (key & 31)) & 1u);
}

// row statistics of the staged f32 score tile (pass 1 of both kernels), mask from the shared-memory bits
__device__ __forceinline__ void row_stats_update_bits(const float* Sf, const uint32_t* Mw, bool has_mask, bool per_row, int Tq,
                                                      int k0, int nk, float* row_m, float* row_l, int* row_any, int warp,
                                                      int lane) {
  for (int i = warp; i < Tq; i += NTH / 32) {
    float mx = -INFINITY;
    int any = 0;
    bool kept[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = lane + 32 * c;
      kept[c] = j < nk && mask_kept(Mw, has_mask, per_row, i, k0 + j);
      if (kept[c]) { mx = fmaxf(mx, Sf[i * PF + j]); any = 1; }
    }
    mx = cn_warp_max(mx);
    any = __any_sync(0xffffffffu, any);
    if (!any) continue;
    const float m_old = row_m[i], m_new = fmaxf(m_old, mx);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (kept[c]) s += __expf(Sf[i * PF + lane + 32 * c] - m_new);
    s = cn_warp_sum(s);
    if (lane == 0) {
      row_l[i] = row_l[i] * (m_old == -INFINITY ? 0.f : __expf(m_old - m_new)) + s;
      row_m[i] = m_new;
      row_any[i] = 1;
    }
  }
}

template <int RA>
__global__ void __launch_bounds__(NTH) attn_bwd_small_mma_kernel(BwdArgs p) {
  constexpr int TQP = 16 * RA;
  extern __shared__ __align__(16) unsigned char smraw[];
  __nv_bfloat16* Qb = reinterpret_cast<__nv_bfloat16*>(smraw);      // [TQP][PB]
  __nv_bfloat16* Gb = Qb + TQP * PB;                                 // dO
  __nv_bfloat16* Kb = Gb + TQP * PB;                                 // [2][KT][PB]  (double-buffered key tiles)
  __nv_bfloat16* Vb = Kb + 2 * KT * PB;                              // [2][KT][PB]
  __nv_bfloat16* Pb = Vb + 2 * KT * PB;                              // [TQP][PB] dropped probabilities of the tile
  __nv_bfloat16* Sb = Pb + TQP * PB;                                 // [TQP][PB] dS * scale
  float* Sf = reinterpret_cast<float*>(Sb + TQP * PB);               // [TQP][PF] scores of pass 1
  __shared__ float row_m[TQ_MAX], row_l[TQ_MAX], row_D[TQ_MAX];
  __shared__ int row_any[TQ_MAX];
  uint32_t* Mw = reinterpret_cast<uint32_t*>(Sf + TQP * PF);        // [TQP][MW] keep-mask bits
  typedef __nv_bfloat16 T;

  const int b = blockIdx.x / p.h, hh = blockIdx.x % p.h;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int Tq = p.Tq, Tk = p.Tk;
  const T* q = (const T*)p.q + (int64_t)b * p.sbq + hh * DK;
  const T* k = (const T*)p.k + (int64_t)b * p.sbk + hh * DK;
  const T* v = (const T*)p.v + (int64_t)b * p.sbv + hh * DK;
  const T* o = (const T*)p.o + (int64_t)b * p.sbo + hh * DK;
  const T* go = (const T*)p.go + (int64_t)b * p.sbg + hh * DK;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * p.mb : nullptr;
  const uint64_t seed = p.seed_dev ? (uint64_t)p.seed_dev[0] : p.seed;
  const bool drop = p.threshold != 0;
  const bool has_mask = mk != nullptr, per_row = p.mi != 0;
  const uint32_t seed_lo = (uint32_t)seed, c2 = p.site * 0xC2B2AE3Du + (uint32_t)(seed >> 32);
  const bool idx32 = (uint64_t)gridDim.x * (uint64_t)Tq * (uint64_t)Tk < 0xffffffffull;
  build_mask_bits(Mw, mk, p.mi, Tq, Tk, warp, lane);
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(Qb), sG = (uint32_t)__cvta_generic_to_shared(Gb),
                 sK0 = (uint32_t)__cvta_generic_to_shared(Kb), sV0 = (uint32_t)__cvta_generic_to_shared(Vb),
                 sP = (uint32_t)__cvta_generic_to_shared(Pb), sS = (uint32_t)__cvta_generic_to_shared(Sb);

  load_tile_b<TQP>(Qb, q, p.ldq, Tq, p.vec);
  load_tile_b<TQP>(Gb, go, p.ldg, Tq, p.vec);
  for (int i = tid; i < TQ_MAX; i += NTH) { row_m[i] = -INFINITY; row_l[i] = 0.f; row_any[i] = 0; row_D[i] = 0.f; }
  __syncthreads();
  for (int i = warp; i < Tq; i += NTH / 32) {           // D_i = dO_i . O_i
    float s = __bfloat162float(Gb[i * PB + lane]) * __bfloat162float(o[(int64_t)i * p.ldo + lane]) +
              __bfloat162float(Gb[i * PB + lane + 32]) * __bfloat162float(o[(int64_t)i * p.ldo + lane + 32]);
    s = cn_warp_sum(s);
    if (lane == 0) row_D[i] = s;
  }
  const int n_tiles = (Tk + KT - 1) / KT;
  auto fetch_k = [&](int t) {
    if (t < n_tiles) load_tile_b_async<KT>(Kb + (t & 1) * KT * PB, k + (int64_t)t * KT * p.ldk, p.ldk, min(KT, Tk - t * KT), p.vec);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto fetch_v = [&](int t) {
    if (t < n_tiles) load_tile_b_async<KT>(Vb + (t & 1) * KT * PB, v + (int64_t)t * KT * p.ldv, p.ldv, min(KT, Tk - t * KT), p.vec);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- pass 1: softmax statistics
  fetch_k(0);
  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    fetch_k(t + 1);                                  // (its buffer was last read two tiles ago, behind a barrier)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const uint32_t sK = sK0 + (uint32_t)((t & 1) * KT * PB * 2);
    float acc[RA][4];
    mma_nt<RA>(sQ, sK, acc, warp, lane);
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      Sf[(16 * a + g) * PF + 8 * warp + 2 * t4] = acc[a][0] * p.scale;
      Sf[(16 * a + g) * PF + 8 * warp + 2 * t4 + 1] = acc[a][1] * p.scale;
      Sf[(16 * a + g + 8) * PF + 8 * warp + 2 * t4] = acc[a][2] * p.scale;
      Sf[(16 * a + g + 8) * PF + 8 * warp + 2 * t4 + 1] = acc[a][3] * p.scale;
    }
    __syncthreads();
    row_stats_update_bits(Sf, Mw, has_mask, per_row, Tq, k0, nk, row_m, row_l, row_any, warp, lane);
    __syncthreads();
  }
  // ---- pass 2
  float dq[RA][4];
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[a][e] = 0.f;
  const float inv_tk = 1.f / (float)Tk;
  T* dvp = (T*)p.dv + (int64_t)b * p.sbdv + hh * DK;
  T* dkp = (T*)p.dk + (int64_t)b * p.sbdk + hh * DK;
  fetch_k(0);
  fetch_v(0);
  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    fetch_k(t + 1);
    fetch_v(t + 1);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    const uint32_t sK = sK0 + (uint32_t)((t & 1) * KT * PB * 2), sV = sV0 + (uint32_t)((t & 1) * KT * PB * 2);
    {
      float acc[RA][4], acg[RA][4];
      mma_nt<RA>(sQ, sK, acc, warp, lane);        // scores
      mma_nt<RA>(sG, sV, acg, warp, lane);        // dP~ = dO V^T
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {           // the fragment's two rows
          const int i = 16 * a + g + 8 * hr;
          const bool row_ok = i < Tq;
          const uint64_t rowbase = ((uint64_t)(b * p.h + hh) * Tq + i) * Tk + (uint64_t)(k0 + 8 * warp + 2 * t4);
          const uint32_t mword = has_mask ? Mw[(per_row ? i : 0) * MW + ((k0 + 8 * warp) >> 5)] >> ((k0 + 8 * warp + 2 * t4) & 31) : 3u;
          const int any = row_ok ? row_any[i] : 0;
          const float m = row_m[i], inv_l = any ? 1.f / row_l[i] : 0.f, D = row_D[i];
          float pr2[2], ds2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 8 * warp + 2 * t4 + e;
            float pr = 0.f, ds = 0.f;
            if (row_ok && j < nk) {
              const bool kept = (mword >> e) & 1u;
              float keepf = 1.f;
              if (drop) {
                keepf = idx32 ? keep_factor32(seed_lo, c2, (uint32_t)rowbase + e, p.threshold, p.inv_keep)
                              : (dropout_keep(seed, p.site, rowbase + e, p.threshold) ? p.inv_keep : 0.f);
              }
              if (any) {
                pr = kept ? __expf(acc[a][2 * hr + e] * p.scale - m) * inv_l : 0.f;
                ds = pr * (acg[a][2 * hr + e] * keepf - D);
              } else {
                pr = inv_tk;
              }
              pr *= keepf;
            }
            pr2[e] = pr;
            ds2[e] = ds * p.scale;
          }
          *reinterpret_cast<__nv_bfloat162*>(Pb + i * PB + 8 * warp + 2 * t4) = __floats2bfloat162_rn(pr2[0], pr2[1]);
          *reinterpret_cast<__nv_bfloat162*>(Sb + i * PB + 8 * warp + 2 * t4) = __floats2bfloat162_rn(ds2[0], ds2[1]);
        }
    }
    __syncthreads();
    // dV = P~^T dO, dK = dS^T Q: [64 keys x 64 features], contraction over the query rows; warp w: features 8w..8w+7
    {
      uint32_t bg[RA][2], bq[RA][2];
#pragma unroll
      for (int ks = 0; ks < RA; ++ks) {
        ldsm_x2_t(bg[ks], b_kn(sG, 8 * warp, 16 * ks, lane));
        ldsm_x2_t(bq[ks], b_kn(sQ, 8 * warp, 16 * ks, lane));
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        float cv[4] = {0.f, 0.f, 0.f, 0.f}, ck[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < RA; ++ks) {
          uint32_t ap[4], as[4];
          ldsm_x4_t(ap, a_transposed(sP, 16 * mt, 16 * ks, lane));
          ldsm_x4_t(as, a_transposed(sS, 16 * mt, 16 * ks, lane));
          mma16816(cv, ap, bg[ks]);
          mma16816(ck, as, bq[ks]);
        }
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int j = 16 * mt + g + 8 * hr;
          if (j < nk) {
            *reinterpret_cast<__nv_bfloat162*>(dvp + (int64_t)(k0 + j) * p.lddv + 8 * warp + 2 * t4) = __floats2bfloat162_rn(cv[2 * hr], cv[2 * hr + 1]);
            *reinterpret_cast<__nv_bfloat162*>(dkp + (int64_t)(k0 + j) * p.lddk + 8 * warp + 2 * t4) = __floats2bfloat162_rn(ck[2 * hr], ck[2 * hr + 1]);
          }
        }
      }
    }
    // dQ += dS K: [TQP x 64 features], contraction over the 64 keys
    {
      uint32_t bk[4][2];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x2_t(bk[ks], b_kn(sK, 8 * warp, 16 * ks, lane));
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t as[4];
          ldsm_x4(as, a_rowmajor(sS, 16 * a, 16 * ks, lane));
          mma16816(dq[a], as, bk[ks]);
        }
    }
    __syncthreads();
  }
  T* dqp = (T*)p.dq + (int64_t)b * p.sbdq + hh * DK;
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int i = 16 * a + g + 8 * hr;
      if (i < Tq)
        *reinterpret_cast<__nv_bfloat162*>(dqp + (int64_t)i * p.lddq + 8 * warp + 2 * t4) = __floats2bfloat162_rn(dq[a][2 * hr], dq[a][2 * hr + 1]);
    }
}

template <int RA>
__global__ void __launch_bounds__(NTH) attn_fwd_small_mma_kernel(FwdArgs p) {
  constexpr int TQP = 16 * RA;
  extern __shared__ __align__(16) unsigned char smraw[];
  __nv_bfloat16* Qb = reinterpret_cast<__nv_bfloat16*>(smraw);      // [TQP][PB]
  __nv_bfloat16* Kb = Qb + TQP * PB;                                 // [2][KT][PB]
  __nv_bfloat16* Vb = Kb + 2 * KT * PB;                              // [2][KT][PB]
  __nv_bfloat16* Pb = Vb + 2 * KT * PB;                              // [TQP][PB]
  float* Sf = reinterpret_cast<float*>(Pb + TQP * PB);               // [TQP][PF]
  __shared__ float row_m[TQ_MAX], row_l[TQ_MAX];
  __shared__ int row_any[TQ_MAX];
  uint32_t* Mw = reinterpret_cast<uint32_t*>(Sf + TQP * PF);        // [TQP][MW] keep-mask bits
  typedef __nv_bfloat16 T;

  const int b = blockIdx.x / p.h, hh = blockIdx.x % p.h;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int Tq = p.Tq, Tk = p.Tk;
  const T* q = (const T*)p.q + (int64_t)b * p.sbq + hh * DK;
  const T* k = (const T*)p.k + (int64_t)b * p.sbk + hh * DK;
  const T* v = (const T*)p.v + (int64_t)b * p.sbv + hh * DK;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * p.mb : nullptr;
  const uint64_t seed = p.seed_dev ? (uint64_t)p.seed_dev[0] : p.seed;
  const bool drop = p.threshold != 0;
  const bool has_mask = mk != nullptr, per_row = p.mi != 0;
  const uint32_t seed_lo = (uint32_t)seed, c2 = p.site * 0xC2B2AE3Du + (uint32_t)(seed >> 32);
  const bool idx32 = (uint64_t)gridDim.x * (uint64_t)Tq * (uint64_t)Tk < 0xffffffffull;
  build_mask_bits(Mw, mk, p.mi, Tq, Tk, warp, lane);
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(Qb), sK0 = (uint32_t)__cvta_generic_to_shared(Kb),
                 sV0 = (uint32_t)__cvta_generic_to_shared(Vb), sP = (uint32_t)__cvta_generic_to_shared(Pb);
  load_tile_b<TQP>(Qb, q, p.ldq, Tq, p.vec);
  for (int i = tid; i < TQ_MAX; i += NTH) { row_m[i] = -INFINITY; row_l[i] = 0.f; row_any[i] = 0; }
  const int n_tiles = (Tk + KT - 1) / KT;
  auto fetch_k = [&](int t) {
    if (t < n_tiles) load_tile_b_async<KT>(Kb + (t & 1) * KT * PB, k + (int64_t)t * KT * p.ldk, p.ldk, min(KT, Tk - t * KT), p.vec);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto fetch_v = [&](int t) {
    if (t < n_tiles) load_tile_b_async<KT>(Vb + (t & 1) * KT * PB, v + (int64_t)t * KT * p.ldv, p.ldv, min(KT, Tk - t * KT), p.vec);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch_k(0);
  for (int t = 0; t < n_tiles; ++t) {              // pass 1: row statistics
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    fetch_k(t + 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const uint32_t sK = sK0 + (uint32_t)((t & 1) * KT * PB * 2);
    float acc[RA][4];
    mma_nt<RA>(sQ, sK, acc, warp, lane);
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      Sf[(16 * a + g) * PF + 8 * warp + 2 * t4] = acc[a][0] * p.scale;
      Sf[(16 * a + g) * PF + 8 * warp + 2 * t4 + 1] = acc[a][1] * p.scale;
      Sf[(16 * a + g + 8) * PF + 8 * warp + 2 * t4] = acc[a][2] * p.scale;
      Sf[(16 * a + g + 8) * PF + 8 * warp + 2 * t4 + 1] = acc[a][3] * p.scale;
    }
    __syncthreads();
    row_stats_update_bits(Sf, Mw, has_mask, per_row, Tq, k0, nk, row_m, row_l, row_any, warp, lane);
    __syncthreads();
  }
  float oacc[RA][4];
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[a][e] = 0.f;
  const float inv_tk = 1.f / (float)Tk;
  fetch_k(0);
  fetch_v(0);
  for (int t = 0; t < n_tiles; ++t) {              // pass 2: P~ and O += P~ V
    const int k0 = t * KT, nk = min(KT, Tk - k0);
    fetch_k(t + 1);
    fetch_v(t + 1);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    const uint32_t sK = sK0 + (uint32_t)((t & 1) * KT * PB * 2), sV = sV0 + (uint32_t)((t & 1) * KT * PB * 2);
    {
      float acc[RA][4];
      mma_nt<RA>(sQ, sK, acc, warp, lane);
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int i = 16 * a + g + 8 * hr;
          const bool row_ok = i < Tq;
          const uint64_t rowbase = ((uint64_t)(b * p.h + hh) * Tq + i) * Tk + (uint64_t)(k0 + 8 * warp + 2 * t4);
          const uint32_t mword = has_mask ? Mw[(per_row ? i : 0) * MW + ((k0 + 8 * warp) >> 5)] >> ((k0 + 8 * warp + 2 * t4) & 31) : 3u;
          const int any = row_ok ? row_any[i] : 0;
          const float m = row_m[i], inv_l = any ? 1.f / row_l[i] : 0.f;
          float pr2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 8 * warp + 2 * t4 + e;
            float pr = 0.f;
            if (row_ok && j < nk) {
              const bool kept = (mword >> e) & 1u;
              pr = any ? (kept ? __expf(acc[a][2 * hr + e] * p.scale - m) * inv_l : 0.f) : inv_tk;
              if (drop && pr != 0.f) {
                pr *= idx32 ? keep_factor32(seed_lo, c2, (uint32_t)rowbase + e, p.threshold, p.inv_keep)
                            : (dropout_keep(seed, p.site, rowbase + e, p.threshold) ? p.inv_keep : 0.f);
              }
            }
            pr2[e] = pr;
          }
          *reinterpret_cast<__nv_bfloat162*>(Pb + i * PB + 8 * warp + 2 * t4) = __floats2bfloat162_rn(pr2[0], pr2[1]);
        }
    }
    __syncthreads();
    {
      uint32_t bv[4][2];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x2_t(bv[ks], b_kn(sV, 8 * warp, 16 * ks, lane));
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t ap[4];
          ldsm_x4(ap, a_rowmajor(sP, 16 * a, 16 * ks, lane));
          mma16816(oacc[a], ap, bv[ks]);
        }
    }
    __syncthreads();
  }
  T* op = (T*)p.out + (int64_t)b * p.sbo + hh * DK;
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int i = 16 * a + g + 8 * hr;
      if (i < Tq)
        *reinterpret_cast<__nv_bfloat162*>(op + (int64_t)i * p.ldo + 8 * warp + 2 * t4) = __floats2bfloat162_rn(oacc[a][2 * hr], oacc[a][2 * hr + 1]);
    }
}

template <int RA>
int launch_bwd_mma(const BwdArgs& p, int64_t B, cudaStream_t st) {
  constexpr int TQP = 16 * RA;
  const size_t smem = (size_t)(4 * TQP + 4 * KT) * PB * 2 + (size_t)TQP * PF * 4 + (size_t)TQP * MW * 4;
  if (cudaFuncSetAttribute(attn_bwd_small_mma_kernel<RA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  attn_bwd_small_mma_kernel<RA><<<(unsigned)(B * p.h), NTH, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}
template <int RA>
int launch_fwd_mma(const FwdArgs& p, int64_t B, cudaStream_t st) {
  constexpr int TQP = 16 * RA;
  const size_t smem = (size_t)(2 * TQP + 4 * KT) * PB * 2 + (size_t)TQP * PF * 4 + (size_t)TQP * MW * 4;
  if (cudaFuncSetAttribute(attn_fwd_small_mma_kernel<RA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  attn_fwd_small_mma_kernel<RA><<<(unsigned)(B * p.h), NTH, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}
#define CN_RA_SWITCH(FN, P, B_, ST)                \
  switch ((P.Tq + 15) / 16) {                      \
    case 1: return FN<1>(P, B_, ST);               \
    case 2: return FN<2>(P, B_, ST);               \
    case 3: return FN<3>(P, B_, ST);               \
    case 4: return FN<4>(P, B_, ST);               \
    case 5: return FN<5>(P, B_, ST);               \
    case 6: return FN<6>(P, B_, ST);               \
    case 7: return FN<7>(P, B_, ST);               \
    case 8: return FN<8>(P, B_, ST);               \
    default: return CN_ERR_UNSUPPORTED;            \
  }
int dispatch_bwd_mma(const BwdArgs& p, int64_t B, cudaStream_t st) { CN_RA_SWITCH(launch_bwd_mma, p, B, st) }
int dispatch_fwd_mma(const FwdArgs& p, int64_t B, cudaStream_t st) { CN_RA_SWITCH(launch_fwd_mma, p, B, st) }

template <typename T, int RA>
int launch_bwd(const BwdArgs& p, int64_t B, cudaStream_t st) {
  constexpr int TQP = 16 * RA;
  const size_t smem = (size_t)(5 * TQP + 2 * KT) * PITCH * sizeof(float);
  if (cudaFuncSetAttribute(attn_bwd_small_kernel<T, RA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  attn_bwd_small_kernel<T, RA><<<(unsigned)(B * p.h), NTH, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

template <typename T>
int dispatch_bwd(const BwdArgs& p, int64_t B, cudaStream_t st) {
  const int ra = (p.Tq + 15) / 16;
  switch (ra) {
    case 1: return launch_bwd<T, 1>(p, B, st);
    case 2: return launch_bwd<T, 2>(p, B, st);
    case 3: return launch_bwd<T, 3>(p, B, st);
    case 4: return launch_bwd<T, 4>(p, B, st);
    case 5: return launch_bwd<T, 5>(p, B, st);
    case 6: return launch_bwd<T, 6>(p, B, st);
    case 7: return launch_bwd<T, 7>(p, B, st);
    case 8: return launch_bwd<T, 8>(p, B, st);
    default: return CN_ERR_UNSUPPORTED;
  }
}

}  // namespace

extern "C" int cn_attention_bwd_small(const void* q, int64_t ldq, int64_t sbq, const void* k, int64_t ldk, int64_t sbk,
                                      const void* v, int64_t ldv, int64_t sbv, const void* o, int64_t ldo, int64_t sbo,
                                      const void* go, int64_t ldg, int64_t sbg, int is_bf16, const uint8_t* mask, int64_t mb,
                                      int64_t mi, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk, float p_drop,
                                      const int64_t* seed_dev, int64_t seed, int64_t site, void* dq, int64_t lddq, int64_t sbdq,
                                      void* dk_out, int64_t lddk, int64_t sbdk, void* dv, int64_t lddv, int64_t sbdv,
                                      cn_stream_t stream) {
  if (!q || !k || !v || !o || !go || !dq || !dk_out || !dv || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (!(p_drop >= 0.f) || p_drop >= 1.f) return CN_ERR_BAD_ARG;
  if (dk != DK || Tq > TQ_MAX || B * h > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  if (B == 0 || Tq == 0) return CN_OK;
  BwdArgs p;
  p.q = q; p.k = k; p.v = v; p.o = o; p.go = go;
  p.ldq = ldq; p.sbq = sbq; p.ldk = ldk; p.sbk = sbk; p.ldv = ldv; p.sbv = sbv; p.ldo = ldo; p.sbo = sbo; p.ldg = ldg; p.sbg = sbg;
  p.dq = dq; p.dk = dk_out; p.dv = dv;
  p.lddq = lddq; p.sbdq = sbdq; p.lddk = lddk; p.sbdk = sbdk; p.lddv = lddv; p.sbdv = sbdv;
  p.mask = mask; p.mb = mb; p.mi = mi;
  p.h = (int)h; p.Tq = (int)Tq; p.Tk = (int)Tk;
  p.scale = 1.f / sqrtf((float)dk);
  p.threshold = (uint32_t)((double)p_drop * 4294967296.0);
  p.inv_keep = 1.f / (1.f - p_drop);
  {
    const int64_t es = is_bf16 ? 2 : 4;
    const uintptr_t ptrs = (uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)go;
    const int64_t pit = (ldq | sbq | ldk | sbk | ldv | sbv | ldg | sbg) * es;
    p.vec = !(ptrs & 15) && !(pit & 15);
  }
  p.site = (uint32_t)site; p.seed_dev = seed_dev; p.seed = (uint64_t)seed;
  // bf16: tensor-core kernel when the outputs allow paired stores; f32: FFMA kernel
  if (is_bf16 && Tk <= 32 * MW && !((lddq | sbdq | lddk | sbdk | lddv | sbdv) & 1) && !(((uintptr_t)dq | (uintptr_t)dk_out | (uintptr_t)dv) & 3))
    return dispatch_bwd_mma(p, B, (cudaStream_t)stream);
  return is_bf16 ? dispatch_bwd<__nv_bfloat16>(p, B, (cudaStream_t)stream) : dispatch_bwd<float>(p, B, (cudaStream_t)stream);
}

extern "C" int cn_attention_fwd_small(const void* q, int64_t ldq, int64_t sbq, const void* k, int64_t ldk, int64_t sbk,
                                      const void* v, int64_t ldv, int64_t sbv, int is_bf16, const uint8_t* mask, int64_t mb,
                                      int64_t mi, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk, float p_drop,
                                      const int64_t* seed_dev, int64_t seed, int64_t site, void* out, int64_t ldo, int64_t sbo,
                                      cn_stream_t stream) {
  if (!q || !k || !v || !out || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (!(p_drop >= 0.f) || p_drop >= 1.f) return CN_ERR_BAD_ARG;
  if (dk != DK || Tq > TQ_MAX || B * h > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  if (B == 0 || Tq == 0) return CN_OK;
  FwdArgs p;
  p.q = q; p.k = k; p.v = v; p.out = out;
  p.ldq = ldq; p.sbq = sbq; p.ldk = ldk; p.sbk = sbk; p.ldv = ldv; p.sbv = sbv; p.ldo = ldo; p.sbo = sbo;
  p.mask = mask; p.mb = mb; p.mi = mi;
  p.h = (int)h; p.Tq = (int)Tq; p.Tk = (int)Tk;
  p.scale = 1.f / sqrtf((float)dk);
  p.threshold = (uint32_t)((double)p_drop * 4294967296.0);
  p.inv_keep = 1.f / (1.f - p_drop);
  {
    const int64_t es = is_bf16 ? 2 : 4;
    const uintptr_t ptrs = (uintptr_t)q | (uintptr_t)k | (uintptr_t)v;
    const int64_t pit = (ldq | sbq | ldk | sbk | ldv | sbv) * es;
    p.vec = !(ptrs & 15) && !(pit & 15);
  }
  p.site = (uint32_t)site; p.seed_dev = seed_dev; p.seed = (uint64_t)seed;
  if (is_bf16 && Tk <= 32 * MW && !((ldo | sbo) & 1) && !((uintptr_t)out & 3)) return dispatch_fwd_mma(p, B, (cudaStream_t)stream);
  return is_bf16 ? dispatch_fwd<__nv_bfloat16>(p, B, (cudaStream_t)stream) : dispatch_fwd<float>(p, B, (cudaStream_t)stream);
}
