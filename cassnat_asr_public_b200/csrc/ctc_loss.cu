// K10: CTC loss forward-backward fused with the log-softmax backward.
//   replaces: torch.nn.CTCLoss(reduction='mean', zero_infinity=True)   src/models/cassnat.py:126,232
//             (arithmetic lives in ATen's ctc_loss; restated here from the published alpha-beta
//             recursion, Graves et al. 2006, with PyTorch's conventions: blank-augmented target
//             l' of 2L+1 states, per-utterance nll / max(L,1), mean over the batch, inf -> 0)
// Three launches:
//   1. gather   E[b,t,i] = logp[b,t,l_i]            (i = 0 is blank)       fully parallel
//   2. alpha-beta: two warps per utterance running CONCURRENTLY -- warp 0 the forward variable from frame 0 up,
//      warp 1 the backward variable from the last frame down -- lattice states strided over lanes (SPL per lane),
//      values in registers / neighbours by shuffle, the rows written to the workspace.  The recursion runs in the
//      log2 domain: the three-way log-sum-exp of a state is max + lg2(ex2 + ex2 + ex2) on the MUFU unit (rows are
//      stored in nats); rows are renormalised (max state = 0, the removed scale accumulated in double) every 8 frames
//      instead of every frame; the emissions of the next 16 frames arrive by cp.async while the current 16 are used
//   3. gradient w.r.t. the LOGITS, one CTA per frame: softmax - occupancy, written once:
//      d loss / d z[b,t,c] = scale_b * (exp(logp[b,t,c]) - sum_{s: l'_s = c} exp(ab_t(s) + nll_b - logp[b,t,c]))
#include "common.cuh"

namespace {

__device__ __forceinline__ float log_add(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

__global__ void __launch_bounds__(256) ctc_gather_kernel(const float* __restrict__ logp, int64_t sb, int64_t st,
                                                         const int64_t* __restrict__ targets, int64_t tgt_stride,
                                                         int T, int L, int V, int64_t blank, float* __restrict__ E) {
  const int t = blockIdx.x, b = blockIdx.y;
  const float* row = logp + (int64_t)b * sb + (int64_t)t * st;
  float* out = E + ((int64_t)b * T + t) * (L + 1);
  for (int i = threadIdx.x; i <= L; i += blockDim.x) {
    int64_t c = (i == 0) ? blank : targets[(int64_t)b * tgt_stride + (i - 1)];
    c = c < 0 ? 0 : (c >= V ? V - 1 : c);
    out[i] = cn_ld_stream_f32(row + c);
  }
}

struct AbArgs {
  const float* E; const int64_t* targets; int64_t tgt_stride;
  const int64_t* in_len; const int64_t* tgt_len;
  int T, L; float* ab; float* bb; float* nll;
  double* off_a;      // [B,T] log-scale (nats) removed from the stored alpha rows
  double* off_b;      // [B,T] the same for the beta rows
  double* nll_d;      // [B]
};

constexpr int kCtcChunk = 16;        // frames per staged emission chunk
constexpr int kCtcNorm = 8;          // frames between renormalisations
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// log2(2^a + 2^b + 2^c); -inf operands contribute nothing, all -inf gives -inf
__device__ __forceinline__ float lse2_3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  const float ms = (m == -INFINITY) ? 0.f : m;
  return ms + fast_lg2(fast_ex2(a - ms) + fast_ex2(b - ms) + fast_ex2(c - ms));
}
__device__ __forceinline__ float lse2_2(float a, float b) {
  const float m = fmaxf(a, b);
  const float ms = (m == -INFINITY) ? 0.f : m;
  return ms + fast_lg2(fast_ex2(a - ms) + fast_ex2(b - ms));
}

// CTA = 2 warps per utterance: warp 0 alpha (frames up), warp 1 beta (frames down); lane owns states
// [lane*SPL, lane*SPL + SPL); SPL even -> odd j = label state
template <int SPL>
__global__ void __launch_bounds__(64) ctc_alpha_beta_kernel(AbArgs p) {
  extern __shared__ float esm[];                          // [2 warps][2 buffers][kCtcChunk][L+1] staged emissions
  __shared__ float fin[32 * SPL];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, T = p.T, L1 = L + 1;
  int Tb = (int)p.in_len[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int Lb = (int)p.tgt_len[b];
  Lb = Lb < 0 ? 0 : (Lb > L ? L : Lb);
  const int Sb = 2 * Lb + 1;                              // states of this utterance
  const int S = 32 * SPL;                                 // padded row length in the workspace
  const float* Eb = p.E + (int64_t)b * T * L1;
  const int64_t* tg = p.targets + (int64_t)b * p.tgt_stride;
  // skip transitions: into odd state s from s-2 iff labels differ (forward); out of s into s+2 (backward)
  uint32_t skip_in = 0, skip_out = 0, live = 0;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = lane * SPL + j;
    if (s < Sb) live |= 1u << j;
    if ((j & 1) && s < Sb) {
      const int i = (s - 1) >> 1;                         // label index 0..Lb-1
      if (i >= 1 && tg[i] != tg[i - 1]) skip_in |= 1u << j;
      if (i + 1 < Lb && tg[i + 1] != tg[i]) skip_out |= 1u << j;
    }
  }
  if (Tb == 0) {
    if (threadIdx.x == 0) { p.nll[b] = (Lb == 0) ? 0.f : INFINITY; p.nll_d[b] = (Lb == 0) ? 0.0 : (double)INFINITY; }
    return;
  }
  // ---- emission staging: chunk c = frames [c*16, c*16+16), both warps walk their own sequence of chunks
  const int n_chunks = (Tb + kCtcChunk - 1) / kCtcChunk;
  float* ering = esm + warp * 2 * kCtcChunk * L1;
  auto stage = [&](int c, int buf) {                      // asynchronous copy of chunk c (contiguous rows of E)
    if (c >= 0 && c < n_chunks) {
      const int t0 = c * kCtcChunk, n = (min(kCtcChunk, Tb - t0)) * L1;
      const float* src = Eb + (int64_t)t0 * L1;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ering + buf * kCtcChunk * L1);
      for (int i = lane; i < n; i += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * (uint32_t)i), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // emissions of frame t (log2 units) for this lane's states, out of the staged chunk
  auto emissions = [&](const float* row, float& eb, float (&el)[SPL / 2]) {
    eb = row[0] * kLog2e;
#pragma unroll
    for (int j = 0; j < SPL / 2; ++j) {
      const int s = lane * SPL + 2 * j + 1;
      el[j] = s < Sb ? row[(s + 1) >> 1] * kLog2e : -INFINITY;
    }
  };
  // Rows are kept NORMALISED (max state = 0) and the removed log-scale is accumulated in double:
  // raw alpha reaches ~ -8 nats/frame, where fp32 has only ~1e-4 of absolute resolution.
  auto normalise = [&](float* r) -> float {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < SPL; ++j) m = fmaxf(m, r[j]);
    m = cn_warp_max(m);
    if (m > -INFINITY) {
#pragma unroll
      for (int j = 0; j < SPL; ++j) r[j] -= m;
      return m;
    }
    return 0.f;
  };
  float x[SPL];                                           // alpha (warp 0) / beta (warp 1), log2 units
  double cx = 0.0;                                        // removed scale, log2 units
  if (warp == 0) {
    // =========================== forward
    float* rows = p.ab + (int64_t)b * T * S;
    double* offs = p.off_a + (int64_t)b * T;
    stage(0, 0);
    stage(1, 1);
    for (int c = 0; c < n_chunks; ++c) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
      const float* e = ering + (c & 1) * kCtcChunk * L1;
      const int t0 = c * kCtcChunk, nf = min(kCtcChunk, Tb - t0);
      for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        float eb, el[SPL / 2];
        emissions(e + f * L1, eb, el);
        if (t == 0) {
#pragma unroll
          for (int j = 0; j < SPL; ++j) {
            const int s = lane * SPL + j;
            x[j] = (s < 2 && s < Sb) ? ((j & 1) ? el[j >> 1] : eb) : -INFINITY;
          }
        } else {
          float up1 = __shfl_up_sync(0xffffffffu, x[SPL - 1], 1), up2 = __shfl_up_sync(0xffffffffu, x[SPL - 2], 1);
          if (lane == 0) { up1 = -INFINITY; up2 = -INFINITY; }
          float nx[SPL];
#pragma unroll
          for (int j = 0; j < SPL; ++j) {
            const float c1 = (j >= 1) ? x[j - 1] : up1;
            float v;
            if (j & 1) {
              const float c2 = ((skip_in >> j) & 1u) ? ((j >= 2) ? x[j - 2] : up1) : -INFINITY;
              v = lse2_3(x[j], c1, c2);
            } else {
              v = lse2_2(x[j], c1);
            }
            nx[j] = ((live >> j) & 1u) ? v + ((j & 1) ? el[j >> 1] : eb) : -INFINITY;
          }
#pragma unroll
          for (int j = 0; j < SPL; ++j) x[j] = nx[j];
        }
        if (t % kCtcNorm == 0 || t == Tb - 1) cx += (double)normalise(x);
        if (lane == 0) offs[t] = cx * (double)kLn2;
#pragma unroll
        for (int j = 0; j < SPL; ++j) rows[(int64_t)t * S + lane * SPL + j] = x[j] * kLn2;
      }
      __syncwarp();
      stage(c + 2, c & 1);
    }
    // ---- nll = -log( alpha_{Tb-1}(Sb-1) + alpha_{Tb-1}(Sb-2) )
#pragma unroll
    for (int j = 0; j < SPL; ++j) fin[lane * SPL + j] = x[j];
    __syncwarp();
    const float l1 = fin[Sb - 1], l2 = (Sb >= 2) ? fin[Sb - 2] : -INFINITY;
    const double nll_d = -((double)lse2_2(l1, l2) + cx) * (double)kLn2;
    if (lane == 0) { p.nll[b] = (float)nll_d; p.nll_d[b] = nll_d; }
  } else {
    // =========================== backward
    float* rows = p.bb + (int64_t)b * T * S;
    double* offs = p.off_b + (int64_t)b * T;
    stage(n_chunks - 1, (n_chunks - 1) & 1);
    stage(n_chunks - 2, (n_chunks - 2) & 1);
    for (int c = n_chunks - 1; c >= 0; --c) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
      const float* e = ering + (c & 1) * kCtcChunk * L1;
      const int t0 = c * kCtcChunk, nf = min(kCtcChunk, Tb - t0);
      for (int f = nf - 1; f >= 0; --f) {
        const int t = t0 + f;
        float eb, el[SPL / 2];
        emissions(e + f * L1, eb, el);
        if (t == Tb - 1) {
#pragma unroll
          for (int j = 0; j < SPL; ++j) {
            const int s = lane * SPL + j;
            x[j] = (s < Sb && s >= Sb - 2) ? ((j & 1) ? el[j >> 1] : eb) : -INFINITY;
          }
        } else {
          float dn1 = __shfl_down_sync(0xffffffffu, x[0], 1), dn2 = __shfl_down_sync(0xffffffffu, x[1], 1);
          if (lane == 31) { dn1 = -INFINITY; dn2 = -INFINITY; }
          float nx[SPL];
#pragma unroll
          for (int j = 0; j < SPL; ++j) {
            const float c1 = (j + 1 < SPL) ? x[j + 1] : dn1;
            float v;
            if (j & 1) {
              const float c2 = ((skip_out >> j) & 1u) ? ((j + 2 < SPL) ? x[j + 2] : dn2) : -INFINITY;
              v = lse2_3(x[j], c1, c2);
            } else {
              v = lse2_2(x[j], c1);
            }
            nx[j] = ((live >> j) & 1u) ? v + ((j & 1) ? el[j >> 1] : eb) : -INFINITY;
          }
#pragma unroll
          for (int j = 0; j < SPL; ++j) x[j] = nx[j];
        }
        if (t % kCtcNorm == 0 || t == Tb - 1) cx += (double)normalise(x);
        if (lane == 0) offs[t] = cx * (double)kLn2;
#pragma unroll
        for (int j = 0; j < SPL; ++j) rows[(int64_t)t * S + lane * SPL + j] = x[j] * kLn2;
      }
      __syncwarp();
      stage(c - 2, c & 1);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

struct GradArgs {
  const float* logp; int64_t sb, st;
  const int64_t* targets; int64_t tgt_stride;
  const int64_t* in_len; const int64_t* tgt_len;
  const float* ab; const float* bb; const float* nll; const double* off_a; const double* off_b; const double* nll_d;
  int T, L, V, S, B; int64_t blank; float grad_out;
  float* grad; int64_t gsb, gst;
};

// one CTA per (frame, utterance): dense softmax part, then the <= L+1 occupancy corrections
__global__ void __launch_bounds__(256) ctc_grad_kernel(GradArgs p) {
  const int t = blockIdx.x, b = blockIdx.y;
  const float* row = p.logp + (int64_t)b * p.sb + (int64_t)t * p.st;
  float* g = p.grad + (int64_t)b * p.gsb + (int64_t)t * p.gst;
  int Tb = (int)p.in_len[b];
  Tb = Tb > p.T ? p.T : Tb;
  int Lb = (int)p.tgt_len[b];
  Lb = Lb < 0 ? 0 : (Lb > p.L ? p.L : Lb);
  const float nll = p.nll[b];
  const bool dead = (t >= Tb) || !(nll < INFINITY);       // padded frame, or zero_infinity
  const float scale = dead ? 0.f : p.grad_out / ((float)p.B * (float)(Lb > 1 ? Lb : 1));
  for (int c = threadIdx.x; c < p.V; c += blockDim.x) g[c] = dead ? 0.f : scale * expf(row[c]);
  if (dead) return;
  __syncthreads();
  const int64_t* tg = p.targets + (int64_t)b * p.tgt_stride;
  const float* abt = p.ab + ((int64_t)b * p.T + t) * p.S;      // alpha_t rows; beta_t rows at the same offset of bb
  const float* bbt = p.bb + ((int64_t)b * p.T + t) * p.S;
  // log-scale of this frame's rows + nll
  const float shift = (float)(p.off_a[(int64_t)b * p.T + t] + p.off_b[(int64_t)b * p.T + t] + p.nll_d[b]);
  // blank column: all even states, warp-tree reduction in a fixed order (deterministic)
  if (threadIdx.x < 32) {
    float acc = -INFINITY;
    for (int i = threadIdx.x; i <= Lb; i += 32) acc = log_add(acc, abt[2 * i] + bbt[2 * i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = log_add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (threadIdx.x == 0) {
      const int64_t cb = p.blank;
      g[cb] = scale * (expf(row[cb]) - expf(acc + shift - row[cb]));
    }
  }
  // label columns: the first occurrence of a label sums all its occurrences, in target order
  for (int i = threadIdx.x; i < Lb; i += blockDim.x) {
    const int64_t c = tg[i];
    bool leader = true;
    for (int k = 0; k < i; ++k)
      if (tg[k] == c) { leader = false; break; }
    if (!leader || c < 0 || c >= p.V || c == p.blank) continue;
    float acc = abt[2 * i + 1] + bbt[2 * i + 1];
    for (int k = i + 1; k < Lb; ++k)
      if (tg[k] == c) acc = log_add(acc, abt[2 * k + 1] + bbt[2 * k + 1]);
    g[c] = scale * (expf(row[c]) - expf(acc + shift - row[c]));
  }
}

__global__ void ctc_reduce_loss_kernel(const float* __restrict__ nll, const int64_t* __restrict__ tgt_len, int B, int L,
                                       float* __restrict__ loss) {
  // mean over the batch of nll / max(L_b, 1), inf -> 0 ; sequential for a reproducible sum
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      int Lb = (int)tgt_len[b];
      Lb = Lb > L ? L : Lb;
      const float v = nll[b];
      if (v < INFINITY) acc += v / (float)(Lb > 1 ? Lb : 1);
    }
    *loss = acc / (float)B;
  }
}

int spl_for(int64_t L) {
  int spl = (int)((2 * L + 1 + 31) / 32);
  if (spl < 2) spl = 2;
  if (spl & 1) ++spl;
  if (spl > 16) spl = ((spl + 3) / 4) * 4;
  return spl;
}

template <int SPL>
int launch_ab(const AbArgs& p, int64_t B, cudaStream_t s) {
  const size_t smem = (size_t)2 * 2 * kCtcChunk * (p.L + 1) * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(ctc_alpha_beta_kernel<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  ctc_alpha_beta_kernel<SPL><<<(unsigned)B, 64, smem, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

}  // namespace

extern "C" int64_t cn_ctc_loss_workspace_bytes(int64_t B, int64_t T, int64_t L) {
  if (B < 0 || T < 0 || L < 0 || L > 511) return -1;
  const int64_t S = 32 * (int64_t)spl_for(L);
  return (B * T * (L + 1) + 2 * B * T * S) * 4 + (2 * B * T + B) * 8 + 512;
}

extern "C" int cn_ctc_loss(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                           const int64_t* targets, int64_t L, int64_t tgt_stride, const int64_t* input_lengths,
                           const int64_t* target_lengths, int64_t blank, float grad_out, float* loss, float* nll,
                           float* grad_logits, int64_t gsb, int64_t gst, void* workspace, int64_t workspace_bytes,
                           cn_stream_t stream_) {
  if (!logp || !targets || !input_lengths || !target_lengths || !loss || !nll || !workspace) return CN_ERR_BAD_ARG;
  if (B <= 0 || T <= 0 || V <= 0 || L < 0 || tgt_stride < L) return CN_ERR_BAD_ARG;
  if (L > 511 || B > 65535 || T > 65535) return CN_ERR_UNSUPPORTED;
  if (workspace_bytes < cn_ctc_loss_workspace_bytes(B, T, L)) return CN_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int spl = spl_for(L);
  const int S = 32 * spl;
  float* E = (float*)workspace;
  float* ab = E + B * T * (L + 1);
  float* bb = ab + B * T * S;
  double* off_a = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(bb + B * T * S) + 255) & ~uintptr_t(255));
  double* off_b = off_a + B * T;
  double* nll_dbl = off_b + B * T;
  ctc_gather_kernel<<<dim3((unsigned)T, (unsigned)B), 256, 0, stream>>>(logp, sb, st, targets, tgt_stride, (int)T, (int)L,
                                                                         (int)V, blank, E);
  CN_CHECK_LAUNCH();
  AbArgs p;
  p.E = E; p.targets = targets; p.tgt_stride = tgt_stride; p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.T = (int)T; p.L = (int)L; p.ab = ab; p.bb = bb; p.nll = nll; p.off_a = off_a; p.off_b = off_b; p.nll_d = nll_dbl;
  int rc;
  switch (spl) {
    case 2: rc = launch_ab<2>(p, B, stream); break;
    case 4: rc = launch_ab<4>(p, B, stream); break;
    case 6: rc = launch_ab<6>(p, B, stream); break;
    case 8: rc = launch_ab<8>(p, B, stream); break;
    case 10: rc = launch_ab<10>(p, B, stream); break;
    case 12: rc = launch_ab<12>(p, B, stream); break;
    case 14: rc = launch_ab<14>(p, B, stream); break;
    case 16: rc = launch_ab<16>(p, B, stream); break;
    case 20: rc = launch_ab<20>(p, B, stream); break;
    case 24: rc = launch_ab<24>(p, B, stream); break;
    case 28: rc = launch_ab<28>(p, B, stream); break;
    case 32: rc = launch_ab<32>(p, B, stream); break;
    default: return CN_ERR_UNSUPPORTED;
  }
  if (rc != CN_OK) return rc;
  ctc_reduce_loss_kernel<<<1, 32, 0, stream>>>(nll, target_lengths, (int)B, (int)L, loss);
  CN_CHECK_LAUNCH();
  if (grad_logits) {
    GradArgs g;
    g.logp = logp; g.sb = sb; g.st = st; g.targets = targets; g.tgt_stride = tgt_stride; g.in_len = input_lengths;
    g.tgt_len = target_lengths; g.ab = ab; g.bb = bb; g.nll = nll; g.off_a = off_a; g.off_b = off_b; g.nll_d = nll_dbl; g.T = (int)T; g.L = (int)L; g.V = (int)V; g.S = S; g.B = (int)B;
    g.blank = blank; g.grad_out = grad_out; g.grad = grad_logits; g.gsb = gsb; g.gst = gst;
    ctc_grad_kernel<<<dim3((unsigned)T, (unsigned)B), 256, 0, stream>>>(g);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}
