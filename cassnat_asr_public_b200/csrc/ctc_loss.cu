// K10: CTC loss forward-backward fused with the log-softmax backward.
//   replaces: torch.nn.CTCLoss(reduction='mean', zero_infinity=True)   src/models/cassnat.py:126,232
//             (arithmetic lives in ATen's ctc_loss; restated here from the published alpha-beta
//             recursion, Graves et al. 2006, with PyTorch's conventions: blank-augmented target
//             l' of 2L+1 states, per-utterance nll / max(L,1), mean over the batch, inf -> 0)
// Three launches:
//   1. gather   E[b,t,i] = logp[b,t,l_i]            (i = 0 is blank)       fully parallel
//   2. alpha-beta: one warp per utterance, lattice states strided over lanes (SPL per lane),
//      alpha in registers / neighbours by shuffle, alpha_t spilled to the workspace, beta fused
//      on the way back so that the workspace ends up holding alpha_t(s) + beta_t(s)
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
  int T, L; float* ab; float* nll;
  double* off_a;      // [B,T] running log-scale of the normalised alpha rows
  float* off;         // [B,T] off_a[t] + off_b[t] + nll: what turns (alpha^ + beta^) into a log-occupancy
};

// one warp per utterance; lane owns states [lane*SPL, lane*SPL + SPL); SPL even -> odd j = label state
template <int SPL>
__global__ void __launch_bounds__(32) ctc_alpha_beta_kernel(AbArgs p) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int L = p.L, T = p.T;
  int Tb = (int)p.in_len[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  int Lb = (int)p.tgt_len[b];
  Lb = Lb < 0 ? 0 : (Lb > L ? L : Lb);
  const int Sb = 2 * Lb + 1;                              // states of this utterance
  const int S = 32 * SPL;                                 // padded row length in the workspace
  const float* Eb = p.E + (int64_t)b * T * (L + 1);
  float* abb = p.ab + (int64_t)b * T * S;
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
    if (lane == 0) p.nll[b] = (Lb == 0) ? 0.f : INFINITY;
    return;
  }
  auto emission = [&](int t, int j) -> float {
    const int s = lane * SPL + j;
    if (s >= Sb) return -INFINITY;
    return Eb[(int64_t)t * (L + 1) + ((j & 1) ? ((s + 1) >> 1) : 0)];
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
  double* offa = p.off_a + (int64_t)b * T;
  float* offo = p.off + (int64_t)b * T;
  float a[SPL];
  double ca = 0.0;
  // ---- forward
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = lane * SPL + j;
    a[j] = (s < 2 && s < Sb) ? emission(0, j) : -INFINITY;
  }
  ca += (double)normalise(a);
  if (lane == 0) offa[0] = ca;
#pragma unroll
  for (int j = 0; j < SPL; ++j) abb[lane * SPL + j] = a[j];
  for (int t = 1; t < Tb; ++t) {
    float up1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1), up2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
    if (lane == 0) { up1 = -INFINITY; up2 = -INFINITY; }
    float na[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const float c1 = (j >= 1) ? a[j - 1] : up1;
      const float c2 = ((skip_in >> j) & 1u) ? ((j >= 2) ? a[j - 2] : (j == 1 ? up1 : up2)) : -INFINITY;
      float v = log_add(log_add(a[j], c1), c2);
      na[j] = ((live >> j) & 1u) ? v + emission(t, j) : -INFINITY;
    }
    ca += (double)normalise(na);
    if (lane == 0) offa[t] = ca;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      a[j] = na[j];
      abb[(int64_t)t * S + lane * SPL + j] = a[j];
    }
  }
  // ---- nll = -log( alpha_{Tb-1}(Sb-1) + alpha_{Tb-1}(Sb-2) )
  __shared__ float fin[32 * SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) fin[lane * SPL + j] = a[j];
  __syncwarp();
  const float l1 = fin[Sb - 1], l2 = (Sb >= 2) ? fin[Sb - 2] : -INFINITY;
  const double nll_d = -((double)log_add(l1, l2) + ca);
  if (lane == 0) p.nll[b] = (float)nll_d;
  // ---- backward, fused with ab = alpha + beta (overwrites the stored alpha)
  float be[SPL];
  double cb = 0.0;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = lane * SPL + j;
    be[j] = (s < Sb && s >= Sb - 2) ? emission(Tb - 1, j) : -INFINITY;
  }
  cb += (double)normalise(be);
  if (lane == 0) offo[Tb - 1] = (float)(offa[Tb - 1] + cb + nll_d);
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    float* slot = abb + (int64_t)(Tb - 1) * S + lane * SPL + j;
    *slot = *slot + be[j];
  }
  for (int t = Tb - 2; t >= 0; --t) {
    float dn1 = __shfl_down_sync(0xffffffffu, be[0], 1), dn2 = __shfl_down_sync(0xffffffffu, be[1], 1);
    if (lane == 31) { dn1 = -INFINITY; dn2 = -INFINITY; }
    float nb[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const float c1 = (j + 1 < SPL) ? be[j + 1] : dn1;
      const float c2 = ((skip_out >> j) & 1u) ? ((j + 2 < SPL) ? be[j + 2] : (j + 2 == SPL ? dn1 : dn2)) : -INFINITY;
      float v = log_add(log_add(be[j], c1), c2);
      nb[j] = ((live >> j) & 1u) ? v + emission(t, j) : -INFINITY;
    }
    cb += (double)normalise(nb);
    if (lane == 0) offo[t] = (float)(offa[t] + cb + nll_d);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      be[j] = nb[j];
      float* slot = abb + (int64_t)t * S + lane * SPL + j;
      *slot = *slot + be[j];
    }
  }
}

struct GradArgs {
  const float* logp; int64_t sb, st;
  const int64_t* targets; int64_t tgt_stride;
  const int64_t* in_len; const int64_t* tgt_len;
  const float* ab; const float* nll; const float* off;
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
  const float* abt = p.ab + ((int64_t)b * p.T + t) * p.S;
  const float shift = p.off[(int64_t)b * p.T + t];       // log-scale of this frame's rows + nll
  // blank column: all even states, warp-tree reduction in a fixed order (deterministic)
  if (threadIdx.x < 32) {
    float acc = -INFINITY;
    for (int i = threadIdx.x; i <= Lb; i += 32) acc = log_add(acc, abt[2 * i]);
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
    float acc = abt[2 * i + 1];
    for (int k = i + 1; k < Lb; ++k)
      if (tg[k] == c) acc = log_add(acc, abt[2 * k + 1]);
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
  ctc_alpha_beta_kernel<SPL><<<(unsigned)B, 32, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

}  // namespace

extern "C" int64_t cn_ctc_loss_workspace_bytes(int64_t B, int64_t T, int64_t L) {
  if (B < 0 || T < 0 || L < 0 || L > 511) return -1;
  const int64_t S = 32 * (int64_t)spl_for(L);
  return (B * T * (L + 1) + B * T * S + B * T) * 4 + B * T * 8 + 512;
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
  float* off = ab + B * T * S;
  double* off_a = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(off + B * T) + 255) & ~uintptr_t(255));
  ctc_gather_kernel<<<dim3((unsigned)T, (unsigned)B), 256, 0, stream>>>(logp, sb, st, targets, tgt_stride, (int)T, (int)L,
                                                                         (int)V, blank, E);
  CN_CHECK_LAUNCH();
  AbArgs p;
  p.E = E; p.targets = targets; p.tgt_stride = tgt_stride; p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.T = (int)T; p.L = (int)L; p.ab = ab; p.nll = nll; p.off_a = off_a; p.off = off;
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
    g.tgt_len = target_lengths; g.ab = ab; g.nll = nll; g.off = off; g.T = (int)T; g.L = (int)L; g.V = (int)V; g.S = S; g.B = (int)B;
    g.blank = blank; g.grad_out = grad_out; g.grad = grad_logits; g.gsb = gsb; g.gst = gst;
    ctc_grad_kernel<<<dim3((unsigned)T, (unsigned)B), 256, 0, stream>>>(g);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}
