// K3: CTC forced alignment (Viterbi) — bit-exact restatement of CassNAT.viterbi_align
// (src/models/cassnat.py:272-353) as one CTA per utterance:
//   warp 0      : the max-plus DP, lattice states strided over lanes (SPL consecutive states
//                 per lane, alpha in registers, s-1 / s-2 neighbours by warp shuffle),
//                 2-bit back-pointers packed 16 states / word and streamed to the workspace;
//   warps 1..7  : gather the (U+1) emissions each frame needs out of the [T,V] posterior rows
//                 into a double-buffered shared-memory ring, one chunk of frames ahead of the DP.
// Back-trace re-stages the back-pointer words through shared memory chunk by chunk, then the
// whole CTA maps states -> labels, collapses repeats and shifts by one frame.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kFrames = 32;          // frames per emission chunk
constexpr int kBtBytes = 8192;       // back-trace staging bytes (fits inside the emission ring)

struct VitArgs {
  const float* logp; int64_t sb, st;
  const float* emis; int64_t eb, et;   // compact emissions [B,T,U+1] (slot 0 = blank, slot u = label u) instead of logp
  const uint8_t* src_mask; const int64_t* src_size;
  const int64_t* ys; int64_t ys_stride; const int64_t* ylens;
  int T, U, V; int64_t blank;
  const int32_t* t_sample; int n_sample;
  int64_t* aligned; uint32_t* bp;
  int n_sm;                            // SM count: co-resident CTAs put their DP warp on different schedulers
};

template <int SPL>
__global__ void __launch_bounds__(kThreads) viterbi_kernel(VitArgs p) {
  constexpr int NWORD = (SPL + 15) / 16;
  constexpr int EW = 16 * SPL + 1;               // emission row: [0]=blank, [1..U]=labels, rest zero
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* emis = reinterpret_cast<float*>(smem_raw);                       // [2][kFrames][EW]
  int* cols = reinterpret_cast<int*>(emis + 2 * kFrames * EW);            // [EW] vocab column per emission slot
  unsigned short* states = reinterpret_cast<unsigned short*>(cols + EW);  // [T]
  __shared__ float fin[32 * SPL];
  __shared__ int samp[16];

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // The DP is one serial warp per utterance.  CTAs are handed out one per SM and then wrap around, so the k-th CTA
  // of an SM runs its DP on warp k (scheduler k): two co-resident utterances never share an issue port.
  const int dp_warp = ((int)blockIdx.x / p.n_sm) & 3;
  const int T = p.T, U = p.U;
  const int S = 2 * U + 1;
  int xb = (int)p.src_size[b];
  xb = xb < 0 ? 0 : (xb > T ? T : xb);
  const int ylen_b = (int)p.ylens[b];
  const int path_len = 2 * ylen_b + 1;
  const float* logp_b = p.emis ? p.emis + (int64_t)b * p.eb : p.logp + (int64_t)b * p.sb;
  const int64_t row_pitch = p.emis ? p.et : p.st;
  const uint8_t* mask_b = p.src_mask + (int64_t)b * T;
  uint32_t* bp_b = p.bp + (int64_t)b * T * NWORD * 32;

  for (int u = tid; u < EW; u += kThreads) {
    int64_t col = (u == 0 || u > U) ? p.blank : p.ys[(int64_t)b * p.ys_stride + (u - 1)];
    // the reference would raise on an out-of-vocabulary label; stay memory-safe instead
    cols[u] = (int)(col < 0 ? 0 : (col >= p.V ? p.V - 1 : col));
  }
  if (tid < 16) samp[tid] = (tid < p.n_sample) ? p.t_sample[tid] : -1;
  __syncthreads();

  auto fill = [&](int chunk, int buf, bool all_warps) {
    const int t0 = chunk * kFrames;
    const int nf = min(kFrames, xb - t0);
    float* e = emis + buf * kFrames * EW;
    const int n_warps = all_warps ? kThreads / 32 : kThreads / 32 - 1;
    const int fw = all_warps ? warp : (warp < dp_warp ? warp : warp - 1);     // rank among the filling warps
    for (int f = fw; f < nf; f += n_warps) {
      const bool valid = mask_b[t0 + f] != 0;
      const float* row = logp_b + (int64_t)(t0 + f) * row_pitch;
      for (int u0 = 0; u0 < EW; u0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int u = u0 + k * 32 + lane;
          // compact emissions: slot u of the row, coalesced; posteriors: the label's vocabulary column (a gather)
          v[k] = (u <= U) ? (valid ? cn_ld_stream_f32(row + (p.emis ? u : cols[u])) : CN_LOGZERO) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int u = u0 + k * 32 + lane;
          if (u < EW) e[f * EW + u] = v[k];
        }
      }
    }
  };

  // ---- DP warp state
  float a[SPL];
  uint32_t skip_ok = 0, inside = 0;
  if (warp == dp_warp) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = lane * SPL + j;
      a[j] = (s == 0) ? 0.f : CN_LOGZERO;
      // odd states carry labels; a skip s-2 -> s is allowed only between different labels
      if ((j & 1) && s >= 3 && s < S) {
        if (cols[(s + 1) >> 1] != cols[(s - 1) >> 1]) skip_ok |= 1u << j;
      }
      if (s < path_len && s < S) inside |= 1u << j;
    }
  }

  const int n_chunks = (xb + kFrames - 1) / kFrames;
  if (n_chunks > 0) fill(0, 0, true);
  __syncthreads();

  for (int c = 0; c < n_chunks; ++c) {
    if (warp == dp_warp) {
      const int t0 = c * kFrames;
      const int nf = min(kFrames, xb - t0);
      const float* e = emis + (c & 1) * kFrames * EW;
      for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        const float* ef = e + f * EW;
        const float eb = ef[0];
        float el[SPL / 2];
#pragma unroll
        for (int j = 0; j < SPL / 2; ++j) el[j] = ef[lane * (SPL / 2) + j + 1];
        bool second = false;
        if (p.n_sample > 0) {
          for (int i = 0; i < p.n_sample && i < 16; ++i) second |= (samp[i] == t);
        }
        float up1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
        float up2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
        if (lane == 0) { up1 = CN_LOGZERO; up2 = CN_LOGZERO; }
        float na[SPL];
        uint32_t words[NWORD];
#pragma unroll
        for (int w = 0; w < NWORD; ++w) words[w] = 0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          const float c0 = a[j];
          const float c1 = (j >= 1) ? a[j - 1] : up1;
          float c2 = (j >= 2) ? a[j - 2] : (j == 1 ? up1 : up2);
          if (!((skip_ok >> j) & 1u)) c2 = CN_LOGZERO;
          float best = c0;
          int k = 0;
          if (c1 > best) { best = c1; k = 1; }
          if (c2 > best) { best = c2; k = 2; }
          if (second && (lane != 0 || j != 0)) {
            // 2nd entry of a stable descending sort of (c0,c1,c2): cassnat.py:321-326
            float vi, vj; int ii, jj;
            if (k == 0) { vi = c1; ii = 1; vj = c2; jj = 2; }
            else if (k == 1) { vi = c0; ii = 0; vj = c2; jj = 2; }
            else { vi = c0; ii = 0; vj = c1; jj = 1; }
            if (vj > vi) { best = vj; k = jj; } else { best = vi; k = ii; }
          }
          if (!((inside >> j) & 1u)) best = CN_LOGZERO;
          words[j >> 4] |= (uint32_t)k << (2 * (j & 15));
          na[j] = best + ((j & 1) ? el[j >> 1] : eb);
        }
#pragma unroll
        for (int j = 0; j < SPL; ++j) a[j] = na[j];
#pragma unroll
        for (int w = 0; w < NWORD; ++w) bp_b[((int64_t)t * NWORD + w) * 32 + lane] = words[w];
      }
    } else if (c + 1 < n_chunks) {
      fill(c + 1, (c + 1) & 1, false);
    }
    __syncthreads();
  }

  // ---- termination (cassnat.py:337-339): strict '>' between the last two path states at t = xb
  if (warp == dp_warp) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) fin[lane * SPL + j] = a[j];
  }
  __threadfence_block();
  __syncthreads();
  int state = 0;
  if (tid == 0 && xb > 0) {
    int i1 = path_len - 1, i2 = path_len - 2;
    if (i1 < 0) i1 += S;
    if (i2 < 0) i2 += S;                      // python negative-index wrap (ylens == 0)
    i1 = min(max(i1, 0), S - 1);
    i2 = min(max(i2, 0), S - 1);
    state = (fin[i1] > fin[i2]) ? i1 : i2;
    states[xb - 1] = (unsigned short)state;
  }
  // ---- back-trace: stage back-pointer words for kBt frames at a time (device-wide visible
  // because the same CTA wrote them; __syncthreads orders the global writes for this CTA)
  constexpr int kBt = kBtBytes / (NWORD * 128);
  uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw);
  for (int hi = xb - 1; hi >= 1; hi -= kBt) {
    const int lo = max(hi - kBt + 1, 1);          // frames lo..hi are consumed in this round
    const int n = (hi - lo + 1) * NWORD * 32;
    for (int i = tid; i < n; i += kThreads) stage[i] = bp_b[(int64_t)lo * NWORD * 32 + i];
    __syncthreads();
    if (tid == 0) {
      for (int t = hi; t >= lo; --t) {
        const int ln = state / SPL, j = state - ln * SPL;
        const uint32_t w = stage[((t - lo) * NWORD + (j >> 4)) * 32 + ln];
        state -= (int)((w >> (2 * (j & 15))) & 3u);
        states[t - 1] = (unsigned short)state;
      }
    }
    __syncthreads();
  }
  __syncthreads();

  // ---- states -> labels -> collapse repeats -> shift by one frame (cassnat.py:344-350)
  auto label = [&](int t) -> int64_t {
    if (t < 0) return 0;
    if (t >= xb) return p.blank;
    const int s = states[t];
    return (s & 1) ? (int64_t)cols[(s + 1) >> 1] : p.blank;
  };
  int64_t* out = p.aligned + (int64_t)b * T;
  for (int t = tid; t < T; t += kThreads) {
    int64_t v = 0;
    if (t > 0) {
      const int64_t cur = label(t - 1), prev = label(t - 2);
      v = (cur != prev) ? cur : 0;
    }
    out[t] = v;
  }
}

template <int SPL>
int launch(const VitArgs& p, int64_t B, cudaStream_t stream) {
  constexpr int EW = 16 * SPL + 1;
  size_t smem = (size_t)2 * kFrames * EW * sizeof(float) + (size_t)EW * sizeof(int) +
                (size_t)((p.T + 1) / 2 * 2) * sizeof(unsigned short) + 16;
  if (smem < (size_t)kBtBytes) smem = kBtBytes;
  if (smem > 200 * 1024) return CN_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(viterbi_kernel<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  viterbi_kernel<SPL><<<(unsigned)B, kThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

int spl_for(int64_t U) {
  const int64_t S = 2 * U + 1;
  int spl = (int)((S + 31) / 32);
  if (spl < 2) spl = 2;
  if (spl & 1) ++spl;
  if (spl > 16) spl = ((spl + 3) / 4) * 4;
  return spl;
}

}  // namespace

extern "C" int64_t cn_viterbi_workspace_bytes(int64_t B, int64_t T, int64_t U) {
  if (B < 0 || T < 0 || U < 1 || U > 511) return -1;
  const int spl = spl_for(U);
  const int nword = (spl + 15) / 16;
  return B * T * nword * 32 * 4;
}

extern "C" int cn_viterbi_align(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                                const uint8_t* src_mask, const int64_t* src_size, const int64_t* ys, int64_t U,
                                int64_t ys_stride, const int64_t* ylens, int64_t blank, const int32_t* t_sample,
                                int64_t n_sample, int64_t* aligned, void* workspace, int64_t workspace_bytes,
                                cn_stream_t stream_) {
  if (!logp || !src_mask || !src_size || !ys || !ylens || !aligned || !workspace) return CN_ERR_BAD_ARG;
  if (B < 0 || T <= 0 || V <= 0 || ys_stride < U || n_sample < 0) return CN_ERR_BAD_ARG;
  if (U < 1 || U > 511 || T > 65535 || n_sample > 16) return CN_ERR_UNSUPPORTED;
  if (n_sample > 0 && !t_sample) return CN_ERR_BAD_ARG;
  if (workspace_bytes < cn_viterbi_workspace_bytes(B, T, U)) return CN_ERR_WORKSPACE;
  if (B == 0) return CN_OK;
  VitArgs p;
  p.emis = nullptr; p.eb = 0; p.et = 0;
  p.logp = logp; p.sb = sb; p.st = st; p.src_mask = src_mask; p.src_size = src_size;
  p.ys = ys; p.ys_stride = ys_stride; p.ylens = ylens; p.T = (int)T; p.U = (int)U; p.V = (int)V; p.blank = blank;
  p.t_sample = t_sample; p.n_sample = (int)n_sample; p.aligned = aligned; p.bp = (uint32_t*)workspace;
  { int dev = 0, n = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); p.n_sm = n > 0 ? n : 148; }
  cudaStream_t stream = (cudaStream_t)stream_;
  switch (spl_for(U)) {
    case 2: return launch<2>(p, B, stream);
    case 4: return launch<4>(p, B, stream);
    case 6: return launch<6>(p, B, stream);
    case 8: return launch<8>(p, B, stream);
    case 10: return launch<10>(p, B, stream);
    case 12: return launch<12>(p, B, stream);
    case 14: return launch<14>(p, B, stream);
    case 16: return launch<16>(p, B, stream);
    case 20: return launch<20>(p, B, stream);
    case 24: return launch<24>(p, B, stream);
    case 28: return launch<28>(p, B, stream);
    case 32: return launch<32>(p, B, stream);
    default: return CN_ERR_UNSUPPORTED;
  }
}

// ---------------------------------------------------------------- compact emissions
// The forced alignment only ever looks at U+1 of the V posteriors of a frame.  With dumped posteriors the gather is
// the cost (one 64-byte DRAM atom per 4-byte value); when the caller can hand over (or produce once and share with
// the CTC loss) the compact [B,T,U+1] emissions, the dynamic program reads 4*(U+1) coalesced bytes per frame.
namespace {
__global__ void __launch_bounds__(256) gather_emissions_kernel(const float* __restrict__ logp, int64_t sb, int64_t st, int T, int V,
                                                               const int64_t* __restrict__ ys, int64_t ys_stride, int U, int64_t blank,
                                                               float* __restrict__ emis) {
  extern __shared__ int gcols[];                               // [U+1] vocabulary column per slot
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  for (int u = threadIdx.x; u <= U; u += 256) {
    int64_t col = u == 0 ? blank : ys[(int64_t)b * ys_stride + (u - 1)];
    gcols[u] = (int)(col < 0 ? 0 : (col >= V ? V - 1 : col));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = warp; f < 32 && t0 + f < T; f += 8) {           // a warp per frame, 4 frames per warp
    const float* row = logp + (int64_t)b * sb + (int64_t)(t0 + f) * st;
    float* out = emis + ((int64_t)b * T + t0 + f) * (U + 1);
    for (int u0 = 0; u0 <= U; u0 += 32 * 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int u = u0 + k * 32 + lane;
        v[k] = u <= U ? cn_ld_stream_f32(row + gcols[u]) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int u = u0 + k * 32 + lane;
        if (u <= U) out[u] = v[k];
      }
    }
  }
}
}  // namespace

extern "C" int cn_gather_emissions(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* ys,
                                   int64_t U, int64_t ys_stride, int64_t blank, float* emis, cn_stream_t stream_) {
  if (!logp || !ys || !emis || B < 0 || T <= 0 || V <= 0 || U < 1 || ys_stride < U) return CN_ERR_BAD_ARG;
  if (B > 65535 || U > 8191) return CN_ERR_UNSUPPORTED;
  if (B == 0) return CN_OK;
  dim3 grid((unsigned)((T + 31) / 32), (unsigned)B);
  gather_emissions_kernel<<<grid, 256, (size_t)(U + 1) * sizeof(int), (cudaStream_t)stream_>>>(logp, sb, st, (int)T, (int)V, ys, ys_stride,
                                                                                             (int)U, blank, emis);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_viterbi_align_compact(const float* emis, int64_t B, int64_t T, int64_t V, int64_t eb, int64_t et,
                                        const uint8_t* src_mask, const int64_t* src_size, const int64_t* ys, int64_t U,
                                        int64_t ys_stride, const int64_t* ylens, int64_t blank, const int32_t* t_sample,
                                        int64_t n_sample, int64_t* aligned, void* workspace, int64_t workspace_bytes,
                                        cn_stream_t stream_) {
  if (!emis || !src_mask || !src_size || !ys || !ylens || !aligned || !workspace) return CN_ERR_BAD_ARG;
  if (B < 0 || T <= 0 || V <= 0 || ys_stride < U || n_sample < 0 || et < U + 1) return CN_ERR_BAD_ARG;
  if (U < 1 || U > 511 || T > 65535 || n_sample > 16) return CN_ERR_UNSUPPORTED;
  if (n_sample > 0 && !t_sample) return CN_ERR_BAD_ARG;
  if (workspace_bytes < cn_viterbi_workspace_bytes(B, T, U)) return CN_ERR_WORKSPACE;
  if (B == 0) return CN_OK;
  VitArgs p;
  p.emis = emis; p.eb = eb; p.et = et;
  p.logp = nullptr; p.sb = 0; p.st = 0; p.src_mask = src_mask; p.src_size = src_size;
  p.ys = ys; p.ys_stride = ys_stride; p.ylens = ylens; p.T = (int)T; p.U = (int)U; p.V = (int)V; p.blank = blank;
  p.t_sample = t_sample; p.n_sample = (int)n_sample; p.aligned = aligned; p.bp = (uint32_t*)workspace;
  { int dev = 0, n = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); p.n_sm = n > 0 ? n : 148; }
  cudaStream_t stream = (cudaStream_t)stream_;
  switch (spl_for(U)) {
    case 2: return launch<2>(p, B, stream);
    case 4: return launch<4>(p, B, stream);
    case 6: return launch<6>(p, B, stream);
    case 8: return launch<8>(p, B, stream);
    case 10: return launch<10>(p, B, stream);
    case 12: return launch<12>(p, B, stream);
    case 14: return launch<14>(p, B, stream);
    case 16: return launch<16>(p, B, stream);
    case 20: return launch<20>(p, B, stream);
    case 24: return launch<24>(p, B, stream);
    case 28: return launch<28>(p, B, stream);
    case 32: return launch<32>(p, B, stream);
    default: return CN_ERR_UNSUPPORTED;
  }
}
