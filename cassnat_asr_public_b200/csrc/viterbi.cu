// K3: CTC forced alignment (Viterbi) — bit-exact restatement of CassNAT.viterbi_align
// (src/models/cassnat.py:272-353) as one CTA per utterance:
//   DP warps    : the max-plus DP.  The 2U+1 lattice states are laid out SPL consecutive states per lane over
//                 NW x 32 lanes (NW = 4 warps, one per scheduler, once the lattice has more than 128 states: a
//                 single warp spends ~16 issue slots per state and frame, so the serial chain per frame is what
//                 bounds the kernel); alpha lives in registers, the s-1 / s-2 neighbours come by warp shuffle and,
//                 across the warp seams, through a double-buffered shared-memory slot + one named barrier per
//                 frame.  2-bit back-pointers are packed per lane over FPW frames and streamed to the workspace;
//   fill warps  : gather the (U+1) emissions each frame needs out of the [T,V] posterior rows (or copy the compact
//                 [T,U+1] rows) into a double-buffered shared-memory ring, one chunk of frames ahead of the DP.
// The ring is small enough for several utterances per SM, so that a whole cfg2 batch is resident in one wave.
// Back-trace re-stages the back-pointer words through shared memory chunk by chunk, then the
// whole CTA maps states -> labels, collapses repeats and shifts by one frame.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kFrames = 16;          // frames per emission chunk
constexpr int kRing = 3;             // chunks in the shared-memory emission ring
constexpr int kBtBytes = 8192;       // back-trace staging bytes (overlays the emission ring)

struct VitArgs {
  const float* logp; int64_t sb, st;
  const float* emis; int64_t eb, et;   // compact emissions [B,T,U+1] (slot 0 = blank, slot u = label u) instead of logp
  const uint8_t* src_mask; const int64_t* src_size;
  const int64_t* ys; int64_t ys_stride; const int64_t* ylens;
  int T, U, V; int64_t blank;
  const int32_t* t_sample; int n_sample;
  int64_t* aligned; uint32_t* bp;
  int n_sm;                            // SM count: co-resident CTAs put a single DP warp on different schedulers
};

struct cn_true { static constexpr bool value = true; };
struct cn_false { static constexpr bool value = false; };

template <int SPL, int NW>
struct VitShape {
  static constexpr int LANES = NW * 32;
  static constexpr int BITS = 2 * SPL;             // back-pointer bits per lane and frame
  static constexpr int FPW = 32 / BITS;            // frames packed per 32-bit word
  static constexpr int EW = LANES * SPL / 2 + 1;   // emission row: [0]=blank, [1..U]=labels, rest zero
  static constexpr int RING = (kRing * kFrames * EW * 4 > kBtBytes) ? kRing * kFrames * EW * 4 : kBtBytes;
};

template <int SPL, int NW>
__global__ void __launch_bounds__(kThreads) viterbi_kernel(VitArgs p) {
  using Sh = VitShape<SPL, NW>;
  constexpr int LANES = Sh::LANES, BITS = Sh::BITS, FPW = Sh::FPW, EW = Sh::EW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* emis = reinterpret_cast<float*>(smem_raw);                       // [kRing][kFrames][EW]
  int* cols = reinterpret_cast<int*>(smem_raw + Sh::RING);                // [EW] vocab column per emission slot
  unsigned short* states = reinterpret_cast<unsigned short*>(cols + EW);  // [T]
  uint8_t* maskb = reinterpret_cast<uint8_t*>(states + (p.T + 1) / 2 * 2);  // [T] frame mask of this utterance
  __shared__ float fin[LANES * SPL];
  __shared__ float seam[2][NW][2];     // alpha of the last two states of a DP warp, for its right-hand neighbour
  __shared__ int samp[16];

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // one DP warp: CTAs are handed out one per SM and then wrap around, so the k-th CTA of an SM runs its DP on warp k
  // (scheduler k) and co-resident utterances do not share an issue port; four DP warps: warps 0..3 = schedulers 0..3
  const int dp_warp = ((int)blockIdx.x / p.n_sm) & 3;
  const int w = NW == 1 ? (warp == dp_warp ? 0 : -1) : (warp < NW ? warp : -1);        // rank among the DP warps
  const int fill_rank = NW == 1 ? (warp < dp_warp ? warp : warp - 1) : warp - NW;
  const int g = (w < 0 ? 0 : w) * 32 + lane;                                           // lattice lane
  const int T = p.T, U = p.U;
  const int S = 2 * U + 1;
  int xb = (int)p.src_size[b];
  xb = xb < 0 ? 0 : (xb > T ? T : xb);
  const int ylen_b = (int)p.ylens[b];
  const int path_len = 2 * ylen_b + 1;
  const float* logp_b = p.emis ? p.emis + (int64_t)b * p.eb : p.logp + (int64_t)b * p.sb;
  const int64_t row_pitch = p.emis ? p.et : p.st;
  const uint8_t* mask_b = p.src_mask + (int64_t)b * T;
  uint32_t* bp_b = p.bp + (int64_t)b * ((T + FPW - 1) / FPW) * LANES;

  for (int u = tid; u < EW; u += kThreads) {
    int64_t col = (u == 0 || u > U) ? p.blank : p.ys[(int64_t)b * p.ys_stride + (u - 1)];
    // the reference would raise on an out-of-vocabulary label; stay memory-safe instead
    cols[u] = (int)(col < 0 ? 0 : (col >= p.V ? p.V - 1 : col));
  }
  if (tid < 16) samp[tid] = (tid < p.n_sample) ? p.t_sample[tid] : -1;
  if (tid < 2 * NW * 2) (&seam[0][0][0])[tid] = CN_LOGZERO;
  __syncthreads();

  // Emission staging into a ring of kRing chunks, two chunks ahead of the DP; a fill warp takes whole frames.
  //   compact rows : asynchronous global -> shared copies (cp.async, coalesced, no registers held): while the DP
  //                  consumes chunk c the fill warps issue chunk c+2 and wait for chunk c+1, whose DRAM latency was
  //                  paid during the previous round; frames masked out inside the valid length (not produced by the
  //                  reference's prefix masks, but defined) are patched to LOGZERO by the copying thread;
  //   posteriors   : the (U+1) columns of a frame are a gather -- streaming loads (no L1 allocation), eight per lane
  //                  in flight, then the shared-memory stores.
  constexpr int N_FILL = kThreads / 32 - NW;
  const int U1 = U + 1;
  const bool compact = p.emis != nullptr;
  auto copy_async = [&](int chunk, int n_warps, int first_warp) {
    const int t0 = chunk * kFrames, nf = min(kFrames, xb - t0);
    const uint32_t e = (uint32_t)__cvta_generic_to_shared(emis + (chunk % kRing) * kFrames * EW);
    for (int f = first_warp; f < nf; f += n_warps) {
      const float* row = logp_b + (int64_t)(t0 + f) * row_pitch;
      for (int u = lane; u < U1; u += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(e + 4u * (uint32_t)(f * EW + u)), "l"(row + u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto patch = [&](int chunk, int n_warps, int first_warp) {
    const int t0 = chunk * kFrames, nf = min(kFrames, xb - t0);
    float* e = emis + (chunk % kRing) * kFrames * EW;
    for (int f = first_warp; f < nf; f += n_warps) {
      if (maskb[t0 + f]) continue;
      for (int u = lane; u < U1; u += 32) e[f * EW + u] = CN_LOGZERO;
    }
  };
  constexpr int FW = (kFrames + N_FILL - 1) / N_FILL;       // frames per fill warp and chunk: all gathered at once
  auto gather = [&](int chunk, int n_warps, int first_warp) {
    const int t0 = chunk * kFrames, nf = min(kFrames, xb - t0);
    float* e = emis + (chunk % kRing) * kFrames * EW;
    for (int u0 = 0; u0 < U1; u0 += 32 * 8) {
      float v[FW][8];
      int col[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) col[k] = cols[min(u0 + k * 32 + lane, U1 - 1)];
#pragma unroll
      for (int i = 0; i < FW; ++i) {
        const int f = first_warp + i * n_warps;
        const bool valid = f < nf && maskb[t0 + min(f, nf - 1)] != 0;
        const float* row = logp_b + (int64_t)(t0 + min(f, nf - 1)) * row_pitch;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[i][k] = (u0 + k * 32 + lane < U1 && valid) ? cn_ld_stream_f32(row + col[k]) : CN_LOGZERO;
      }
#pragma unroll
      for (int i = 0; i < FW; ++i) {
        const int f = first_warp + i * n_warps;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int u = u0 + k * 32 + lane;
          if (f < nf && u < U1) e[f * EW + u] = v[i][k];
        }
      }
    }
  };
  for (int i = tid; i < kRing * kFrames * EW; i += kThreads) emis[i] = 0.f;      // (slots beyond U stay zero)
  for (int t = tid; t < xb; t += kThreads) maskb[t] = mask_b[t];
  __syncthreads();
  const int n_chunks = (xb + kFrames - 1) / kFrames;
  if (compact) {
    if (n_chunks > 0) copy_async(0, kThreads / 32, warp);
    if (w < 0 && n_chunks > 1) copy_async(1, N_FILL, fill_rank);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (n_chunks > 0) patch(0, kThreads / 32, warp);
    if (w < 0 && n_chunks > 1) patch(1, N_FILL, fill_rank);
  } else {
    if (n_chunks > 0) gather(0, kThreads / 32, warp);
    if (w < 0 && n_chunks > 1) gather(1, N_FILL, fill_rank);
  }

  // ---- DP state of this lane.  States at or beyond the path length 2*ylen+1 are not masked here: a state only
  // ever feeds states above it, termination and back-trace stay below the path length, so whatever those lanes hold
  // never reaches the result (the reference's masked value is reconstructed at termination where python's negative
  // index reads one, ylens == 0).
  float a[SPL];
  uint32_t skip_ok = 0;
  if (w >= 0) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = g * SPL + j;
      a[j] = (s == 0) ? 0.f : CN_LOGZERO;
      // odd states carry labels; a skip s-2 -> s is allowed only between different labels
      if ((j & 1) && s >= 3 && s < S) {
        if (cols[(s + 1) >> 1] != cols[(s - 1) >> 1]) skip_ok |= 1u << j;
      }
    }
  }
  const float* seam_in = &seam[1][w > 0 ? w - 1 : 0][0];       // parity (t & 1) ^ 1 at t = 0
  float* seam_out = &seam[0][w > 0 ? w : 0][0];
  const int seam_flip = NW * 2;                                // floats between the two parities
  const bool seam_rd = NW > 1 && w > 0 && lane == 0, seam_wr = NW > 1 && w >= 0 && w < NW - 1 && lane == 31;
  __syncthreads();

  uint32_t acc = 0;                    // back-pointers of this lane for the current group of FPW frames
  // one frame of the dynamic program; SECOND: frames of a sampled alignment take the runner-up predecessor
  auto frame = [&](auto second_tag, int t, const float* ef, bool odd_frame) {
    constexpr bool SECOND = decltype(second_tag)::value;
    const float eb = ef[0];
    float el[SPL / 2];
#pragma unroll
    for (int j = 0; j < SPL / 2; ++j) el[j] = ef[g * (SPL / 2) + j + 1];
    float up1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
    float up2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
    if (lane == 0) {
      up1 = CN_LOGZERO;
      up2 = CN_LOGZERO;
      if (seam_rd) {
        const float* si = odd_frame ? seam_in - seam_flip : seam_in;
        up1 = si[0];
        up2 = si[1];
      }
    }
    float na[SPL];
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const float c0 = a[j];
      const float c1 = (j >= 1) ? a[j - 1] : up1;
      float best = c0;
      uint32_t k = 0;
      if (!(j & 1) && !SECOND) {
        // blank state: stay or advance (the skip never applies)
        k = c1 > c0 ? 1u : 0u;
        best = c1 > c0 ? c1 : c0;
      } else {
        float c2 = (j >= 2) ? a[j - 2] : (j == 1 ? up1 : up2);
        if (!((skip_ok >> j) & 1u)) c2 = CN_LOGZERO;
        if (c1 > best) { best = c1; k = 1; }
        if (c2 > best) { best = c2; k = 2; }
        if (SECOND && (g != 0 || j != 0)) {
          // 2nd entry of a stable descending sort of (c0,c1,c2): cassnat.py:321-326
          float vi, vj; uint32_t ii, jj;
          if (k == 0) { vi = c1; ii = 1; vj = c2; jj = 2; }
          else if (k == 1) { vi = c0; ii = 0; vj = c2; jj = 2; }
          else { vi = c0; ii = 0; vj = c1; jj = 1; }
          if (vj > vi) { best = vj; k = jj; } else { best = vi; k = ii; }
        }
      }
      bits |= k << (2 * j);
      na[j] = best + ((j & 1) ? el[j >> 1] : eb);
    }
#pragma unroll
    for (int j = 0; j < SPL; ++j) a[j] = na[j];
    if (seam_wr) {
      float* so = odd_frame ? seam_out + seam_flip : seam_out;
      so[0] = na[SPL - 1];
      so[1] = na[SPL - 2];
    }
    const int slot = t % FPW;
    acc |= bits << (slot * BITS);
    if (slot == FPW - 1 || t == xb - 1) {
      bp_b[(int64_t)(t / FPW) * LANES + g] = acc;
      acc = 0;
    }
    if (NW > 1) asm volatile("bar.sync 2, %0;" ::"n"(LANES) : "memory");     // the seam values of frame t are visible
  };

  for (int c = 0; c < n_chunks; ++c) {
    if (w >= 0) {
      const int t0 = c * kFrames;
      const int nf = min(kFrames, xb - t0);
      const float* e = emis + (c % kRing) * kFrames * EW;
      uint32_t second_mask = 0;        // frames of this chunk that take the runner-up predecessor (sampled alignments)
      for (int i = 0; i < p.n_sample && i < 16; ++i) {
        const int d = samp[i] - t0;
        if (d >= 0 && d < nf) second_mask |= 1u << d;
      }
      if (second_mask == 0) {
#pragma unroll 2
        for (int f = 0; f < nf; ++f) frame(cn_false{}, t0 + f, e + f * EW, (f & 1) != 0);
      } else {
        for (int f = 0; f < nf; ++f) {
          if ((second_mask >> f) & 1u) frame(cn_true{}, t0 + f, e + f * EW, (f & 1) != 0);
          else frame(cn_false{}, t0 + f, e + f * EW, (f & 1) != 0);
        }
      }
    } else if (compact) {
      if (c + 2 < n_chunks) {
        copy_async(c + 2, N_FILL, fill_rank);
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // chunk c+1 (issued a round ago) has landed
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      if (c + 1 < n_chunks && c > 0) patch(c + 1, N_FILL, fill_rank);
    } else if (c + 2 < n_chunks) {
      gather(c + 2, N_FILL, fill_rank);
    }
    __syncthreads();
  }

  // ---- termination (cassnat.py:337-339): strict '>' between the last two path states at t = xb
  if (w >= 0) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) fin[g * SPL + j] = a[j];
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
    // alpha of state q after the last frame; beyond the path length: the reference's masked value LOGZERO + emission
    auto final_alpha = [&](int q) -> float {
      if (q < path_len) return fin[q];
      float em = CN_LOGZERO;
      if (maskb[xb - 1]) {
        const int u = (q & 1) ? (q + 1) >> 1 : 0;
        em = logp_b[(int64_t)(xb - 1) * row_pitch + (p.emis ? u : cols[u])];
      }
      return CN_LOGZERO + em;
    };
    state = (final_alpha(i1) > final_alpha(i2)) ? i1 : i2;
    states[xb - 1] = (unsigned short)state;
  }
  // ---- back-trace: stage the back-pointer words of kBtG frame groups at a time (device-wide visible
  // because the same CTA wrote them; __syncthreads orders the global writes for this CTA)
  constexpr int kBtG = kBtBytes / (LANES * 4);
  uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw);
  for (int ghi = (xb - 1) / FPW; ghi >= 0 && xb > 1; ghi -= kBtG) {
    const int glo = max(ghi - kBtG + 1, 0);       // frame groups glo..ghi are consumed in this round
    const int n = (ghi - glo + 1) * LANES;
    for (int i = tid; i < n; i += kThreads) stage[i] = bp_b[(int64_t)glo * LANES + i];
    __syncthreads();
    if (tid == 0) {
      const int thi = min(xb - 1, ghi * FPW + FPW - 1), tlo = max(glo * FPW, 1);
      for (int t = thi; t >= tlo; --t) {
        const int ln = state / SPL, j = state - ln * SPL;
        const uint32_t wd = stage[(t / FPW - glo) * LANES + ln];
        state -= (int)((wd >> ((t % FPW) * BITS + 2 * j)) & 3u);
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

template <int SPL, int NW>
int launch(const VitArgs& p, int64_t B, cudaStream_t stream) {
  using Sh = VitShape<SPL, NW>;
  size_t smem = (size_t)Sh::RING + (size_t)Sh::EW * sizeof(int) + (size_t)((p.T + 1) / 2 * 2) * sizeof(unsigned short) + (size_t)p.T + 16;
  if (smem > 200 * 1024) return CN_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(viterbi_kernel<SPL, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  viterbi_kernel<SPL, NW><<<(unsigned)B, kThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

// lattice shape for U labels: states per lane and DP warps
void shape_for(int64_t U, int& spl, int& nw) {
  const int64_t S = 2 * U + 1;
  if (S <= 64) { spl = 2; nw = 1; }
  else if (S <= 128) { spl = 4; nw = 1; }
  else { nw = 4; spl = 2 * (int)((S + 255) / 256); }
}

int dispatch(const VitArgs& p, int64_t B, int64_t U, cudaStream_t stream) {
  int spl, nw;
  shape_for(U, spl, nw);
  if (nw == 1) return spl == 2 ? launch<2, 1>(p, B, stream) : launch<4, 1>(p, B, stream);
  switch (spl) {
    case 2: return launch<2, 4>(p, B, stream);
    case 4: return launch<4, 4>(p, B, stream);
    case 6: return launch<6, 4>(p, B, stream);
    case 8: return launch<8, 4>(p, B, stream);
    default: return CN_ERR_UNSUPPORTED;
  }
}

}  // namespace

extern "C" int64_t cn_viterbi_workspace_bytes(int64_t B, int64_t T, int64_t U) {
  if (B < 0 || T < 0 || U < 1 || U > 511) return -1;
  int spl, nw;
  shape_for(U, spl, nw);
  const int fpw = 32 / (2 * spl);
  return B * ((T + fpw - 1) / fpw) * nw * 32 * 4;
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
  return dispatch(p, B, U, stream);
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
  return dispatch(p, B, U, stream);
}
