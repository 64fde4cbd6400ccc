// K3: CTC forced alignment (Viterbi) — bit-exact restatement of CassNAT.viterbi_align
// (src/models/cassnat.py:272-353) as one CTA per utterance:
//   DP warps    : the max-plus DP.  The 2U+1 lattice states are laid out SPL consecutive states per lane over
//                 NW x 32 lanes (NW = 4 warps, one per scheduler, once the lattice has more than 128 states: a
//                 single warp spends ~13 issue slots per state and frame, and the serial chain per frame is what
//                 bounds the kernel).  alpha lives in registers, the s-1 / s-2 neighbours come by warp shuffle.
//                 The warps run SKEWED by one chunk of kFrames frames: in round r warp w works on chunk r - w, and
//                 takes the alpha values that cross the warp seam from what warp w-1 left in shared memory for that
//                 chunk one round earlier -- no per-frame synchronisation, one CTA barrier per round.
//                 2-bit back-pointers are packed per lane over FPW frames and streamed to the workspace;
//   fill warps  : stage, two chunks ahead of each DP warp, the emissions that warp needs (its slice of label slots
//                 + the blank) out of the [T,V] posterior rows (a gather) or the compact [T,U+1] rows (cp.async)
//                 into per-warp shared-memory rings.
// The rings are small enough for two utterances per SM, so that a whole cfg2 batch is resident in one wave.
// Back-trace re-stages the back-pointer words through shared memory, then the whole CTA maps states -> labels,
// collapses repeats and shifts by one frame.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
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

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

// (x & m) | (LOGZERO & ~m): a select on a loop-invariant condition without spending a predicate register -- with the
// handful the ISA has, the compiler otherwise funnels every compare of the frame through the same one or two
// predicates and the seven compare/select pairs of a frame run strictly one after the other
__device__ __forceinline__ float keep_or_logzero(float x, uint32_t m) {
  return __int_as_float((__float_as_int(x) & (int)m) | (__float_as_int(CN_LOGZERO) & (int)~m));
}
__device__ __forceinline__ float pick(float x, float y, uint32_t m) {           // m ? y : x
  return __int_as_float((__float_as_int(x) & (int)~m) | (__float_as_int(y) & (int)m));
}

// One frame of the max-plus recursion for the SPL consecutive states of a lane (cassnat.py:314-333).
//   a      : alpha of the lane's states, updated in place
//   up1/2  : alpha of the two states below the lane's first one (previous frame)
//   eb, el : blank / label emissions of the frame for this lane
//   skm    : per label state (odd j) all-ones when the skip s-2 -> s is allowed, else zero
//   MODE   : 0 = utterance without masked frames: blank states take the two-way maximum (their forbidden skip holds
//            LOGZERO, which can only win once alpha itself has dropped below LOGZERO, i.e. after a masked frame);
//            1 = the full three-way comparison; 2 = frames of a sampled alignment, which take the runner-up
//            predecessor (cassnat.py:321-326)
// Returns the lane's 2-bit back-pointers (state j at bits 2j).  alpha = max over the predecessors: fmaxf equals the
// reference's sequence of strict '>' selects up to the sign of a zero, which no comparison can see.
template <int SPL, int MODE>
__device__ __forceinline__ uint32_t vit_frame(float (&a)[SPL], const uint32_t (&skm)[SPL / 2], float up1, float up2, float eb,
                                              const float (&el)[SPL / 2], bool first_lane) {
  float na[SPL];
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const float c0 = a[j];
    const float c1 = (j >= 1) ? a[j - 1] : up1;
    float best;
    constexpr bool SECOND = MODE == 2;
    if (!SECOND) {
      const float m01 = fmaxf(c0, c1);
      const uint32_t adv = c1 > c0 ? 1u : 0u;
      if (!(j & 1) && MODE == 0) {
        // blank state: stay or advance (the forbidden skip cannot win)
        best = m01;
        bits |= adv << (2 * j);
      } else {
        const float c2 = (j & 1) ? keep_or_logzero((j >= 2) ? a[j - 2] : up1, skm[j >> 1]) : CN_LOGZERO;
        best = fmaxf(m01, c2);
        bits |= (c2 > m01 ? 2u : adv) << (2 * j);
      }
    } else {
      float c2 = (j >= 2) ? a[j - 2] : (j == 1 ? up1 : up2);
      if (!(j & 1)) c2 = CN_LOGZERO; else c2 = keep_or_logzero(c2, skm[j >> 1]);
      uint32_t k = 0;
      best = c0;
      if (c1 > best) { best = c1; k = 1; }
      if (c2 > best) { best = c2; k = 2; }
      if (!(first_lane && j == 0)) {
        // 2nd entry of a stable descending sort of (c0,c1,c2)
        float vi, vj; uint32_t ii, jj;
        if (k == 0) { vi = c1; ii = 1; vj = c2; jj = 2; }
        else if (k == 1) { vi = c0; ii = 0; vj = c2; jj = 2; }
        else { vi = c0; ii = 0; vj = c1; jj = 1; }
        if (vj > vi) { best = vj; k = jj; } else { best = vi; k = ii; }
      }
      bits |= k << (2 * j);
    }
    na[j] = best + ((j & 1) ? el[j >> 1] : eb);
  }
#pragma unroll
  for (int j = 0; j < SPL; ++j) a[j] = na[j];
  return bits;
}

template <int SPL, int NW, int kFrames>      // kFrames: frames per emission chunk
struct VitShape {
  static constexpr int LANES = NW * 32;
  static constexpr int BITS = 2 * SPL;             // back-pointer bits per lane and frame
  static constexpr int FPW = 32 / BITS;            // frames packed per 32-bit word
  static constexpr int EW = LANES * SPL / 2 + 1;   // label -> column table: [0]=blank, [1..U]=labels, rest blank
  static constexpr int NL = 32 * SPL / 2;          // label slots of one DP warp
  static constexpr int SW = NL + 2;                // emission slice of a warp and frame: NL labels, the blank, a pad
                                                   // (even: a lane's label pair is one 8-byte load)
  static constexpr int CHUNK = NW * kFrames * SW * 4;                 // bytes of one staged chunk (all slices)
  static constexpr int KR = kFrames == 8 ? 7 : 6 * CHUNK <= 100 * 1024 ? 6 : 3;   // chunks in the ring: the lead over the DP has to
                                                                      // cover the DRAM latency, two CTAs share an SM
  static constexpr int RING = (KR * CHUNK > kBtBytes) ? KR * CHUNK : kBtBytes;
};

template <int SPL, int NW, int kFrames>
__global__ void __launch_bounds__(kThreads) viterbi_kernel(VitArgs p) {
  using Sh = VitShape<SPL, NW, kFrames>;
  constexpr int LANES = Sh::LANES, BITS = Sh::BITS, FPW = Sh::FPW, EW = Sh::EW, NL = Sh::NL, SW = Sh::SW, kRing = Sh::KR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* emis = reinterpret_cast<float*>(smem_raw);                       // [kRing][NW][kFrames][SW]
  int* cols = reinterpret_cast<int*>(smem_raw + Sh::RING);                // [EW] vocab column per emission slot
  unsigned short* states = reinterpret_cast<unsigned short*>(cols + EW);  // [T]
  uint8_t* maskb = reinterpret_cast<uint8_t*>(states + (p.T + 1) / 2 * 2);  // [T] frame mask of this utterance
  __shared__ float fin[LANES * SPL];
  __shared__ float seam[3][NW][kFrames][2];     // alpha of the last two states of a DP warp after each frame of a chunk
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
  for (int i = tid; i < kRing * NW * kFrames * SW; i += kThreads) emis[i] = 0.f;      // (slots beyond U stay zero)
  for (int t = tid; t < xb; t += kThreads) maskb[t] = mask_b[t];
  __syncthreads();

  // Emission staging.  The slice of DP warp sl for one frame is SW floats: index i < NL = label slot 1 + sl * NL + i
  // of the lattice, index NL = the blank.  Fill warp i serves slice i % NW (frames i / NW, + N_FILL / NW, ...); in
  // round r it stages chunk r + kRing - 1 - sl, the chunk DP warp sl reaches kRing - 1 rounds later.
  //   compact rows : asynchronous global -> shared copies (cp.async, coalesced, no registers held); a round waits
  //                  for the copies issued kRing - 2 rounds earlier, whose DRAM latency is paid by then; frames masked
  //                  out inside the valid length (not produced by the reference's prefix masks, but defined) are
  //                  patched to LOGZERO by the copying thread;
  //   posteriors   : see gather_pair below.
  constexpr int N_FILL = kThreads / 32 - NW;
  constexpr int F_STRIDE = N_FILL / NW;                          // fill warps per slice
  constexpr int IPP = (NL + 31) / 32;                            // label iterations per slice and frame
  static_assert(N_FILL % NW == 0, "fill warps are split evenly over the slices");
  const int U1 = U + 1;
  const bool compact = p.emis != nullptr;
  const int n_chunks = (xb + kFrames - 1) / kFrames;
  const int my_sl = w < 0 ? fill_rank % NW : 0, f_first = w < 0 ? fill_rank / NW : 0;
  auto slice_ptr = [&](int cc, int sl, int f) { return emis + (((cc % kRing) * NW + sl) * kFrames + f) * SW; };
  bool any_masked = false;
  for (int t = tid; t < xb; t += kThreads) any_masked |= maskb[t] == 0;
  any_masked = __syncthreads_or(any_masked);
  auto copy_async = [&](int r) {
    const int cc = r + kRing - 1 - my_sl;
    if (cc >= 0 && cc < n_chunks) {
      const int nf = min(kFrames, xb - cc * kFrames);
      const float* row = logp_b + (int64_t)(cc * kFrames + f_first) * row_pitch + 1 + my_sl * NL + lane;
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(slice_ptr(cc, my_sl, f_first)) + 4u * (uint32_t)lane;
      for (int f = f_first; f < nf; f += F_STRIDE) {
#pragma unroll
        for (int k = 0; k < IPP; ++k)
          if (k * 32 + lane < NL && 1 + my_sl * NL + k * 32 + lane < U1)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 128u * k), "l"(row + 32 * k) : "memory");
        if (lane == 0)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * NL), "l"(row - 1 - my_sl * NL) : "memory");
        row += (int64_t)F_STRIDE * row_pitch;
        dst += 4u * F_STRIDE * SW;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto patch = [&](int r) {
    const int cc = r + kRing - 1 - my_sl;
    if (!any_masked || cc < 0 || cc >= n_chunks) return;
    const int nf = min(kFrames, xb - cc * kFrames);
    for (int f = f_first; f < nf; f += F_STRIDE) {
      if (maskb[cc * kFrames + f]) continue;
      float* dst = slice_ptr(cc, my_sl, f);
      for (int idx = lane; idx < SW; idx += 32)
        if (idx == NL || 1 + my_sl * NL + idx < U1) dst[idx] = CN_LOGZERO;
    }
  };
  // posteriors, rings of at least NW + 3 chunks: all slices of a frame are gathered together, whole frames per fill
  // warp, two chunks (cc0, cc0 + 1 = r + 2, r + 3) every other round -- so that the U+1 columns of a 4V-byte
  // posterior row are touched within a few hundred nanoseconds of each other (fetching the four slices of a row in
  // four different rounds halves the DRAM efficiency: fewer hits per open page) and a lane keeps ~32 loads in flight
  static_assert(kFrames != 8 || kRing >= NW + 3, "posterior chunks live from two/three rounds before DP warp 0 to DP warp NW-1");
  constexpr int FPG = (2 * kFrames + N_FILL - 1) / N_FILL;       // frames per fill warp and chunk pair
  constexpr int KPF = (NW * NL + 31) / 32;                       // label iterations per whole frame
  auto gather_pair = [&](int cc0) {
    if (cc0 >= n_chunks) return;
    const int nf = min(2 * kFrames, xb - cc0 * kFrames);         // frames of the pair
    for (int k0 = 0; k0 < KPF; k0 += 8) {
      float v[FPG][8];
      float vb[FPG];
      int col[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) col[k] = cols[min(1 + (k0 + k) * 32 + lane, EW - 1)];
      const int colb = cols[0];
#pragma unroll
      for (int i = 0; i < FPG; ++i) {
        const int f = fill_rank + i * N_FILL;
        const int t = cc0 * kFrames + min(f, nf - 1);
        const bool valid = f < nf && maskb[t] != 0;
        const float* row = logp_b + (int64_t)t * row_pitch;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[i][k] = (valid && k0 + k < KPF && 1 + (k0 + k) * 32 + lane < U1) ? cn_ld_stream_f32(row + col[k]) : CN_LOGZERO;
        vb[i] = (valid && k0 == 0 && lane < NW) ? cn_ld_stream_f32(row + colb) : CN_LOGZERO;
      }
#pragma unroll
      for (int i = 0; i < FPG; ++i) {
        const int f = fill_rank + i * N_FILL;
        if (f >= nf) continue;
        const int cc = cc0 + f / kFrames, fc = f % kFrames;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int li = (k0 + k) * 32 + lane;                   // label slot - 1
          if (k0 + k < KPF && 1 + li < U1) slice_ptr(cc, li / NL, fc)[li % NL] = v[i][k];
        }
        if (k0 == 0 && lane < NW) slice_ptr(cc, lane, fc)[NL] = vb[i];      // the blank, once per slice
      }
    }
  };
  // what the first kRing - 1 rounds need (virtual rounds -(kRing-1) .. -1)
  if (w < 0) {
    if (compact) {
      for (int r = -(kRing - 1); r < 0; ++r) copy_async(r);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      for (int r = -(kRing - 1); r < 0; ++r) patch(r);
    } else {
      gather_pair(0);
    }
  }

  // ---- DP state of this lane.  States at or beyond the path length 2*ylen+1 are not masked here: a state only
  // ever feeds states above it, termination and back-trace stay below the path length, so whatever those lanes hold
  // never reaches the result (the reference's masked value is reconstructed at termination where python's negative
  // index reads one, ylens == 0).
  float a[SPL];
  uint32_t skm[SPL / 2];               // per label state of the lane: all-ones when the skip s-2 -> s is allowed
#pragma unroll
  for (int j = 0; j < SPL / 2; ++j) skm[j] = 0;
  if (w >= 0) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = g * SPL + j;
      a[j] = (s == 0) ? 0.f : CN_LOGZERO;
      // odd states carry labels; a skip s-2 -> s is allowed only between different labels
      if ((j & 1) && s >= 3 && s < S) {
        if (cols[(s + 1) >> 1] != cols[(s - 1) >> 1]) skm[j >> 1] = 0xffffffffu;
      }
    }
  }
  const uint32_t lane0_m = lane == 0 ? 0xffffffffu : 0u, has_in_m = (NW > 1 && w > 0) ? 0xffffffffu : 0u;
  const bool seam_rd = NW > 1 && w > 0 && lane == 0, seam_wr = NW > 1 && w >= 0 && w < NW - 1 && lane == 31;
  const uint32_t emis_s = (uint32_t)__cvta_generic_to_shared(emis);
  const uint32_t seam_s = (uint32_t)__cvta_generic_to_shared(&seam[0][0][0][0]);
  __syncthreads();

  uint32_t acc = 0;                    // back-pointers of this lane for the current group of FPW frames
  const int n_rounds = n_chunks > 0 ? n_chunks + NW - 1 : 0;
  for (int r = 0; r < n_rounds; ++r) {
    if (w >= 0) {
      const int c = r - w;                       // this warp's chunk of the round
      if (c >= 0 && c < n_chunks) {
        const int t0 = c * kFrames;
        const int nf = min(kFrames, xb - t0);
        // shared-memory addresses: this lane's labels / the blank of frame 0, the seam slots read and written
        const uint32_t e_lab = emis_s + 4u * (uint32_t)((((c % kRing) * NW + w) * kFrames) * SW + lane * (SPL / 2));
        const uint32_t e_blank = emis_s + 4u * (uint32_t)((((c % kRing) * NW + w) * kFrames) * SW + NL);
        const uint32_t s_out = seam_s + 8u * (uint32_t)(((c % 3) * NW + w) * kFrames);
        const uint32_t s_in = seam_s + 8u * (uint32_t)(((c % 3) * NW + (w > 0 ? w - 1 : 0)) * kFrames);
        uint32_t* bp_ptr = bp_b + (int64_t)(t0 / FPW) * LANES + g;
        uint32_t second_mask = 0;      // frames of this chunk that take the runner-up predecessor (sampled alignments)
        for (int i = 0; i < p.n_sample && i < 16; ++i) {
          const int d = samp[i] - t0;
          if (d >= 0 && d < nf) second_mask |= 1u << d;
        }
        // the seam values after the last frame of the previous chunk
        float in1 = CN_LOGZERO, in2 = CN_LOGZERO;
        if (seam_rd && c > 0) {
          const float2 v = lds_f32x2(seam_s + 8u * (uint32_t)((((c - 1) % 3) * NW + (w - 1)) * kFrames + kFrames - 1));
          in1 = v.x;
          in2 = v.y;
        }
        auto load_el = [&](int f, float (&el)[SPL / 2]) {
          if ((SPL / 2) % 2 == 0) {                       // a lane's labels start 8-byte aligned
#pragma unroll
            for (int j = 0; j < SPL / 2; j += 2) {
              const float2 v = lds_f32x2(e_lab + 4u * (uint32_t)(f * SW + j));
              el[j] = v.x;
              el[j + 1 < SPL / 2 ? j + 1 : j] = v.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < SPL / 2; ++j) el[j] = lds_f32(e_lab + 4u * (uint32_t)(f * SW + j));
          }
        };
        if (nf == kFrames && second_mask == 0 && !any_masked) {
          // the common case, fully unrolled: compile-time slots, no branches, every select on an invariant condition
          // done with masks.  Small lattices per lane: the chunk's emissions and seam values are read up front, so
          // that the recursion itself never waits behind the fill warps' traffic in the load/store unit.
          constexpr bool HOIST = SPL <= 4;
          float EB[HOIST ? kFrames : 1], EL[HOIST ? kFrames : 1][SPL / 2], IN1[HOIST ? kFrames : 1];
          if (HOIST) {
#pragma unroll
            for (int f = 0; f < kFrames; ++f) {
              load_el(f, EL[f]);
              EB[f] = lds_f32(e_blank + 4u * (uint32_t)(f * SW));
              IN1[f] = (f > 0 && NW > 1) ? lds_f32(s_in + 8u * (uint32_t)(f - 1)) : in1;     // (every lane: a broadcast)
            }
          }
#pragma unroll
          for (int f = 0; f < kFrames; ++f) {
            float el[SPL / 2];
            float eb, i1;
            if (HOIST) {
#pragma unroll
              for (int j = 0; j < SPL / 2; ++j) el[j] = EL[f][j];
              eb = EB[f];
              i1 = IN1[f];
            } else {
              load_el(f, el);
              eb = lds_f32(e_blank + 4u * (uint32_t)(f * SW));
              i1 = (f > 0 && NW > 1) ? lds_f32(s_in + 8u * (uint32_t)(f - 1)) : in1;
            }
            i1 = keep_or_logzero(i1, has_in_m);
            const float up1 = pick(__shfl_up_sync(0xffffffffu, a[SPL - 1], 1), i1, lane0_m);
            const uint32_t bits = vit_frame<SPL, 0>(a, skm, up1, 0.f, eb, el, g == 0);
            if (seam_wr) sts_f32x2(s_out + 8u * (uint32_t)f, a[SPL - 1], a[SPL - 2]);
            acc |= bits << ((f % FPW) * BITS);
            if (f % FPW == FPW - 1) {
              bp_ptr[(f / FPW) * LANES] = acc;
              acc = 0;
            }
          }
        } else {
          for (int f = 0; f < nf; ++f) {
            float el[SPL / 2];
            load_el(f, el);
            const float eb = lds_f32(e_blank + 4u * (uint32_t)(f * SW));
            if (f > 0 && seam_rd) {
              const float2 v = lds_f32x2(s_in + 8u * (uint32_t)(f - 1));
              in1 = v.x;
              in2 = v.y;
            }
            float up1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
            float up2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
            if (lane == 0) { up1 = in1; up2 = in2; }
            const uint32_t bits = ((second_mask >> f) & 1u) ? vit_frame<SPL, 2>(a, skm, up1, up2, eb, el, g == 0)
                                                            : vit_frame<SPL, 1>(a, skm, up1, up2, eb, el, g == 0);
            if (seam_wr) sts_f32x2(s_out + 8u * (uint32_t)f, a[SPL - 1], a[SPL - 2]);
            acc |= bits << ((f % FPW) * BITS);
            if (f % FPW == FPW - 1 || t0 + f == xb - 1) {
              bp_ptr[(f / FPW) * LANES] = acc;
              acc = 0;
            }
          }
        }
      }
    } else if (compact) {
      copy_async(r);
      asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 2) : "memory");      // the copies issued kRing - 2 rounds ago
      patch(r - (kRing - 2));                                                    // have landed
    } else if (!(r & 1)) {
      gather_pair(r + 2);
    }
    __syncthreads();
  }
  if (compact) asm volatile("cp.async.wait_group 0;" ::: "memory");

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
  constexpr int kBtG = Sh::RING / (LANES * 4);      // the whole emission ring is free by now
  uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw);
  for (int ghi = (xb - 1) / FPW; ghi >= 0 && xb > 1; ghi -= kBtG) {
    const int glo = max(ghi - kBtG + 1, 0);       // frame groups glo..ghi are consumed in this round
    const int n = (ghi - glo + 1) * LANES;
    {
      const uint4* src = reinterpret_cast<const uint4*>(bp_b + (int64_t)glo * LANES);
      uint4* dst = reinterpret_cast<uint4*>(stage);
#pragma unroll 4
      for (int i = tid; i < n / 4; i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    if (tid == 0) {
      // one thread walks the path down, branch-free: the word of (frame group, lane) covers FPW frames and is re-read
      // (predicated) only when the path moves to the lane below
      for (int gi = ghi; gi >= glo; --gi) {
        const uint32_t* base = stage + (gi - glo) * LANES;
        const int thi = min(xb - 1, gi * FPW + FPW - 1), tlo = max(gi * FPW, 1);
        int ln = state / SPL, j = state - ln * SPL;
        uint32_t wd = base[ln];
        int sh = (thi - gi * FPW) * BITS + 2 * j;          // bit position of (frame slot, state of the lane)
        for (int t = thi; t >= tlo; --t) {
          const int k = (int)((wd >> sh) & 3u);
          state -= k;
          j -= k;
          sh -= BITS + 2 * k;
          const bool down = j < 0;                           // the path moved to the lane below
          j += down ? SPL : 0;
          sh += down ? BITS : 0;
          ln = down ? max(ln - 1, 0) : ln;
          if (down) wd = base[ln];
          state = max(state, 0);                             // (lattice corner no feasible path visits)
          states[t - 1] = (unsigned short)state;
        }
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

template <int SPL, int NW, int KF>
int launch(const VitArgs& p, int64_t B, cudaStream_t stream) {
  using Sh = VitShape<SPL, NW, KF>;
  size_t smem = (size_t)Sh::RING + (size_t)Sh::EW * sizeof(int) + (size_t)((p.T + 1) / 2 * 2) * sizeof(unsigned short) + (size_t)p.T + 16;
  if (smem > 200 * 1024) return CN_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(viterbi_kernel<SPL, NW, KF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CN_ERR_LAUNCH;
  viterbi_kernel<SPL, NW, KF><<<(unsigned)B, kThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

// lattice shape for U labels: states per lane and DP warps
void shape_for(int64_t U, int& spl, int& nw) {
  const int64_t S = 2 * U + 1;
  if (S <= 64) { spl = 2; nw = 1; }
  else if (S <= 128) { spl = 4; nw = 1; }
  else { nw = 4; spl = 2 * (int)((S + 255) / 256); }
}

// Compact emissions: chunks of 16 frames.  Posterior rows: chunks of 8, which halves the shared-memory ring -- the
// scattered column gather keeps thousands of sectors in flight per SM and L1 is where they land; with two CTAs of
// ~107 KB each per SM the 30 KB left for L1 throttled the gather to half its rate (measured 1.28 vs 0.62 ms at cfg2).
template <int KF>
int dispatch_kf(const VitArgs& p, int64_t B, int64_t U, cudaStream_t stream) {
  int spl, nw;
  shape_for(U, spl, nw);
  if (nw == 1) return spl == 2 ? launch<2, 1, KF>(p, B, stream) : launch<4, 1, KF>(p, B, stream);
  switch (spl) {
    case 2: return launch<2, 4, KF>(p, B, stream);
    case 4: return launch<4, 4, KF>(p, B, stream);
    case 6: return launch<6, 4, KF>(p, B, stream);
    case 8: return launch<8, 4, KF>(p, B, stream);
    default: return CN_ERR_UNSUPPORTED;
  }
}
int dispatch(const VitArgs& p, int64_t B, int64_t U, cudaStream_t stream) {
  return p.emis ? dispatch_kf<16>(p, B, U, stream) : dispatch_kf<8>(p, B, U, stream);
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
