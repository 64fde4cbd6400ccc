// K2 / K4 / K5: row statistics over posteriors, best-path collapse, trigger masks.
// Integer outputs are bit-exact restatements of src/models/cassnat.py:259-270, 355-389.
#include "common.cuh"

// ------------------------------------------------------------------ row statistics
struct Top2 {
  float v1, v2;
  int i1, i2;
};

__device__ __forceinline__ void top2_push(Top2& s, float v, int i) {
  // elements arrive in increasing index order within a lane, so strict '>' keeps the
  // lowest index among equal values
  if (v > s.v1) {
    s.v2 = s.v1; s.i2 = s.i1; s.v1 = v; s.i1 = i;
  } else if (v > s.v2) {
    s.v2 = v; s.i2 = i;
  }
}

__device__ __forceinline__ void top2_merge(Top2& a, float bv1, int bi1, float bv2, int bi2) {
  // merge two (sorted) pairs under the order: larger value first, then lower index
  if (cn_better(bv1, bi1, a.v1, a.i1)) {
    // b1 is the new best; runner-up is the better of a1 and b2
    if (cn_better(bv2, bi2, a.v1, a.i1)) { a.v2 = bv2; a.i2 = bi2; }
    else { a.v2 = a.v1; a.i2 = a.i1; }
    a.v1 = bv1; a.i1 = bi1;
  } else if (cn_better(bv1, bi1, a.v2, a.i2)) {
    a.v2 = bv1; a.i2 = bi1;
  }
}

// online log-sum-exp accumulator (per lane), merged across lanes at the end
struct Lse {
  float m, s;
};
__device__ __forceinline__ void lse_push(Lse& a, float v) {
  if (v > a.m) {
    a.s = a.s * __expf(a.m - v) + 1.f;
    a.m = v;
  } else {
    a.s += __expf(v - a.m);
  }
}

template <bool WANT_LSE, bool VEC4>
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ x, int64_t rows, int V,
                                                        int64_t row_stride, int32_t* __restrict__ top1,
                                                        int32_t* __restrict__ top2, float* __restrict__ peak,
                                                        float* __restrict__ lse_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps_total) {
    const float* row = x + r * row_stride;
    Top2 s{-INFINITY, -INFINITY, 0x7fffffff, 0x7fffffff};
    Lse l{-FLT_MAX, 0.f};
    if (VEC4) {
      const float4* row4 = reinterpret_cast<const float4*>(row);
      const int n4 = V >> 2;
      int c = lane;
      // 8 independent 16-byte loads in flight per lane
      for (; c + 7 * 32 < n4; c += 8 * 32) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = cn_ld_stream_f32x4(row4 + c + k * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int base = (c + k * 32) << 2;
          top2_push(s, v[k].x, base); top2_push(s, v[k].y, base + 1);
          top2_push(s, v[k].z, base + 2); top2_push(s, v[k].w, base + 3);
          if (WANT_LSE) { lse_push(l, v[k].x); lse_push(l, v[k].y); lse_push(l, v[k].z); lse_push(l, v[k].w); }
        }
      }
      for (; c < n4; c += 32) {
        float4 v = cn_ld_stream_f32x4(row4 + c);
        const int base = c << 2;
        top2_push(s, v.x, base); top2_push(s, v.y, base + 1);
        top2_push(s, v.z, base + 2); top2_push(s, v.w, base + 3);
        if (WANT_LSE) { lse_push(l, v.x); lse_push(l, v.y); lse_push(l, v.z); lse_push(l, v.w); }
      }
      for (int i = (n4 << 2) + lane; i < V; i += 32) {
        float v = cn_ld_stream_f32(row + i);
        top2_push(s, v, i);
        if (WANT_LSE) lse_push(l, v);
      }
    } else {
      for (int i = lane; i < V; i += 32) {
        float v = cn_ld_stream_f32(row + i);
        top2_push(s, v, i);
        if (WANT_LSE) lse_push(l, v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float bv1 = __shfl_xor_sync(0xffffffffu, s.v1, o), bv2 = __shfl_xor_sync(0xffffffffu, s.v2, o);
      int bi1 = __shfl_xor_sync(0xffffffffu, s.i1, o), bi2 = __shfl_xor_sync(0xffffffffu, s.i2, o);
      top2_merge(s, bv1, bi1, bv2, bi2);
    }
    if (WANT_LSE) {
      // exact final pass relative to the row max: sum_i exp(m_i - M) * s_i
      float part = l.s * expf(l.m - s.v1);   // lanes without elements: s == 0
      part = cn_warp_sum(part);
      if (lane == 0 && lse_out) lse_out[r] = s.v1 + logf(part);
    }
    if (lane == 0) {
      top1[r] = s.i1;
      if (top2) top2[r] = (V > 1) ? s.i2 : s.i1;
      if (peak) peak[r] = s.v1;
    }
  }
}

__global__ void __launch_bounds__(256) log_softmax_write_kernel(const float* __restrict__ x, int64_t rows, int V,
                                                                int64_t row_stride, const float* __restrict__ lse,
                                                                float* __restrict__ out, int64_t out_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps_total) {
    const float* row = x + r * row_stride;
    float* orow = out + r * out_stride;
    const float z = lse[r];
    for (int i = lane; i < V; i += 32) orow[i] = row[i] - z;
  }
}

extern "C" int cn_row_stats(const float* x, int64_t rows, int64_t V, int64_t row_stride, int32_t* top1,
                            int32_t* top2, float* peak, float* lse, float* logp_out, int64_t out_stride,
                            cn_stream_t stream_) {
  if (!x || !top1 || rows < 0 || V <= 0 || V > 0x7ffffff0 || row_stride < V) return CN_ERR_BAD_ARG;
  if (logp_out && !lse) return CN_ERR_BAD_ARG;
  if (rows == 0) return CN_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool vec4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (row_stride % 4 == 0) && V >= 4;
  const int warps_per_block = 8;
  int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: 16 CTAs (128 warps) resident per SM
  dim3 grid((unsigned)blocks), block(warps_per_block * 32);
  if (lse) {
    if (vec4) row_stats_kernel<true, true><<<grid, block, 0, stream>>>(x, rows, (int)V, row_stride, top1, top2, peak, lse);
    else row_stats_kernel<true, false><<<grid, block, 0, stream>>>(x, rows, (int)V, row_stride, top1, top2, peak, lse);
  } else {
    if (vec4) row_stats_kernel<false, true><<<grid, block, 0, stream>>>(x, rows, (int)V, row_stride, top1, top2, peak, nullptr);
    else row_stats_kernel<false, false><<<grid, block, 0, stream>>>(x, rows, (int)V, row_stride, top1, top2, peak, nullptr);
  }
  CN_CHECK_LAUNCH();
  if (logp_out) {
    log_softmax_write_kernel<<<grid, block, 0, stream>>>(x, rows, (int)V, row_stride, lse, logp_out, out_stride);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}

// ------------------------------------------------------------------ best-path collapse
// One CTA per utterance.  frame(t) = masked (sampled) arg-max label; a label survives only on
// the first frame of its run and is written one frame later (cassnat.py:380-385).
__global__ void __launch_bounds__(256) bestpath_collapse_kernel(
    const int32_t* __restrict__ top1, const int32_t* __restrict__ top2, const float* __restrict__ peak,
    const uint8_t* __restrict__ select, const uint8_t* __restrict__ src_mask, int T, int64_t blank,
    int sample_num, float threshold, int include_best, int64_t* __restrict__ aligned,
    int64_t* __restrict__ ylen, int32_t* __restrict__ ymax) {
  const int b = blockIdx.x;
  const int64_t base = (int64_t)b * T;
  const bool sampled = sample_num > 1 && select != nullptr && !(include_best && (b % sample_num) == 0);
  auto frame = [&](int t) -> int {
    if (t < 0 || !src_mask[base + t]) return 0;
    int lab = top1[base + t];
    if (sampled && select[base + t] && (expf(peak[base + t]) < threshold)) lab = top2[base + t];
    return lab;
  };
  int count = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int64_t out = 0;
    if (t > 0) {
      const int cur = frame(t - 1), prev = frame(t - 2);
      out = (cur != prev) ? cur : 0;
    }
    aligned[base + t] = out;
    count += (out != blank);
  }
  __shared__ int red[8];
  count = cn_warp_sum_i(count);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = count;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    ylen[b] = tot;
    atomicMax(ymax, tot);
  }
}

extern "C" int cn_bestpath_collapse(const int32_t* top1, const int32_t* top2, const float* peak,
                                    const uint8_t* select, const uint8_t* src_mask, int64_t B, int64_t T,
                                    int64_t blank, int64_t sample_num, float threshold, int include_best,
                                    int64_t* aligned, int64_t* ylen, int32_t* ymax, cn_stream_t stream_) {
  if (!top1 || !src_mask || !aligned || !ylen || !ymax || B < 0 || T <= 0 || T > 0x7ffffff0) return CN_ERR_BAD_ARG;
  if (sample_num > 1 && select && (!top2 || !peak)) return CN_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (cudaMemsetAsync(ymax, 0, sizeof(int32_t), stream) != cudaSuccess) return CN_ERR_LAUNCH;
  if (B == 0) return CN_OK;
  bestpath_collapse_kernel<<<(unsigned)B, 256, 0, stream>>>(top1, top2, peak, select, src_mask, (int)T, blank,
                                                             (int)sample_num, threshold, include_best, aligned,
                                                             ylen, ymax);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int64_t cn_bestpath_workspace_bytes(int64_t B, int64_t T) {
  // top1, top2 (i32) + peak (f32) per frame, each 256-byte aligned
  int64_t n = ((B * T * 4 + 255) / 256) * 256;
  return 3 * n;
}

extern "C" int cn_bestpath_from_logp(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                                     const uint8_t* src_mask, const uint8_t* select, int64_t blank,
                                     int64_t sample_num, float threshold, int include_best, int64_t* aligned,
                                     int64_t* ylen, int32_t* ymax, void* workspace, int64_t workspace_bytes,
                                     cn_stream_t stream) {
  if (!logp || !workspace) return CN_ERR_BAD_ARG;
  if (workspace_bytes < cn_bestpath_workspace_bytes(B, T)) return CN_ERR_WORKSPACE;
  if (sb != T * st) return CN_ERR_UNSUPPORTED;  // frames of all utterances must form one [B*T, V] matrix
  const int64_t n = ((B * T * 4 + 255) / 256) * 256;
  int32_t* top1 = (int32_t*)workspace;
  int32_t* top2 = (int32_t*)((char*)workspace + n);
  float* peak = (float*)((char*)workspace + 2 * n);
  const bool sampled = sample_num > 1 && select;
  int rc = cn_row_stats(logp, B * T, V, st, top1, sampled ? top2 : nullptr, sampled ? peak : nullptr, nullptr,
                        nullptr, 0, stream);
  if (rc != CN_OK) return rc;
  return cn_bestpath_collapse(top1, top2, peak, select, src_mask, B, T, blank, sample_num, threshold,
                              include_best, aligned, ylen, ymax, stream);
}

// ------------------------------------------------------------------ trigger masks
// base(u,t): align_to_mask's value (cassnat.py:358-361) = segment match AND src_mask, OR the
// eos fix-up cell (ylens[b], src_size[b]-1).
struct MaskCtx {
  const int* seg;          // shared: cumsum(aligned != blank) per frame
  const uint8_t* src;      // global: src_mask row
  int T, ylens, fix_t;
};
__device__ __forceinline__ bool mask_base(const MaskCtx& c, int u, int t) {
  if (t < 0 || t >= c.T) return false;
  return (c.seg[t] == u && c.src[t]) || (u == c.ylens && t == c.fix_t);
}
__device__ __forceinline__ bool mask_final(const MaskCtx& c, int u, int t, int left, int right, int and_src) {
  bool m = mask_base(c, u, t);
  if (right) m = m || mask_base(c, u, t - 1);
  if (left) m = m || mask_base(c, u, t + 1);   // (left of the right-dilated mask adds only base(u,t+1))
  if (and_src) m = m && c.src[t];
  return m;
}

__device__ void block_segment_scan(const int64_t* __restrict__ aligned_row, int T, int64_t blank, int* seg,
                                   int* warp_tot) {
  // inclusive prefix count of non-blank entries: a thread owns VT consecutive frames (one pass over its frames for the
  // local count, one block-wide scan of the 256 partial counts, one pass to write)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int VT = (T + (int)blockDim.x - 1) / (int)blockDim.x;
  const int t0 = threadIdx.x * VT, t1 = min(t0 + VT, T);
  int cnt = 0;
  for (int t = t0; t < t1; ++t) cnt += aligned_row[t] != blank ? 1 : 0;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_tot[w];
  int run = before + inc - cnt;                       // exclusive prefix of this thread's first frame
  for (int t = t0; t < t1; ++t) {
    run += aligned_row[t] != blank ? 1 : 0;
    seg[t] = run;
  }
  (void)nwarp;
  __syncthreads();
}

__global__ void __launch_bounds__(256) align_to_mask_kernel(
    const int64_t* __restrict__ aligned, const int64_t* __restrict__ ylens, const uint8_t* __restrict__ src_mask,
    const int64_t* __restrict__ src_size, int T, int n_rows, int64_t blank, int left, int right, int and_src,
    int rows_per_cta, uint8_t* __restrict__ mask_out, int64_t* __restrict__ ylen_out, int32_t* __restrict__ seg_out) {
  extern __shared__ int seg[];
  __shared__ int warp_tot[8];
  const int b = blockIdx.x;
  block_segment_scan(aligned + (int64_t)b * T, T, blank, seg, warp_tot);
  __syncthreads();
  MaskCtx c;
  c.seg = seg;
  c.src = src_mask + (int64_t)b * T;
  c.T = T;
  c.ylens = (int)ylens[b];
  int fx = (int)src_size[b] - 1;
  c.fix_t = fx < 0 ? fx + T : fx;  // python-style negative index wrap (src_size == 0)
  if (blockIdx.y == 0) {
    if (ylen_out && threadIdx.x == 0) ylen_out[b] = ylens[b] + 1;
    if (seg_out)
      for (int t = threadIdx.x; t < T; t += blockDim.x) seg_out[(int64_t)b * T + t] = seg[t];
  }
  if (!mask_out) return;
  const int u0 = blockIdx.y * rows_per_cta;
  const int u1 = min(u0 + rows_per_cta, n_rows);
  uint8_t* out = mask_out + ((int64_t)b * n_rows) * T;
  if ((T & 7) == 0 && (reinterpret_cast<uintptr_t>(mask_out) & 7) == 0) {
    // Word-wise synthesis: seg is non-decreasing, so row u's frames are one range [lo, hi] (two binary searches);
    // a lane builds the 32-frame word of the row from range masks, the source-mask word, the eos cell and the
    // dilation carries of the neighbouring words, then expands the bits to bytes (8-byte stores).  The
    // per-element predicate version below spent ~15 instructions per output byte.
    uint32_t* srcw = reinterpret_cast<uint32_t*>(seg + T);              // [ceil(T/32)] source-mask bits
    const int n_words = (T + 31) >> 5;
    for (int t0 = 0; t0 < n_words * 32; t0 += blockDim.x) {
      const int t = t0 + threadIdx.x;
      const uint32_t w = __ballot_sync(0xffffffffu, t < T && c.src[t] != 0);
      if ((threadIdx.x & 31) == 0 && (t >> 5) < n_words) srcw[t >> 5] = w;
    }
    __syncthreads();
    auto lower_bound = [&](int val) {                 // first t with seg[t] >= val
      int lo = 0, hi = T;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg[mid] < val) lo = mid + 1; else hi = mid; }
      return lo;
    };
    int* first = reinterpret_cast<int*>(srcw + n_words);               // [rows + 1]: first frame of every row of the CTA
    for (int i = threadIdx.x; i <= u1 - u0; i += blockDim.x) first[i] = lower_bound(u0 + i);      // (searches in parallel)
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int u = u0 + warp; u < u1; u += nwarp) {
      const int lo = first[u - u0], hi = first[u - u0 + 1] - 1;
      const bool eos_row = u == c.ylens && c.fix_t >= 0 && c.fix_t < T;
      auto base = [&](int w) -> uint32_t {            // align_to_mask's own value for frames [32w, 32w+32)
        if (w < 0 || w >= n_words) return 0u;
        const int a = max(lo - w * 32, 0), e = min(hi - w * 32, 31);
        uint32_t m = 0;
        if (a <= e) m = (e == 31 ? 0xffffffffu : ((1u << (e + 1)) - 1u)) & ~((1u << a) - 1u);
        m &= srcw[w];
        if (eos_row && (c.fix_t >> 5) == w) m |= 1u << (c.fix_t & 31);
        return m;
      };
      // a lane owns 8 frames (consecutive lanes write consecutive 8-byte pieces: 256 contiguous bytes per warp store);
      // the four lanes of a word recompute it
      for (int o8 = lane; o8 < (T >> 3); o8 += 32) {
        const int w = o8 >> 2, q = o8 & 3;
        const uint32_t b0 = base(w);
        uint32_t m = b0;
        if (right) m |= (b0 << 1) | (base(w - 1) >> 31);
        if (left) m |= (b0 >> 1) | (base(w + 1) << 31);
        if (and_src) m &= srcw[w];
        const uint32_t bits8 = m >> (8 * q);
        const uint32_t lo4 = ((bits8 & 0xFu) * 0x00204081u) & 0x01010101u;          // 4 bits -> 4 bytes
        const uint32_t hi4 = (((bits8 >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint2*>(out + (int64_t)u * T + o8 * 8) = make_uint2(lo4, hi4);
      }
    }
  } else if ((T & 3) == 0) {
    const int T4 = T >> 2;
    for (int idx = threadIdx.x; idx < (u1 - u0) * T4; idx += blockDim.x) {
      const int u = u0 + idx / T4, t = (idx % T4) << 2;
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) w |= (mask_final(c, u, t + k, left, right, and_src) ? 1u : 0u) << (8 * k);
      *reinterpret_cast<uint32_t*>(out + (int64_t)u * T + t) = w;
    }
  } else {
    for (int idx = threadIdx.x; idx < (u1 - u0) * T; idx += blockDim.x) {
      const int u = u0 + idx / T, t = idx % T;
      out[(int64_t)u * T + t] = mask_final(c, u, t, left, right, and_src) ? 1 : 0;
    }
  }
}

// Bit-packed per-row masks for the tcgen05 attention: bits[b][u][w] (16 words per row, Tk <= 512),
// bit (t & 31) of word (t >> 5) = keep(u, t).
__global__ void __launch_bounds__(256) trigger_mask_bits_kernel(const int32_t* __restrict__ seg_g,
                                                                const int64_t* __restrict__ ylens,
                                                                const uint8_t* __restrict__ src_mask,
                                                                const int64_t* __restrict__ src_size, int T, int n_rows,
                                                                int left, int right, uint32_t* __restrict__ bits) {
  // seg is non-decreasing along t, so row u's frames are one range [lo, hi]: two binary searches and
  // word-wise range masks instead of T predicate evaluations per row
  extern __shared__ int seg[];
  __shared__ uint32_t srcw[16];
  const int b = blockIdx.x, lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < T; t += blockDim.x) seg[t] = seg_g[(int64_t)b * T + t];
  for (int t0 = 0; t0 < 512; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const uint32_t w = __ballot_sync(0xffffffffu, t < T && src_mask[(int64_t)b * T + t] != 0);
    if (lane == 0) srcw[t >> 5] = w;
  }
  __syncthreads();
  const int yl = (int)ylens[b];
  int fx = (int)src_size[b] - 1;
  fx = fx < 0 ? fx + T : fx;
  auto lower_bound = [&](int val) {                 // first t with seg[t] >= val
    int lo = 0, hi = T;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg[mid] < val) lo = mid + 1; else hi = mid; }
    return lo;
  };
  for (int u = threadIdx.x; u < n_rows; u += blockDim.x) {
    const int lo = lower_bound(u), hi = lower_bound(u + 1) - 1;           // frames with seg == u (may be empty)
    uint32_t base[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const int a = max(lo - w * 32, 0), e = min(hi - w * 32, 31);
      uint32_t m = 0;
      if (a <= e) m = (e == 31 ? 0xffffffffu : ((1u << (e + 1)) - 1u)) & ~((1u << a) - 1u);
      m &= srcw[w];
      if (u == yl && fx >= 0 && fx < T && (fx >> 5) == w) m |= 1u << (fx & 31);      // eos row is never empty
      base[w] = m;
    }
    uint32_t* out = bits + ((int64_t)b * n_rows + u) * 16;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      uint32_t m = base[w];
      if (right) m |= (base[w] << 1) | (w > 0 ? base[w - 1] >> 31 : 0u);
      if (left) m |= (base[w] >> 1) | (w < 15 ? base[w + 1] << 31 : 0u);
      out[w] = m & srcw[w];
    }
  }
}

__global__ void __launch_bounds__(256) pack_mask_bits_kernel(const uint8_t* __restrict__ mask, int64_t mb, int64_t mi,
                                                             int Tq, int Tk, int64_t rows, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / Tq;
  const int i = (int)(row % Tq);
  const uint8_t* m = mask + b * mb + (int64_t)i * mi;
  for (int w = 0; w < 16; ++w) {
    const int t = w * 32 + lane;
    const uint32_t word = __ballot_sync(0xffffffffu, t < Tk && m[t] != 0);
    if (lane == 0) bits[row * 16 + w] = word;
  }
}

// Length hints of one decode step for the tensor-core attention, in ONE launch (replaces ~25 tiny elementwise /
// sort launches): rows the pick loop can read, last valid token / frame per utterance, and the heaviest-first
// processing orders of the self- and the source-attention.  Single CTA, B <= 1024.
__global__ void __launch_bounds__(1024) decode_hints_kernel(const int64_t* __restrict__ ylen, int ymax,
                                                            const uint8_t* __restrict__ src_mask, int B, int T, int use_q,
                                                            int32_t* __restrict__ need_rows, int32_t* __restrict__ kv_tgt,
                                                            int32_t* __restrict__ last_key, int32_t* __restrict__ order_tgt,
                                                            int32_t* __restrict__ order_src) {
  __shared__ int key_t[1024], key_s[1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = warp; b < B; b += 32) {                 // last unmasked frame + 1 (scan backwards, 32 frames per step)
    const uint8_t* m = src_mask + (int64_t)b * T;
    int last = 0;
    for (int t0 = ((T - 1) / 32) * 32; t0 >= 0; t0 -= 32) {
      const int t = t0 + lane;
      const uint32_t bits = __ballot_sync(0xffffffffu, t < T && m[t] != 0);
      if (bits) { last = t0 + 32 - __clz(bits); break; }
    }
    if (lane == 0) {
      const int yl = (int)ylen[b];
      const int nr = min(yl + 1, ymax), kt = min(yl, ymax);
      need_rows[b] = nr; kv_tgt[b] = kt; last_key[b] = last;
      const int qt = use_q ? (nr + 127) / 128 : 1;
      key_t[b] = ((kt + 127) / 128) * qt;
      key_s[b] = ((last + 127) / 128) * qt;
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {    // stable descending rank by counting
    int rt = 0, rs = 0;
    const int kt = key_t[b], ks = key_s[b];
    for (int j = 0; j < B; ++j) {
      rt += (key_t[j] > kt) || (key_t[j] == kt && j < b);
      rs += (key_s[j] > ks) || (key_s[j] == ks && j < b);
    }
    order_tgt[rt] = b;
    order_src[rs] = b;
  }
}

extern "C" int cn_decode_hints(const int64_t* ylen, int64_t ymax, const uint8_t* src_mask, int64_t B, int64_t T, int use_q,
                               int32_t* need_rows, int32_t* kv_tgt, int32_t* last_key, int32_t* order_tgt,
                               int32_t* order_src, cn_stream_t stream) {
  if (!ylen || !src_mask || !need_rows || !kv_tgt || !last_key || !order_tgt || !order_src || B < 0 || T <= 0 || ymax < 0)
    return CN_ERR_BAD_ARG;
  if (B > 1024 || ymax > 0x7ffffff) return CN_ERR_UNSUPPORTED;
  if (B == 0) return CN_OK;
  decode_hints_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ylen, (int)ymax, src_mask, (int)B, (int)T, use_q, need_rows, kv_tgt,
                                                          last_key, order_tgt, order_src);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_trigger_mask_bits(const int32_t* seg, const int64_t* ylens, const uint8_t* src_mask,
                                    const int64_t* src_size, int64_t B, int64_t T, int64_t n_rows, int left, int right,
                                    uint32_t* bits, cn_stream_t stream) {
  if (!seg || !ylens || !src_mask || !src_size || !bits || B < 0 || T <= 0 || n_rows < 0) return CN_ERR_BAD_ARG;
  if (T > 512 || B > 0x7fffffff) return CN_ERR_UNSUPPORTED;
  if (B == 0 || n_rows == 0) return CN_OK;
  trigger_mask_bits_kernel<<<(unsigned)B, 256, (size_t)T * sizeof(int), (cudaStream_t)stream>>>(
      seg, ylens, src_mask, src_size, (int)T, (int)n_rows, left, right, bits);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_pack_mask_bits(const uint8_t* mask, int64_t mb, int64_t mi, int64_t B, int64_t Tq, int64_t Tk,
                                 uint32_t* bits, cn_stream_t stream) {
  if (!mask || !bits || B < 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (Tk > 512) return CN_ERR_UNSUPPORTED;
  const int64_t rows = B * Tq;
  if (rows == 0) return CN_OK;
  pack_mask_bits_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(mask, mb, mi, (int)Tq, (int)Tk, rows, bits);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_align_to_mask(const int64_t* aligned, const int64_t* ylens, const uint8_t* src_mask,
                                const int64_t* src_size, int64_t B, int64_t T, int64_t ymax, int64_t blank,
                                int left, int right, int and_src_mask, uint8_t* mask_out, int64_t* ylen_out,
                                int32_t* seg_out, cn_stream_t stream_) {
  if (!aligned || !ylens || !src_mask || !src_size || B < 0 || T <= 0 || ymax < 0) return CN_ERR_BAD_ARG;
  if (T > 12000) return CN_ERR_UNSUPPORTED;  // segment ids are staged in shared memory (4 B / frame)
  if (B == 0) return CN_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n_rows = (int)ymax + 1;
  // aim for >= 4 waves of CTAs over 148 SMs while keeping >= 8 rows (8*T bytes of stores) per CTA
  int chunks = (int)((148 * 8 + B - 1) / B);
  int rows_per_cta = (n_rows + chunks - 1) / chunks;
  if (rows_per_cta < 8) rows_per_cta = 8;
  chunks = (n_rows + rows_per_cta - 1) / rows_per_cta;
  if (!mask_out) chunks = 1;
  dim3 grid((unsigned)B, (unsigned)chunks);
  align_to_mask_kernel<<<grid, 256, (size_t)T * sizeof(int) + (size_t)((T + 31) / 32) * sizeof(uint32_t) + (size_t)(rows_per_cta + 1) * sizeof(int), stream>>>(
      aligned, ylens, src_mask, src_size, (int)T, n_rows, blank, left != 0, right != 0, and_src_mask != 0,
      rows_per_cta, mask_out, ylen_out, seg_out);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

__global__ void __launch_bounds__(256) expand_mask_kernel(const uint8_t* __restrict__ m, int64_t rows, int T,
                                                          int left, int right, uint8_t* __restrict__ out) {
  const int64_t n = rows * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    bool v = m[i] != 0;
    if (right && t > 0) v = v || (m[i - 1] != 0);
    if (left && t + 1 < T) v = v || (m[i + 1] != 0);
    out[i] = v ? 1 : 0;
  }
}

extern "C" int cn_expand_trigger_mask(const uint8_t* mask, int64_t B, int64_t U, int64_t T, int left, int right,
                                      uint8_t* out, cn_stream_t stream_) {
  if (!mask || !out || mask == out || B < 0 || U < 0 || T <= 0) return CN_ERR_BAD_ARG;
  const int64_t n = B * U * T;
  if (n == 0) return CN_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  expand_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(mask, B * U, (int)T, left > 0, right > 0, out);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

__global__ void tgt_length_mask_kernel(const int64_t* __restrict__ ylen, int64_t B, int ymax,
                                       int64_t* __restrict__ o64, uint8_t* __restrict__ o8) {
  const int64_t n = B * ymax;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ymax, u = i % ymax;
    const int v = u < ylen[b];
    if (o64) o64[i] = v;
    if (o8) o8[i] = (uint8_t)v;
  }
}

extern "C" int cn_tgt_length_mask(const int64_t* ylen, int64_t B, int64_t ymax, int64_t* out_i64, uint8_t* out_u8,
                                  cn_stream_t stream_) {
  if (!ylen || B < 0 || ymax < 0 || ymax > 0x7ffffff0) return CN_ERR_BAD_ARG;
  if (B * ymax == 0) return CN_OK;
  int64_t blocks = (B * ymax + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  tgt_length_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(ylen, B, (int)ymax, out_i64, out_u8);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// ------------------------------------------------------------------ K11 sampled-alignment rescoring
//   replaces: gather / masked_fill / sum / count / max        src/models/cassnat.py:517-520
// one CTA per utterance, one warp per sample: mean over the valid tokens of the ranker's log-prob at
// the predicted id (entries that are exactly 0 do not count, as in the reference), then the first
// arg-max over the S samples (NaN ranks highest, like torch.max).
__global__ void __launch_bounds__(1024) rescore_select_kernel(const float* __restrict__ lm_out, int64_t s_row, int64_t s_tok,
                                                              const int32_t* __restrict__ ids,
                                                              const uint8_t* __restrict__ tgt_mask, int S, int U,
                                                              float* __restrict__ mean_out, int64_t* __restrict__ best) {
  __shared__ float means[32];
  const int b = blockIdx.x, s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (s < S) {
    const int64_t r = (int64_t)b * S + s;
    float sum = 0.f;
    int cnt = 0;
    for (int u = lane; u < U; u += 32) {
      float v = 0.f;
      if (tgt_mask[r * U + u]) v = lm_out[r * s_row + (int64_t)u * s_tok + ids[r * U + u]];
      sum += v;
      cnt += (v != 0.f);
    }
    sum = cn_warp_sum(sum);
    cnt = cn_warp_sum_i(cnt);
    if (lane == 0) {
      const float m = sum / (float)cnt;
      means[s] = m;
      mean_out[r] = m;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int arg = 0;
    float top = means[0];
    for (int k = 1; k < S; ++k) {
      const float v = means[k];
      if (top != top) break;                         // a NaN already won
      if (v != v || v > top) { top = v; arg = k; }
    }
    best[b] = arg;
  }
}

extern "C" int cn_rescore_select(const float* lm_out, int64_t row_stride, int64_t tok_stride, const int32_t* ids,
                                 const uint8_t* tgt_mask, int64_t B, int64_t S, int64_t U, float* mean_out,
                                 int64_t* best, cn_stream_t stream_) {
  if (!lm_out || !ids || !tgt_mask || !mean_out || !best || B < 0 || S <= 0 || U <= 0) return CN_ERR_BAD_ARG;
  if (S > 32) return CN_ERR_UNSUPPORTED;
  if (B == 0) return CN_OK;
  rescore_select_kernel<<<(unsigned)B, (unsigned)(32 * S), 0, (cudaStream_t)stream_>>>(lm_out, row_stride, tok_stride, ids,
                                                                                      tgt_mask, (int)S, (int)U, mean_out, best);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
