// Masked multi-head attention in split-bf16 ("bf16x3") arithmetic on the 5th-gen tensor cores: fp32-grade scores and
// context vectors at tensor-core speed for the precise mode (d_k = 64, Tk <= 2048, key masks).
//
// Q, K, V arrive as (hi, lo) bf16 pairs (hi = bf16(v), lo = bf16(v - hi): what the bf16x3 GEMMs emit).  Per
// (utterance, head, 128-query tile) item and 128-key tile j:
//   S_j = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T      12 tcgen05.mma (SS) into one of TWO S buffers in TMEM, so the
//                                                      issuer computes S_{j+1} while the softmax works on S_j;
//   softmax warps (thread = query row, the 128 scores of the tile in registers): scale, mask (finfo.min semantics of
//   modules/attention.py:19-21), exponentials against a lazily advanced maximum, P split into (hi, lo) bf16 pairs and
//   stored to TMEM (tcgen05.st);
//   O += P_hi V_hi + P_lo V_hi + P_hi V_lo             24 tcgen05.mma (A operand from TMEM) accumulating in TMEM;
//   the row sum normalises once at the end, the context rows are written in fp32.
// Relative error of S and O ~2^-17 (the dropped lo*lo terms), against 2^-9 for the bf16 kernel.
// Replaces (precise mode) attention() of src/models/modules/attention.py:13-24.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int QT = 128, KT = 128, DK = 64;
constexpr int KV_SLOTS = 2;          // K / V ring slots (hi + lo sub-tiles each)
constexpr int MAX_TK = 2048;
constexpr int NTHREADS = 192;        // warps 0-3 softmax, warp 4 MMA issuer, warp 5 TMA producer
constexpr float RESCALE_LOG2 = 8.f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();     // protocol bug -> launch failure, not a hang
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
        "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]),
        "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]),
        "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

struct AttnX3Args {
  const uint8_t* mask; int64_t mb;          // key mask u8 [B,Tk] (row pitch mb) or NULL
  int B, h, Tq, Tk;
  float* out; int64_t ldo, sbo;             // fp32 context [B,Tq,h*64], or (out == NULL)
  __nv_bfloat16* out16; int64_t lo_off;     // split-bf16 context: hi at out16, lo at out16 + lo_off (same pitches ldo / sbo)
  const int32_t* kv_len; const int32_t* q_len; const int32_t* order;     // as cn_attention_bf16
};

// smem: Q[2]{hi,lo} | K[2]{hi,lo} | V[2]{hi,lo} | barriers, key-mask words.  16 KB per sub-tile; the query tile is
// double-buffered so that the next item's Q is in flight while the current item computes.
constexpr uint32_t OFF_Q = 0, OFF_K = 4 * 16384, OFF_V = OFF_K + KV_SLOTS * 2 * 16384, OFF_BAR = OFF_V + KV_SLOTS * 2 * 16384,
                   SMEM_BYTES = OFF_BAR + 1024 + 1024 + 1024;     // (+1 KB alignment slack)
constexpr int B_QFULL = 0, B_QFREE = 2, B_KFULL = 4, B_KFREE = 6, B_VFULL = 8, B_VFREE = 10, B_SFULL = 12, B_SFREE = 14,
              B_PREADY = 16, B_PFREE = 17, N_BARS = 18;

__global__ void __launch_bounds__(NTHREADS, 1) attention_x3_kernel(const __grid_constant__ CUtensorMap map_qh,
                                                                   const __grid_constant__ CUtensorMap map_ql,
                                                                   const __grid_constant__ CUtensorMap map_kh,
                                                                   const __grid_constant__ CUtensorMap map_kl,
                                                                   const __grid_constant__ CUtensorMap map_vh,
                                                                   const __grid_constant__ CUtensorMap map_vl,
                                                                   AttnX3Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 256);
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 288);     // [MAX_TK / 32] key-mask words

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_max = (p.Tk + KT - 1) / KT;
  const int nq_max = (p.Tq + QT - 1) / QT;
  auto tiles_of = [&](int b, int& n_tiles, int& nq) {
    n_tiles = n_tiles_max;
    nq = nq_max;
    if (p.kv_len) { int t = (min(max(p.kv_len[b], 1), p.Tk) + KT - 1) / KT; n_tiles = min(max(t, 1), n_tiles); }
    if (p.q_len) { int t = (min(max(p.q_len[b], 0), p.Tq) + QT - 1) / QT; nq = min(t, nq_max); }
  };
  const int n_items = p.B * p.h * nq_max;              // work item = (utterance, head, query tile)
  auto decode_item = [&](int item, int& b, int& hh, int& qt) {
    const int u = item / (p.h * nq_max), rem = item % (p.h * nq_max);
    b = p.order ? p.order[u] : u;
    hh = rem / nq_max;
    qt = rem % nq_max;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_BARS; ++i) mbar_init(bar(i), (i == B_PREADY || i == B_SFREE || i == B_SFREE + 1) ? 4 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S[0] 0..127 | S[1] 128..255 | O 256..319 | P_hi 320..383 | P_lo 384..447 (bf16 pairs)
  auto tm_s = [&](int buf) { return tmem + buf * 128; };
  const uint32_t tm_o = tmem + 256, tm_ph = tmem + 320, tm_pl = tmem + 384;

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t q_cnt = 0, k_fill[KV_SLOTS] = {0, 0}, v_fill[KV_SLOTS] = {0, 0};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, hh, qt;
        decode_item(item, b, hh, qt);
        int n_tiles, nq;
        tiles_of(b, n_tiles, nq);
        if (qt >= nq) continue;
        const int qb = q_cnt & 1;
        mbar_wait(bar(B_QFREE + qb), ((q_cnt >> 1) & 1) ^ 1);
        mbar_expect_tx(bar(B_QFULL + qb), 2 * QT * DK * 2);
        tma_load_3d(&map_qh, bar(B_QFULL + qb), sbase + OFF_Q + qb * 32768, hh * DK, qt * QT, b);
        tma_load_3d(&map_ql, bar(B_QFULL + qb), sbase + OFF_Q + qb * 32768 + 16384, hh * DK, qt * QT, b);
        ++q_cnt;
        for (int j = 0; j < n_tiles; ++j) {
          const int sl = j % KV_SLOTS;
          mbar_wait(bar(B_KFREE + sl), (k_fill[sl] & 1) ^ 1);
          mbar_expect_tx(bar(B_KFULL + sl), 2 * KT * DK * 2);
          tma_load_3d(&map_kh, bar(B_KFULL + sl), sbase + OFF_K + sl * 32768, hh * DK, j * KT, b);
          tma_load_3d(&map_kl, bar(B_KFULL + sl), sbase + OFF_K + sl * 32768 + 16384, hh * DK, j * KT, b);
          ++k_fill[sl];
          mbar_wait(bar(B_VFREE + sl), (v_fill[sl] & 1) ^ 1);
          mbar_expect_tx(bar(B_VFULL + sl), 2 * KT * DK * 2);
          tma_load_3d(&map_vh, bar(B_VFULL + sl), sbase + OFF_V + sl * 32768, hh * DK, j * KT, b);
          tma_load_3d(&map_vl, bar(B_VFULL + sl), sbase + OFF_V + sl * 32768 + 16384, hh * DK, j * KT, b);
          ++v_fill[sl];
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(QT, KT, false);
      constexpr uint32_t idesc_o = idesc_bf16(QT, DK, true);
      uint32_t q_cnt = 0, k_seen[KV_SLOTS] = {0, 0}, v_seen[KV_SLOTS] = {0, 0}, s_uses[2] = {0, 0}, p_cnt = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, hh, qt;
        decode_item(item, b, hh, qt);
        int n_tiles, nq;
        tiles_of(b, n_tiles, nq);
        if (qt >= nq) continue;
        auto issue_s = [&](int j) {
          const int sl = j % KV_SLOTS, buf = j & 1, qb = q_cnt & 1;
          if (j == 0) mbar_wait(bar(B_QFULL + qb), (q_cnt >> 1) & 1);
          if (s_uses[buf] > 0) mbar_wait(bar(B_SFREE + buf), (s_uses[buf] - 1) & 1);   // the softmax has read this buffer
          ++s_uses[buf];
          mbar_wait(bar(B_KFULL + sl), k_seen[sl] & 1);
          ++k_seen[sl];
          fence_after();
          const uint32_t qh = sbase + OFF_Q + qb * 32768, ql = qh + 16384, kh = sbase + OFF_K + sl * 32768, kl = kh + 16384;
          const uint64_t dqh = desc_kmajor(qh), dql = desc_kmajor(ql), dkh = desc_kmajor(kh), dkl = desc_kmajor(kl);
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) umma(tm_s(buf), dqh + (uint64_t)(k * 2), dkh + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) umma(tm_s(buf), dql + (uint64_t)(k * 2), dkh + (uint64_t)(k * 2), idesc_s, 1u);
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) umma(tm_s(buf), dqh + (uint64_t)(k * 2), dkl + (uint64_t)(k * 2), idesc_s, 1u);
          tcgen05_commit(bar(B_SFULL + buf));
          tcgen05_commit(bar(B_KFREE + sl));
          if (j == n_tiles - 1) { tcgen05_commit(bar(B_QFREE + qb)); ++q_cnt; }
        };
        issue_s(0);
        for (int j = 0; j < n_tiles; ++j) {
          if (j + 1 < n_tiles) issue_s(j + 1);                   // runs ahead of the softmax by one key tile
          mbar_wait(bar(B_PREADY), p_cnt & 1);
          ++p_cnt;
          const int sl = j % KV_SLOTS;
          mbar_wait(bar(B_VFULL + sl), v_seen[sl] & 1);
          ++v_seen[sl];
          fence_after();
          const uint64_t dvh = desc_mnmajor(sbase + OFF_V + sl * 32768), dvl = desc_mnmajor(sbase + OFF_V + sl * 32768 + 16384);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k)
            umma_ts(tm_o, tm_ph + (uint32_t)(k * 8), dvh + (uint64_t)((k * 16 * 128) >> 4), idesc_o, (j | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k)
            umma_ts(tm_o, tm_pl + (uint32_t)(k * 8), dvh + (uint64_t)((k * 16 * 128) >> 4), idesc_o, 1u);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k)
            umma_ts(tm_o, tm_ph + (uint32_t)(k * 8), dvl + (uint64_t)((k * 16 * 128) >> 4), idesc_o, 1u);
          tcgen05_commit(bar(B_PFREE));
          tcgen05_commit(bar(B_VFREE + sl));
        }
      }
    }
  } else {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float scale = 0.125f * 1.4426950408889634f;      // 1/sqrt(64) * log2(e)
    uint32_t s_cnt[2] = {0, 0}, pf_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, hh, qt;
      decode_item(item, b, hh, qt);
      int n_tiles, nq;
      tiles_of(b, n_tiles, nq);
      if (qt >= nq) {
        // query rows the caller does not need: zero-filled
        const int r = qt * QT + row;
        if (r < p.Tq) {
          const int64_t off = (int64_t)b * p.sbo + (int64_t)r * p.ldo + hh * DK;
          if (p.out) {
            float4* dst = reinterpret_cast<float4*>(p.out + off);
#pragma unroll
            for (int d = 0; d < DK / 4; ++d) dst[d] = make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            uint4* dh = reinterpret_cast<uint4*>(p.out16 + off);
            uint4* dl = reinterpret_cast<uint4*>(p.out16 + p.lo_off + off);
#pragma unroll
            for (int d = 0; d < DK / 8; ++d) dh[d] = dl[d] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        continue;
      }
      if (p.mask) {
        asm volatile("bar.sync 1, 128;" ::: "memory");               // previous item's readers are done
        const uint8_t* mk = p.mask + (int64_t)b * p.mb;
        for (int k0 = 0; k0 < p.Tk; k0 += 128) {
          const int k = k0 + row;
          const uint32_t bits = __ballot_sync(0xffffffffu, k < p.Tk && mk[k] != 0);
          if (lane == 0) kbits[k >> 5] = bits;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int qi = qt * QT + row;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        mbar_wait(bar(B_SFULL + buf), s_cnt[buf] & 1);
        ++s_cnt[buf];
        fence_after();
        const int k0 = j * KT;
        float t[KT];
#pragma unroll
        for (int c = 0; c < KT; c += 32) tmem_ld32(tm_s(buf) + lane_addr + c, t + c);
        tmem_wait_ld();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_SFREE + buf));               // the issuer may overwrite this S buffer
        // ---- masks per 32-key chunk; raw chunks keep raw scores (exponent = fma(t, scale, -m)), the others hold
        // final log2-domain values
        uint32_t done = 0;
        float mx = -INFINITY;
#pragma unroll
        for (int ci = 0; ci < KT / 32; ++ci) {
          float* tc = t + ci * 32;
          const int n = p.Tk - (k0 + ci * 32);
          const uint32_t valid = n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
          const uint32_t keep = p.mask ? kbits[(k0 + ci * 32) >> 5] : 0xffffffffu;
          if ((keep & valid) == 0xffffffffu) {
            float cm[4] = {tc[0], tc[1], tc[2], tc[3]};
#pragma unroll
            for (int q = 4; q < 32; ++q) cm[q & 3] = fmaxf(cm[q & 3], tc[q]);
            mx = fmaxf(mx, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) * scale);
          } else {
            done |= 1u << ci;
            float mm = -INFINITY;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              float v = tc[q] * scale;
              v = ((keep >> q) & 1u) ? v : CN_F32_MIN;                  // masked: finfo.min, like the reference
              v = ((valid >> q) & 1u) ? v : -INFINITY;                  // beyond the key range: no weight at all
              tc[q] = v;
              mm = fmaxf(mm, v);
            }
            mx = fmaxf(mx, mm);
          }
        }
        // ---- lazy running maximum
        float corr = 1.f;
        bool rescale = false;
        if (mx > m_run + RESCALE_LOG2 || (m_run == -INFINITY && mx > -INFINITY)) {
          if (m_run != -INFINITY && j > 0) { rescale = true; corr = ex2(m_run - mx); }
          m_run = mx;
        }
        const float m_use = m_run == -INFINITY ? 0.f : m_run;
        // ---- exponentials, row sum, P split into (hi, lo) bf16 pairs (two keys per 32-bit TMEM column)
        float lsv[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t ph[KT / 2], pl[KT / 2];
#pragma unroll
        for (int ci = 0; ci < KT / 32; ++ci) {
          float* tc = t + ci * 32;
          if (!((done >> ci) & 1u)) {
#pragma unroll
            for (int q = 0; q < 32; ++q) tc[q] = fmaf(tc[q], scale, -m_use);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) tc[q] -= m_use;
          }
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float e0 = ex2(tc[q]), e1 = ex2(tc[q + 1]);
            lsv[(q >> 1) & 3] += e0 + e1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
            const float2 hf = __bfloat1622float2(h2);
            __nv_bfloat162 l2 = __floats2bfloat162_rn(e0 - hf.x, e1 - hf.y);
            ph[(ci * 32 + q) >> 1] = *reinterpret_cast<uint32_t*>(&h2);
            pl[(ci * 32 + q) >> 1] = *reinterpret_cast<uint32_t*>(&l2);
          }
        }
        // ---- P / O are free once PV(j-1) has retired
        if (j > 0) {
          mbar_wait(bar(B_PFREE), pf_cnt & 1);
          ++pf_cnt;
          fence_after();
          if (__any_sync(0xffffffffu, rescale)) {
            float o[DK];
            tmem_ld32(tm_o + lane_addr, o);
            tmem_ld32(tm_o + lane_addr + 32, o + 32);
            tmem_wait_ld();
#pragma unroll
            for (int d = 0; d < DK; ++d) o[d] *= corr;
            tmem_st32(tm_o + lane_addr, reinterpret_cast<uint32_t*>(o));
            tmem_st32(tm_o + lane_addr + 32, reinterpret_cast<uint32_t*>(o) + 32);
          }
        }
        tmem_st32(tm_ph + lane_addr, ph);
        tmem_st32(tm_ph + lane_addr + 32, ph + 32);
        tmem_st32(tm_pl + lane_addr, pl);
        tmem_st32(tm_pl + lane_addr + 32, pl + 32);
        tmem_wait_st();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PREADY));
        l_run = l_run * corr + ((lsv[0] + lsv[1]) + (lsv[2] + lsv[3]));
      }
      // ---- epilogue: O / l once the last PV has retired, fp32 rows
      mbar_wait(bar(B_PFREE), pf_cnt & 1);
      ++pf_cnt;
      fence_after();
      {
        float o[DK];
        tmem_ld32(tm_o + lane_addr, o);
        tmem_ld32(tm_o + lane_addr + 32, o + 32);
        tmem_wait_ld();
        fence_before();
        const float inv = 1.f / l_run;
        if (qi < p.Tq) {
          const int64_t off = (int64_t)b * p.sbo + (int64_t)qi * p.ldo + hh * DK;
          if (p.out) {
            float4* dst = reinterpret_cast<float4*>(p.out + off);
#pragma unroll
            for (int d = 0; d < DK; d += 4) dst[d >> 2] = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
          } else {                                       // (hi, lo) pairs: the operand of the three-pass output projection
            uint4* dh = reinterpret_cast<uint4*>(p.out16 + off);
            uint4* dl = reinterpret_cast<uint4*>(p.out16 + p.lo_off + off);
#pragma unroll
            for (int d = 0; d < DK; d += 8) {
              uint32_t wh[4], wl[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = o[d + 2 * e] * inv, c = o[d + 2 * e + 1] * inv;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
                const float2 hf = __bfloat1622float2(h2);
                __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, c - hf.y);
                wh[e] = *reinterpret_cast<uint32_t*>(&h2);
                wl[e] = *reinterpret_cast<uint32_t*>(&l2);
              }
              dh[d >> 3] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
              dl[d >> 3] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
            }
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 4) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// bf16 [B, T, width] (row pitch ld, batch pitch sb elements) -> 3-D map, box = 64 cols x 128 rows x 1
bool make_map3(CUtensorMap* map, const void* base, int64_t B, int64_t T, int64_t width, int64_t ld, int64_t sb) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int cn_attention_bf16x3(const void* Qhi, const void* Qlo, int64_t ldq, int64_t sbq, const void* Khi, const void* Klo,
                                   int64_t ldk, int64_t sbk, const void* Vhi, const void* Vlo, int64_t ldv, int64_t sbv,
                                   const uint8_t* mask, int64_t mb, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                                   float* out, void* out16, int64_t lo_off, int64_t ldo, int64_t sbo,
                                   const int32_t* kv_len, const int32_t* q_len, const int32_t* order, cn_stream_t stream_) {
  if (!Qhi || !Qlo || !Khi || !Klo || !Vhi || !Vlo || (!out == !out16) || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (dk != DK || Tk > MAX_TK || h > 65535 || B > 65535) return CN_ERR_UNSUPPORTED;
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (sbq & 7) || (sbk & 7) || (sbv & 7) || (ldo & 7) || (sbo & 7) || (lo_off & 7) ||
      ((reinterpret_cast<uintptr_t>(Qhi) | reinterpret_cast<uintptr_t>(Qlo) | reinterpret_cast<uintptr_t>(Khi) |
        reinterpret_cast<uintptr_t>(Klo) | reinterpret_cast<uintptr_t>(Vhi) | reinterpret_cast<uintptr_t>(Vlo) |
        reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out16)) & 15))
    return CN_ERR_UNSUPPORTED;
  if (B == 0 || Tq == 0) return CN_OK;
  CUtensorMap mqh, mql, mkh, mkl, mvh, mvl;
  const int64_t width = h * DK;
  if (!make_map3(&mqh, Qhi, B, Tq, width, ldq, sbq) || !make_map3(&mql, Qlo, B, Tq, width, ldq, sbq) ||
      !make_map3(&mkh, Khi, B, Tk, width, ldk, sbk) || !make_map3(&mkl, Klo, B, Tk, width, ldk, sbk) ||
      !make_map3(&mvh, Vhi, B, Tk, width, ldv, sbv) || !make_map3(&mvl, Vlo, B, Tk, width, ldv, sbv))
    return CN_ERR_DEVICE;
  AttnX3Args p;
  p.mask = mask; p.mb = mb; p.B = (int)B; p.h = (int)h; p.Tq = (int)Tq; p.Tk = (int)Tk;
  p.out = out; p.out16 = static_cast<__nv_bfloat16*>(out16); p.lo_off = lo_off; p.ldo = ldo; p.sbo = sbo;
  p.kv_len = mask ? kv_len : nullptr; p.q_len = q_len; p.order = order;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
      return CN_ERR_LAUNCH;
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = B * h * ((Tq + QT - 1) / QT);
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  attention_x3_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(mqh, mql, mkh, mkl, mvh, mvl, p);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
