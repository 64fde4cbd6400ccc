// Masked multi-head attention on the 5th-gen tensor cores (fast mode), d_k = 64, Tk <= 512.
//
// One CTA per (128-query tile, head, utterance).  Warps 0-3 (one thread per query row = one TMEM
// lane) run the softmax; warp 4 issues TMA and tcgen05.mma:
//   * Q tile and ALL K/V tiles of this (utterance, head) are fetched up front by TMA (3-D tensor
//     maps over [B, T, row], 128-byte swizzle), one mbarrier per tile, so S_0 starts as soon as
//     K_0 lands;
//   * S_j = Q K_j^T  (128 x 128 x 64, K-major operands) into a double-buffered TMEM accumulator;
//   * softmax warps read S_j with tcgen05.ld, apply 1/sqrt(d_k) and the mask with the reference's
//     finfo.min semantics (fully masked rows stay uniform over all Tk keys, modules/attention.py:19-21),
//     keep a running max / sum, and write P_j as bf16 into shared memory in the K-major swizzled
//     layout the next MMA reads;
//   * O_j = P_j V_j (128 x 64 x 128; V tile used in place as an MN-major operand) into a
//     double-buffered TMEM slot, folded into per-row fp32 registers with the online-softmax
//     correction.
// Replaces (fast mode) attention() of src/models/modules/attention.py:13-24.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int QT = 128;        // queries per CTA
constexpr int KT = 128;        // keys per tile
constexpr int DK = 64;
constexpr int MAX_TILES = 4;   // Tk <= 512
constexpr int THREADS = 160;

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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major operand, 128-byte swizzle, 8-row groups 1024 B apart (same as the GEMM)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major operand (rows = K index, 64 contiguous MN elements = one 128-byte swizzle atom):
// stride byte offset = 1024 B between 8-row K groups, leading byte offset (next MN atom) unused.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// D = f32, A = B = bf16; bit 16 set -> B is MN-major
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

struct AttnTcArgs {
  const uint8_t* mask; int64_t mb, mi;      // u8, element (b,i,j) at mask[b*mb + i*mi + j]; mi == 0 -> key mask
  int h, Tq, Tk;
  __nv_bfloat16* out; int64_t ldo, sbo;
};

// smem map (bytes): Q 16K | K 4x16K | V 4x16K | P 32K | barriers
constexpr uint32_t OFF_Q = 0, OFF_K = 16384, OFF_V = OFF_K + MAX_TILES * 16384, OFF_P = OFF_V + MAX_TILES * 16384,
                   OFF_BAR = OFF_P + 32768, SMEM_BYTES = OFF_BAR + 256 + 1024;

__global__ void __launch_bounds__(THREADS, 1) attention_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                  const __grid_constant__ CUtensorMap map_k,
                                                                  const __grid_constant__ CUtensorMap map_v,
                                                                  AttnTcArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  // barriers: 0 q_full | 1..4 k_full[j] | 5..8 v_full[j] | 9,10 s_full[2] | 11 p_ready | 12 p_free | 13,14 o_full[2]
  const uint32_t bars = sbase + OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 128);
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 160);     // 16 words: key mask of this utterance

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT, hh = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Tk + KT - 1) / KT;

  if (threadIdx.x < 16) kbits[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(9), 1); mbar_init(bar(10), 1);
    mbar_init(bar(11), 128); mbar_init(bar(12), 1);
    mbar_init(bar(13), 1); mbar_init(bar(14), 1);
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
  const uint32_t TM_S = tmem, TM_O = tmem + 256;          // S: 2 x 128 columns, O: 2 x 64 columns

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA: Q, then K_j / V_j for every key tile
      mbar_expect_tx(bar(0), QT * DK * 2);
      tma_load_3d(&map_q, bar(0), sbase + OFF_Q, hh * DK, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_expect_tx(bar(1 + j), KT * DK * 2);
        tma_load_3d(&map_k, bar(1 + j), sbase + OFF_K + j * 16384, hh * DK, j * KT, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        mbar_expect_tx(bar(5 + j), KT * DK * 2);
        tma_load_3d(&map_v, bar(5 + j), sbase + OFF_V + j * 16384, hh * DK, j * KT, b);
      }
      // ---- MMA issue
      constexpr uint32_t idesc_s = idesc_bf16(QT, KT, false);
      constexpr uint32_t idesc_o = idesc_bf16(QT, DK, true);
      mbar_wait(bar(0), 0);
      auto issue_s = [&](int j) {
        mbar_wait(bar(1 + j), 0);
        fence_after();
        const uint64_t dq = desc_kmajor(sbase + OFF_Q), dkk = desc_kmajor(sbase + OFF_K + j * 16384);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) umma(TM_S + (j & 1) * 128, dq + (uint64_t)(k * 2), dkk + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
        tcgen05_commit(bar(9 + (j & 1)));
      };
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);               // S_{j+1} overlaps the softmax of tile j
        mbar_wait(bar(11), j & 1);                          // P_j written by all 128 softmax threads
        mbar_wait(bar(5 + j), 0);                           // V_j landed
        fence_after();
        const uint64_t dv = desc_mnmajor(sbase + OFF_V + j * 16384);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          const uint64_t dp = desc_kmajor(sbase + OFF_P + (k >> 2) * 16384) + (uint64_t)((k & 3) * 2);
          umma(TM_O + (j & 1) * 64, dp, dv + (uint64_t)((k * 16 * 128) >> 4), idesc_o, k ? 1u : 0u);
        }
        tcgen05_commit(bar(12));                            // P buffer free again
        tcgen05_commit(bar(13 + (j & 1)));                  // O_j ready
      }
    }
  } else {
    // ---- softmax warps: thread = query row = TMEM lane
    const int row = warp * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float scale = 0.125f * 1.4426950408889634f;      // 1/sqrt(64) * log2(e)
    float o[DK];
#pragma unroll
    for (int d = 0; d < DK; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, corr_prev = 1.f;
    if (p.mask && p.mi == 0) {
      // key-mask bits of this utterance, built by the 128 softmax threads while the TMA loads fly
      const uint8_t* mk = p.mask + (int64_t)b * p.mb;
      for (int k0 = 0; k0 < p.Tk; k0 += 128) {
        const int k = k0 + (int)threadIdx.x;
        const uint32_t bits = __ballot_sync(0xffffffffu, k < p.Tk && mk[k] != 0);
        if (lane == 0) kbits[k >> 5] = bits;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const uint8_t* mrow = p.mask ? p.mask + (int64_t)b * p.mb + (int64_t)(qi < p.Tq ? qi : 0) * p.mi : nullptr;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(bar(9 + (j & 1)), (j >> 1) & 1);
      fence_after();
      // Two sweeps over S_j straight out of TMEM, 32 columns at a time (keeps the loop body small
      // enough for the instruction cache and the row out of the register file): max, then exp/P.
      const int k0 = j * KT;
      const uint32_t s_addr = TM_S + lane_addr + (j & 1) * 128;
      auto chunk_bits = [&](int c, uint32_t& keep, uint32_t& valid) {
        const int n = p.Tk - (k0 + c);
        valid = n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
        if (!p.mask) keep = 0xffffffffu;
        else if (p.mi == 0) keep = kbits[(k0 + c) >> 5];              // key mask: shared by all rows
        else {
          keep = 0;
          for (int q = 0; q < 32; ++q)
            if (k0 + c + q < p.Tk && mrow[k0 + c + q]) keep |= 1u << q;
        }
      };
      // Software pipeline over the four 32-column chunks: the tcgen05.ld of chunk i+1 is in flight
      // while chunk i is processed (tcgen05.wait::ld waits for ALL outstanding loads of the thread).
      // four independent max / sum chains: one softmax warp per scheduler has to supply its own ILP
      float mxv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, mx_masked = -INFINITY;
      auto sweep_max = [&](const float* t, int c) {
        uint32_t keep, valid;
        chunk_bits(c, keep, valid);
        if ((keep & valid) == 0xffffffffu) {                            // common case: nothing masked in this chunk
#pragma unroll
          for (int q = 0; q < 32; ++q) mxv[q & 3] = fmaxf(mxv[q & 3], t[q]);
        } else {
          float mm = -INFINITY;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float v = t[q] * scale;
            v = ((keep >> q) & 1u) ? v : CN_F32_MIN;                  // masked: finfo.min, like the reference
            v = ((valid >> q) & 1u) ? v : -INFINITY;                  // beyond the key range: no weight at all
            mm = fmaxf(mm, v);
          }
          mx_masked = fmaxf(mx_masked, mm);
        }
      };
      float ta[32], tb[32];
      tmem_ld32(s_addr, ta); tmem_wait_ld();
      tmem_ld32(s_addr + 32, tb); sweep_max(ta, 0); tmem_wait_ld();
      tmem_ld32(s_addr + 64, ta); sweep_max(tb, 32); tmem_wait_ld();
      tmem_ld32(s_addr + 96, tb); sweep_max(ta, 64); tmem_wait_ld();
      tmem_ld32(s_addr, ta); sweep_max(tb, 96);                          // chunk 0 again, for the exp sweep
      const float mx = fmaxf(fmaxf(fmaxf(mxv[0], mxv[1]), fmaxf(mxv[2], mxv[3])) * scale, mx_masked);   // scale > 0
      const float m_new = fmaxf(m_run, mx);
      const float corr = ex2(m_run - m_new);
      float lsv[4] = {0.f, 0.f, 0.f, 0.f};
      if (j > 0) mbar_wait(bar(12), (j - 1) & 1);                    // PV_{j-1} finished reading the P buffer
      uint8_t* prow = smem + OFF_P + row * 128;
      auto sweep_exp = [&](float* t, int c) {
        uint32_t keep, valid;
        chunk_bits(c, keep, valid);
        if ((keep & valid) == 0xffffffffu) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            t[q] = ex2(fmaf(t[q], scale, -m_new));
            lsv[q & 3] += t[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float v = t[q] * scale;
            v = ((keep >> q) & 1u) ? v : CN_F32_MIN;
            v = ((valid >> q) & 1u) ? v : -INFINITY;
            t[q] = ex2(v - m_new);
            lsv[q & 3] += t[q];
          }
        }
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          __nv_bfloat162 b0 = __floats2bfloat162_rn(t[g], t[g + 1]), b1 = __floats2bfloat162_rn(t[g + 2], t[g + 3]),
                         b2 = __floats2bfloat162_rn(t[g + 4], t[g + 5]), b3 = __floats2bfloat162_rn(t[g + 6], t[g + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&b0); pk.y = *reinterpret_cast<uint32_t*>(&b1);
          pk.z = *reinterpret_cast<uint32_t*>(&b2); pk.w = *reinterpret_cast<uint32_t*>(&b3);
          const int col = c + g;
          const int atom = col >> 6, chunk = (col & 63) >> 3;         // 64-key swizzle atoms, 16-byte chunks
          *reinterpret_cast<uint4*>(prow + atom * 16384 + ((chunk ^ (row & 7)) << 4)) = pk;
        }
      };
      tmem_wait_ld();
      tmem_ld32(s_addr + 32, tb); sweep_exp(ta, 0); tmem_wait_ld();
      tmem_ld32(s_addr + 64, ta); sweep_exp(tb, 32); tmem_wait_ld();
      tmem_ld32(s_addr + 96, tb); sweep_exp(ta, 64); tmem_wait_ld();
      sweep_exp(tb, 96);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the MMA
      fence_before();                                                  // S_j (TMEM) fully read before it is handed back
      mbar_arrive(bar(11));
      // fold O_{j-1} (finished long ago) while PV_j runs
      if (j > 0) {
        mbar_wait(bar(13 + ((j - 1) & 1)), ((j - 1) >> 1) & 1);
        fence_after();
        float part[DK];
        tmem_ld32(TM_O + lane_addr + ((j - 1) & 1) * 64, part);
        tmem_ld32(TM_O + lane_addr + ((j - 1) & 1) * 64 + 32, part + 32);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < DK; ++d) o[d] = o[d] * corr_prev + part[d];
      }
      l_run = l_run * corr + ((lsv[0] + lsv[1]) + (lsv[2] + lsv[3]));
      m_run = m_new;
      corr_prev = corr;
      fence_before();
    }
    {
      const int j = n_tiles - 1;
      mbar_wait(bar(13 + (j & 1)), (j >> 1) & 1);
      fence_after();
      float part[DK];
      tmem_ld32(TM_O + lane_addr + (j & 1) * 64, part);
      tmem_ld32(TM_O + lane_addr + (j & 1) * 64 + 32, part + 32);
      tmem_wait_ld();
      const float inv = 1.f / l_run;
#pragma unroll
      for (int d = 0; d < DK; ++d) o[d] = (o[d] * corr_prev + part[d]) * inv;
    }
    if (qi < p.Tq) {
      __nv_bfloat16* dst = p.out + (int64_t)b * p.sbo + (int64_t)qi * p.ldo + hh * DK;
#pragma unroll
      for (int d = 0; d < DK; d += 8) {
        __nv_bfloat162 b0 = __floats2bfloat162_rn(o[d], o[d + 1]), b1 = __floats2bfloat162_rn(o[d + 2], o[d + 3]),
                       b2 = __floats2bfloat162_rn(o[d + 4], o[d + 5]), b3 = __floats2bfloat162_rn(o[d + 6], o[d + 7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&b0); pk.y = *reinterpret_cast<uint32_t*>(&b1);
        pk.z = *reinterpret_cast<uint32_t*>(&b2); pk.w = *reinterpret_cast<uint32_t*>(&b3);
        *reinterpret_cast<uint4*>(dst + d) = pk;
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

extern "C" int cn_attention_bf16(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                                 const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb, int64_t mi,
                                 int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk, void* out, int64_t ldo,
                                 int64_t sbo, cn_stream_t stream_) {
  if (!Q || !K || !V || !out || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (dk != DK || Tk > MAX_TILES * KT || h > 65535 || B > 65535) return CN_ERR_UNSUPPORTED;
  // TMA: 16-byte aligned bases and pitches; 16-byte vector stores of the output rows
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (sbq & 7) || (sbk & 7) || (sbv & 7) || (ldo & 7) || (sbo & 7) ||
      ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V) |
        reinterpret_cast<uintptr_t>(out)) & 15))
    return CN_ERR_UNSUPPORTED;
  if (B == 0 || Tq == 0) return CN_OK;
  CUtensorMap mq, mk, mv;
  const int64_t width = h * DK;
  if (!make_map3(&mq, Q, B, Tq, width, ldq, sbq) || !make_map3(&mk, K, B, Tk, width, ldk, sbk) ||
      !make_map3(&mv, V, B, Tk, width, ldv, sbv))
    return CN_ERR_DEVICE;
  AttnTcArgs p;
  p.mask = mask; p.mb = mb; p.mi = mi; p.h = (int)h; p.Tq = (int)Tq; p.Tk = (int)Tk;
  p.out = (__nv_bfloat16*)out; p.ldo = ldo; p.sbo = sbo;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
      return CN_ERR_LAUNCH;
    configured = true;
  }
  dim3 grid((unsigned)((Tq + QT - 1) / QT), (unsigned)h, (unsigned)B);
  attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(mq, mk, mv, p);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
