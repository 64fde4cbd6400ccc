// bf16 tensor-core GEMM for sm_100a: C[M,N] = residual + alpha * act(A[M,K] * W[N,K]^T + bias)
//
//   * persistent, one CTA per SM, warp-specialised (1 TMA producer lane, 1 MMA issuer lane,
//     8 epilogue warps);
//   * operands: bf16, both K-major, staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
//     4-stage shared-memory ring guarded by mbarriers;
//   * math: tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction, fp32 accumulators in
//     TMEM, two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
//     tile i+1;
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> XOR-swizzled shared-memory transpose -> bias /
//     activation / scale / residual -> coalesced fp32 and/or bf16 rows; or (STATS) per-row argmax / runner-up /
//     log-sum-exp / target logit without writing the matrix;
//   * operands may also be MN-major (backward GEMMs), and the same mainloop runs the front-end's 3x3 stride-2
//     convolution as an implicit GEMM (4-D strided TMA boxes).
//
// Replaces (fast mode) the nn.Linear / Conv1d(k=1) / Generator.proj call sites of
// src/models/modules/{attention,positionff,conformer_related}.py and src/models/cassnat.py:105,113.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (rows of A per tile)
constexpr int BK = 64;           // bf16 elements per stage row = 128 bytes = one swizzle atom
constexpr int UK = 16;           // K per tcgen05.mma for 16-bit inputs
constexpr int STAGES = 4;
constexpr int THREADS = 320;     // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..9: epilogue
constexpr int EPI_WARP0 = 2;     // first epilogue warp
constexpr int CONV_THREADS = 320;   // conv2_igemm_kernel keeps the compact layout (warps 2..9 epilogue)

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s at 2 GHz
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1 (sm_100)
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                  // LBO: ignored for a one-atom-wide swizzled K-major tile (CUTLASS sets 1)
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10),
// bit 15 / 16 = A / B is MN-major (0 = K-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// MN-major, 128-byte-swizzled operand: the tile is a row of [64 contraction rows x 64 MN elements] blocks (one TMA
// box each, 8 KB apart = leading byte offset), 8-row groups of a block 1024 B apart (stride byte offset).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Epilogue {
  const float* bias;
  const float* residual; int64_t ldr; int64_t res_rows;
  float alpha; int act;
  float* out_f32; __nv_bfloat16* out_bf16; int64_t ldc;
  int64_t lo_off;        // != 0: out_bf16 is a split (bf16x3) operand, the rounding residue goes lo_off elements further
  // LN mode: the LayerNorm that consumes the fp32 output (next sublayer's pre-norm) is applied by the same kernel
  const float* ln_a; const float* ln_b; float ln_eps; __nv_bfloat16* ln_out; int64_t ld_ln;
  // LayerNorm folding (fast mode), producer side (LNM_PART): besides the fp32 rows (and their raw bf16 copy) the
  // epilogue emits per-row partial sums (sum, sum of squares) of its 128-column halves: part [M][n_parts] float2.
  // Consumer side (LNM_FOLD): A is the RAW bf16 row, W already carries the LayerNorm gain, and the epilogue applies
  //   y = rinv_m * (acc - mean_m * c_n) + bias'_n      (c_n = sum_k W'_nk, bias' = bias + W b)
  // i.e. LayerNorm(x) W^T without a LayerNorm kernel or a normalised copy of x ever existing.
  float* part; int n_rowparts;                             // producer: output; consumer: input (stats of ITS rows)
  const float* fold_c; float fold_eps; int fold_d;         // consumer: column sums of the folded weight, eps, row width
  // STATS mode: per-(row, column-part) partial row statistics instead of (or besides) the matrix
  float* st_max; int32_t* st_idx; float* st_sum; float* st_max2; int32_t* st_idx2; int n_parts;
  const int32_t* st_target; float* st_target_logit;    // optional: the logit of one given column per row
  // CE mode (LNM_CE, fused linear + label-smoothed cross entropy, backward): the tile's logits z = A W^T + bias are
  // turned into the loss gradient  dz = rowscale_m * (exp(z - lse_m) - w),  w = conf at column target_m, eps elsewhere
  // (src/utils/loss.py:21-35: KLDiv against the smoothed one-hot), and only dz (bf16) is written
  const float* ce_lse; const int32_t* ce_target; const float* ce_rowscale; float ce_conf, ce_eps;
};

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if (ACT == CN_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == CN_ACT_SWISH) return __fdividef(v, 1.f + __expf(-v));      // (IEEE division: ~10 more instructions per element)
  return v;
}

constexpr int LNM_NONE = 0, LNM_FUSED = 1, LNM_PART = 2, LNM_FOLD = 3, LNM_CE = 4;

template <int BN, int ACT, bool RES, bool STATS, bool A_MN = false, bool B_MN = false, int LNM = LNM_NONE>
__global__ void __launch_bounds__(THREADS, 1) gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_w,
                                                                  Epilogue ep, int M, int N, int K, int split) {
  constexpr bool LN = LNM == LNM_FUSED;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;                    // two accumulator stages (power of two >= 32)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // keeps the shared address space visible
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int n_tiles = tiles_m * tiles_n;
  // split (bf16x3): both operands are stored [hi | lo] along the contraction axis (2K columns, K % BK == 0); the
  // mainloop walks three parts of K/BK blocks each -- (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo) -- into the same
  // accumulator, i.e. a plain GEMM over a 3K-long contraction whose TMA coordinates are remapped per part
  // split == 2: only W is split (A is a plain bf16 operand, e.g. an attention context): (A, W_hi), (A, W_lo)
  const int k_part = (K + BK - 1) / BK;
  const int k_blocks = split == 1 ? 3 * k_part : (split == 2 ? 2 * k_part : k_part);
  // Tile schedule (every role walks the same list).  Default: tile = blockIdx + i * grid over the n-fastest tile
  // index.  LN: a CTA owns whole 128-row blocks (all N tiles of a block back to back), because the fused LayerNorm
  // needs the complete rows: the row statistics are accumulated across the block's N tiles by the same warps.
  const int n_my = LN ? (((tiles_m - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * tiles_n)
                      : ((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
  auto tile_of = [&](int i) {
    return LN ? ((int)blockIdx.x + (i / tiles_n) * (int)gridDim.x) * tiles_n + (i % tiles_n) : (int)blockIdx.x + i * (int)gridDim.x;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cn_pdl_wait();                 // (prologue above overlapped the predecessor's tail)
  cn_pdl_launch();
  // (setmaxnreg re-balancing between the single-lane roles and the epilogue warps was tried: ptxas then spills in
  // every instantiation, with or without role-pure warpgroups; the epilogue stays inside the common 168 registers)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        const int tile = tile_of(ti);
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          // K-major operand: one [rows x 64 k] box.  MN-major operand (the global tensor is [contraction, MN]
          // row-major, used as is by the backward GEMMs): one [64 k x 64 mn] box per 64 MN elements.
          int ka = kb, kw = kb;
          if (split) {
            const int part = kb / k_part, r = kb - part * k_part;
            if (split == 1) {
              ka = part == 1 ? k_part + r : r;                  // A: hi, lo, hi
              kw = part == 2 ? k_part + r : r;                  // W: hi, hi, lo
            } else {
              ka = r;                                           // A: plain
              kw = part == 1 ? k_part + r : r;                  // W: hi, lo
            }
          }
          if (!A_MN) tma_load_2d(&map_a, &full[stage], sa, ka * BK, m0);
          else
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(&map_a, &full[stage], sa + i * 8192, m0 + i * 64, ka * BK);
          if (!B_MN) tma_load_2d(&map_w, &full[stage], sa + A_BYTES, kw * BK, n0);
          else
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(&map_w, &full[stage], sa + A_BYTES + i * 8192, n0 + i * 64, kw * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);             // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);                       // TMA bytes have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
          const uint64_t da = A_MN ? make_smem_desc_mn(sa) : make_smem_desc(sa);
          const uint64_t db = B_MN ? make_smem_desc_mn(sa + A_BYTES) : make_smem_desc(sa + A_BYTES);
          // advancing 16 contraction steps: +32 bytes inside the swizzle atom (K-major) / +16 rows of 128 B (MN-major)
          constexpr uint32_t ka = A_MN ? (UK * 128) >> 4 : (UK * 2) >> 4, kw = B_MN ? (UK * 128) >> 4 : (UK * 2) >> 4;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_bf16(tmem_d, da + (uint64_t)(k * ka), db + (uint64_t)(k * kw), idesc, (kb | k) ? 1u : 0u);
          tcgen05_commit(&empty[stage]);                        // frees the smem slot when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&acc_full[acc]);                         // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (warps 2..9) =====================
    // TMEM lane quarter = warp % 4 (hardware rule); the two warps of a quarter split the BN columns.
    // TMEM hands every thread one ROW (32 consecutive columns).  Writing rows straight to global
    // memory costs 32 L1 wavefronts per store instruction, so each warp transposes its 32x32 chunk
    // through a private padded smem buffer and stores 4 rows x 128 contiguous bytes per instruction.
    const int quarter = warp & 3, half = (warp - EPI_WARP0) >> 2;
    if constexpr (STATS) {
      // Fused row statistics (K1/K9): every thread owns one output ROW of the tile (its TMEM lane) and
      // scans its half of the tile's columns: running (max, first index), runner-up and online
      // sum of exp.  Partials go to [row][part] arrays; a tiny second kernel merges the parts.
      // The logits themselves never leave the SM.
      float* bias_s = reinterpret_cast<float*>(tiles + STAGES * STAGE_BYTES + 256);       // [2][BN]
      const int tid_e = threadIdx.x - EPI_WARP0 * 32;
      uint32_t acc = 0, acc_phase = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        const int tile = tile_of(ti);
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        if (tid_e < BN) bias_s[acc * BN + tid_e] = (ep.bias && n0 + tid_e < N) ? ep.bias[n0 + tid_e] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait(&acc_full[acc], acc_phase);
        tcgen05_fence_after();
        const int row = m0 + quarter * 32 + lane;
        float m1 = -INFINITY, m2 = -INFINITY, ssum = 0.f;
        int i1 = 0x7fffffff, i2 = 0x7fffffff;
        const int tgt = (ep.st_target && row < M) ? ep.st_target[row] : -1;
#pragma unroll 1
        for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
          const int col0 = n0 + c0;
          if (col0 >= N) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c0, v);
          const float* bz = bias_s + acc * BN + c0;
          float x[32];
          float cm = -INFINITY;
          const float m_old = m1;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bz + j);
            x[j] = (__uint_as_float(v[j]) + b4.x) * ep.alpha;
            x[j + 1] = (__uint_as_float(v[j + 1]) + b4.y) * ep.alpha;
            x[j + 2] = (__uint_as_float(v[j + 2]) + b4.z) * ep.alpha;
            x[j + 3] = (__uint_as_float(v[j + 3]) + b4.w) * ep.alpha;
          }
          if (col0 + 32 > N) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j >= N) x[j] = -INFINITY;
          }
          if (tgt >= col0 && tgt < col0 + 32) {                 // exactly one (tile, half, chunk) owns the target column
            float tl = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) tl = (col0 + j == tgt) ? x[j] : tl;
            ep.st_target_logit[row] = tl;
          }
          if (ep.st_idx2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (x[j] > m1) { m2 = m1; i2 = i1; m1 = x[j]; i1 = col0 + j; }
              else if (x[j] > m2) { m2 = x[j]; i2 = col0 + j; }
              cm = fmaxf(cm, x[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = fmaxf(cm, x[j]);
            if (cm > m1) {                                   // strictly greater: earlier columns keep ties
              int ci = 31;
#pragma unroll
              for (int j = 30; j >= 0; --j) ci = (x[j] == cm) ? j : ci;
              i1 = col0 + ci;
            }
          }
          const float m_new = fmaxf(m_old, cm);
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) part += __expf(x[j] - m_new);
          ssum = ssum * __expf(m_old - m_new) + part;          // m_old: the max ssum was relative to so far
          m1 = m_new;
        }
        if (row < M) {
          const int64_t o = (int64_t)row * ep.n_parts + (tile % tiles_n) * 2 + half;
          ep.st_max[o] = m1; ep.st_idx[o] = i1; ep.st_sum[o] = ssum;
          if (ep.st_idx2) { ep.st_max2[o] = m2; ep.st_idx2[o] = i2; }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else {
    const uint32_t stg = smem_u32(tiles + STAGES * STAGE_BYTES + 256) + (uint32_t)(warp - EPI_WARP0) * 4096u;   // 32 x 32 fp32, XOR-swizzled 16-byte chunks
    const int rr = lane >> 3, cc = (lane & 7) * 4;
    const bool vec_ok = ((ep.ldc & 3) == 0) && (!ep.residual || (ep.ldr & 3) == 0);
    uint32_t acc = 0, acc_phase = 0;
    float ln_s1[8], ln_s2[8];            // LN: sum / sum of squares of this lane's columns of its 8 rows (all N tiles)
#pragma unroll
    for (int it = 0; it < 8; ++it) ln_s1[it] = ln_s2[it] = 0.f;
    for (int ti = 0; ti < n_my; ++ti) {
      const int tile = tile_of(ti);
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int row0 = m0 + quarter * 32 + rr;                  // first of the 8 rows (stride 4) this lane stores
      // Residual rows do not depend on the accumulator: fetch them one chunk ahead so their HBM
      // latency hides behind the MMAs / the previous chunk instead of stalling the epilogue.
      // (Measured alternatives, both slower: a continuous stream two chunks ahead across tile boundaries with the
      // chunk loop unrolled, 35 -> 47 us on [32000x512x512]; setmaxnreg re-balancing, which makes ptxas spill.)
      auto load_res = [&](int c0, float4 (&r)[8]) {
        const int gc = n0 + c0 + cc;
        const bool ok = RES && ep.residual && vec_ok && gc + 3 < N;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row0 + it * 4;
          if (ok && grow < M) {
            const int64_t rrow = ep.res_rows > 0 ? (int64_t)(grow % (int)ep.res_rows) : (int64_t)grow;
            r[it] = *reinterpret_cast<const float4*>(ep.residual + rrow * ep.ldr + gc);
          } else {
            r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      float4 res_cur[8], res_next[8];
      if (RES) load_res(half * (BN / 2), res_cur);
      float f_lse[8], f_rs[8];                                  // LNM_CE: log-sum-exp / gradient scale of this lane's 8 rows
      int f_tgt[8];
      if (LNM == LNM_CE) {
        const int myrow = m0 + quarter * 32 + lane;
        const float lse = myrow < M ? __ldg(ep.ce_lse + myrow) : 0.f, rs = myrow < M ? __ldg(ep.ce_rowscale + myrow) : 0.f;
        const int tg = myrow < M ? __ldg(ep.ce_target + myrow) : -1;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          f_lse[it] = __shfl_sync(0xffffffffu, lse, it * 4 + rr);
          f_rs[it] = __shfl_sync(0xffffffffu, rs, it * 4 + rr);
          f_tgt[it] = __shfl_sync(0xffffffffu, tg, it * 4 + rr);
        }
      }
      float f_mean[8], f_rinv[8];                               // LNM_FOLD: LayerNorm statistics of this lane's 8 rows
      if (LNM == LNM_FOLD) {
        // lane l finalises row (quarter*32 + l) from the producer's partial sums, then the 8 rows a lane stores
        // (rows rr + 4*it of the quarter) are fetched by shuffle
        const int myrow = m0 + quarter * 32 + lane;
        float s1 = 0.f, s2 = 0.f;
        if (myrow < M) {
          const float2* pp = reinterpret_cast<const float2*>(ep.part) + (int64_t)myrow * ep.n_rowparts;
          for (int q = 0; q < ep.n_rowparts; ++q) { const float2 v2 = __ldg(pp + q); s1 += v2.x; s2 += v2.y; }
        }
        const float mean = s1 / (float)ep.fold_d;
        const float var = fmaxf(s2 - s1 * mean, 0.f) / (float)(ep.fold_d - 1);
        const float rinv = 1.f / (sqrtf(var) + ep.fold_eps);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          f_mean[it] = __shfl_sync(0xffffffffu, mean, it * 4 + rr);
          f_rinv[it] = __shfl_sync(0xffffffffu, rinv, it * 4 + rr);
        }
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        const int col0 = n0 + c0;
        if (col0 >= N) break;                                   // warp-uniform
        if (RES && c0 + 32 < (half + 1) * (BN / 2) && col0 + 32 < N) load_res(c0 + 32, res_next);
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + (uint32_t)(lane * 128 + (((j >> 2) ^ (lane & 7)) << 4))), "r"(v[j]),
                       "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3])
                       : "memory");
        __syncwarp();
        const int gcol = col0 + cc;
        float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias) {
          if (gcol + 3 < N) bz = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol));
          else {
            if (gcol < N) bz.x = ep.bias[gcol];
            if (gcol + 1 < N) bz.y = ep.bias[gcol + 1];
            if (gcol + 2 < N) bz.z = ep.bias[gcol + 2];
          }
        }
        float x[8][4];
#pragma unroll
        for (int it = 0; it < 8; ++it)
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(x[it][0]), "=f"(x[it][1]), "=f"(x[it][2]), "=f"(x[it][3])
                       : "r"(stg + (uint32_t)((it * 4 + rr) * 128 + ((((lane & 7)) ^ ((it * 4 + rr) & 7)) << 4))));
        if (LNM == LNM_FOLD) {
          const float4 c4 = gcol + 3 < N ? __ldg(reinterpret_cast<const float4*>(ep.fold_c + gcol)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            x[it][0] = f_rinv[it] * fmaf(-f_mean[it], c4.x, x[it][0]);
            x[it][1] = f_rinv[it] * fmaf(-f_mean[it], c4.y, x[it][1]);
            x[it][2] = f_rinv[it] * fmaf(-f_mean[it], c4.z, x[it][2]);
            x[it][3] = f_rinv[it] * fmaf(-f_mean[it], c4.w, x[it][3]);
          }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          x[it][0] = act_fn<ACT>(x[it][0] + bz.x) * ep.alpha;
          x[it][1] = act_fn<ACT>(x[it][1] + bz.y) * ep.alpha;
          x[it][2] = act_fn<ACT>(x[it][2] + bz.z) * ep.alpha;
          x[it][3] = act_fn<ACT>(x[it][3] + bz.w) * ep.alpha;
        }
        if (LNM == LNM_CE) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              x[it][q] = f_rs[it] * (__expf(x[it][q] - f_lse[it]) - (gcol + q == f_tgt[it] ? ep.ce_conf : ep.ce_eps));
        }
        if (vec_ok && gcol + 3 < N) {
          if (RES) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              x[it][0] += res_cur[it].x; x[it][1] += res_cur[it].y; x[it][2] += res_cur[it].z; x[it][3] += res_cur[it].w;
            }
          }
          if (LN || LNM == LNM_PART) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              ln_s1[it] += (x[it][0] + x[it][1]) + (x[it][2] + x[it][3]);
              ln_s2[it] = fmaf(x[it][0], x[it][0], fmaf(x[it][1], x[it][1], fmaf(x[it][2], x[it][2], fmaf(x[it][3], x[it][3], ln_s2[it]))));
            }
          }
          if (ep.out_f32) {
            float* o = ep.out_f32 + (int64_t)row0 * ep.ldc + gcol;
#pragma unroll
            for (int it = 0; it < 8; ++it, o += 4 * ep.ldc)
              if (row0 + it * 4 < M) *reinterpret_cast<float4*>(o) = make_float4(x[it][0], x[it][1], x[it][2], x[it][3]);
          }
          if (ep.out_bf16) {
            __nv_bfloat16* o = ep.out_bf16 + (int64_t)row0 * ep.ldc + gcol;
#pragma unroll
            for (int it = 0; it < 8; ++it, o += 4 * ep.ldc) {
              if (row0 + it * 4 < M) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(x[it][0], x[it][1]), hi = __floats2bfloat162_rn(x[it][2], x[it][3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(o) = pk;
                if (ep.lo_off) {                               // rounding residue of a split (bf16x3) operand
                  const float2 f0 = __bfloat1622float2(lo), f1 = __bfloat1622float2(hi);
                  __nv_bfloat162 r0 = __floats2bfloat162_rn(x[it][0] - f0.x, x[it][1] - f0.y),
                                 r1 = __floats2bfloat162_rn(x[it][2] - f1.x, x[it][3] - f1.y);
                  pk.x = *reinterpret_cast<uint32_t*>(&r0);
                  pk.y = *reinterpret_cast<uint32_t*>(&r1);
                  *reinterpret_cast<uint2*>(o + ep.lo_off) = pk;
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int grow = row0 + it * 4;
            if (grow >= M) continue;
            const int64_t rrow = ep.res_rows > 0 ? (int64_t)(grow % (int)ep.res_rows) : (int64_t)grow;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (gcol + q < N) {
                float y = x[it][q];
                if (ep.residual) y += ep.residual[rrow * ep.ldr + gcol + q];
                if (ep.out_f32) ep.out_f32[(int64_t)grow * ep.ldc + gcol + q] = y;
                if (ep.out_bf16) {
                  const __nv_bfloat16 yh = __float2bfloat16(y);
                  ep.out_bf16[(int64_t)grow * ep.ldc + gcol + q] = yh;
                  if (ep.lo_off) ep.out_bf16[(int64_t)grow * ep.ldc + gcol + q + ep.lo_off] = __float2bfloat16(y - __bfloat162float(yh));
                }
              }
            }
          }
        }
        __syncwarp();                                            // staging buffer is reused by the next chunk
        if (RES) {
#pragma unroll
          for (int it = 0; it < 8; ++it) res_cur[it] = res_next[it];
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (LNM == LNM_PART) {
        // row partial sums of this warp's 128-column half of the tile -> part[row][tile_n * 2 + half]
#pragma unroll
        for (int it = 0; it < 8; ++it) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            ln_s1[it] += __shfl_xor_sync(0xffffffffu, ln_s1[it], o);
            ln_s2[it] += __shfl_xor_sync(0xffffffffu, ln_s2[it], o);
          }
          if ((lane & 7) == 0 && row0 + it * 4 < M)
            reinterpret_cast<float2*>(ep.part)[(int64_t)(row0 + it * 4) * ep.n_rowparts + (tile % tiles_n) * 2 + half] =
                make_float2(ln_s1[it], ln_s2[it]);
          ln_s1[it] = ln_s2[it] = 0.f;
        }
      }
      if (LN && (tile % tiles_n) == tiles_n - 1) {
        // ---- fused LayerNorm of the 128 finished rows (modules/norm.py:15-18: unbiased std, eps added to std).
        // Row totals: over this lane's columns (registers), over the 8 lanes that share a row (shuffles), over the
        // two warps that share the lane quarter (staging buffer + a 64-thread named barrier).  The MMA warp is
        // already filling the next block's accumulator while this runs.
#pragma unroll
        for (int it = 0; it < 8; ++it) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            ln_s1[it] += __shfl_xor_sync(0xffffffffu, ln_s1[it], o);
            ln_s2[it] += __shfl_xor_sync(0xffffffffu, ln_s2[it], o);
          }
        }
        if ((lane & 7) == 0) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(stg + (uint32_t)((it * 4 + rr) * 8)), "f"(ln_s1[it]), "f"(ln_s2[it]) : "memory");
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        const uint32_t stg_peer = smem_u32(tiles + STAGES * STAGE_BYTES + 256) + (uint32_t)((warp - EPI_WARP0) ^ 4) * 4096u;
        float mean[8], rinv[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float a1, a2, b1, b2;
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a1), "=f"(a2) : "r"(stg + (uint32_t)((it * 4 + rr) * 8)));
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(b1), "=f"(b2) : "r"(stg_peer + (uint32_t)((it * 4 + rr) * 8)));
          const float s1 = a1 + b1, s2 = a2 + b2;
          mean[it] = s1 / (float)N;
          const float var = fmaxf(s2 - s1 * mean[it], 0.f) / (float)(N - 1);
          rinv[it] = 1.f / (sqrtf(var) + ep.ln_eps);
          ln_s1[it] = ln_s2[it] = 0.f;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");     // both warps have read: staging is free again
        // second pass over this lane's own outputs (just written: L2-resident; a thread sees its own stores)
#pragma unroll 1
        for (int nt = 0; nt < tiles_n; ++nt) {
#pragma unroll 1
          for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
            const int gcol = nt * BN + c0 + cc;
            if (gcol + 3 >= N) continue;                             // (N % 4 == 0: a lane is inside or outside)
            const float4 ga = __ldg(reinterpret_cast<const float4*>(ep.ln_a + gcol));
            const float4 be = __ldg(reinterpret_cast<const float4*>(ep.ln_b + gcol));
            float4 xv[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int grow = row0 + it * 4;
              xv[it] = grow < M ? *reinterpret_cast<const float4*>(ep.out_f32 + (int64_t)grow * ep.ldc + gcol)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int grow = row0 + it * 4;
              if (grow >= M) continue;
              const float y0 = fmaf(ga.x * (xv[it].x - mean[it]), rinv[it], be.x), y1 = fmaf(ga.y * (xv[it].y - mean[it]), rinv[it], be.y);
              const float y2 = fmaf(ga.z * (xv[it].z - mean[it]), rinv[it], be.z), y3 = fmaf(ga.w * (xv[it].w - mean[it]), rinv[it], be.w);
              __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(ep.ln_out + (int64_t)grow * ep.ld_ln + gcol) = pk;
            }
          }
        }
      }
    }
    }  // !STATS
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major [rows, K] (row pitch ld elements) -> 2-D map with a [BK x box_rows] box, 128B swizzle
bool make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_sm_limit = 0;          // cn_set_sm_limit: CTAs of the persistent GEMMs (0 = one per SM)
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

template <int BN, int ACT, bool RES, bool STATS = false, bool A_MN = false, bool B_MN = false, int LNM = LNM_NONE>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mw, const Epilogue& ep, int M, int N, int K,
              cudaStream_t stream, int split = 0) {
  constexpr size_t smem = 1024 + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 256 + 8 * 4096;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, ACT, RES, STATS, A_MN, B_MN, LNM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CN_ERR_LAUNCH;
    configured = true;
  }
  const int tiles = LNM == LNM_FUSED ? (M + BM - 1) / BM : ((M + BM - 1) / BM) * ((N + BN - 1) / BN);     // fused LN: whole row blocks per CTA
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (cn_launch_pdl(gemm_bf16_tc_kernel<BN, ACT, RES, STATS, A_MN, B_MN, LNM>, dim3(grid), dim3(THREADS), smem, stream, ma, mw, ep, M, N, K,
                    split) != cudaSuccess)
    return CN_ERR_LAUNCH;
  return cudaGetLastError() == cudaSuccess ? CN_OK : CN_ERR_LAUNCH;
}

}  // namespace

struct LnArgs { const float* a; const float* b; float eps; void* out; int64_t ld; };
// LayerNorm folding: producer (part_out != NULL) or consumer (part_in != NULL)
struct FoldArgs { float* part_out; const float* part_in; int n_parts; const float* c; float eps; int d; };

static int gemm_tile_n(int64_t N, int64_t K, int split) { return (N > 128 && !(N <= 512 && K <= 512 && !split)) ? 256 : 128; }

static int gemm_bf16_impl(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                         const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act,
                         float* out_f32, void* out_bf16, int64_t ldc, int64_t M, int64_t N, int64_t K,
                         int split, int out_split, cn_stream_t stream_, const LnArgs* ln = nullptr,
                         const FoldArgs* fold = nullptr) {
  const int64_t kcols = split ? 2 * K : K, acols = split == 1 ? 2 * K : K;      // stored columns of W and A
  if (!A || !W || (!out_f32 && !out_bf16) || M < 0 || N < 0 || K <= 0 || lda < acols || ldw < kcols || ldc < N)
    return CN_ERR_BAD_ARG;
  if (residual && ldr < N) return CN_ERR_BAD_ARG;
  if (split && (K % BK)) return CN_ERR_UNSUPPORTED;            // the hi | lo boundary must fall on a k-block
  if (out_split && (!out_bf16 || out_f32 || ldc < 2 * N || (N & 3))) return CN_ERR_BAD_ARG;
  // TMA needs 16-byte aligned bases and row pitches
  if ((lda & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return CN_ERR_UNSUPPORTED;
  if (M > 0x7fffff00 || N > 0x7fffff00 || K > 0x3fffff00) return CN_ERR_UNSUPPORTED;
  if (M == 0 || N == 0) return CN_OK;
  // 128x256 tiles unless the problem is the HBM-bound d x d projection shape (N, K <= 512): there the
  // epilogue streams more bytes than the mainloop computes, and 128x128 tiles fill the last wave of the
  // 148 persistent CTAs better (measured on [32000x512x512]: 38.8 -> 34.9 us)
  int bn = gemm_tile_n(N, K, split);
  if (!fold && !ln) {
    // few rows (the small model, the label-token side of training): narrower tiles until the persistent grid has a
    // tile for every SM -- at [4000 x 256 x 2048] a 128x256 tiling keeps 32 of 148 SMs busy
    const int64_t mblocks = (M + BM - 1) / BM;
    while (bn > 64 && mblocks * ((N + bn - 1) / bn) < sm_count()) bn /= 2;
  }
  if (fold) {
    // vector epilogue only; the producer writes fp32 rows (+ raw bf16 copy), the consumer has no residual
    if (ln || split || out_split || (N & 3) || (ldc & 3) || (residual && (ldr & 3))) return CN_ERR_UNSUPPORTED;
    if (fold->part_out && (act != CN_ACT_NONE || !out_f32)) return CN_ERR_BAD_ARG;
    if (fold->part_in && (residual || !fold->c || fold->d < 2 || fold->n_parts < 1)) return CN_ERR_BAD_ARG;
  }
  if (ln) {
    // fused LayerNorm: a CTA owns complete rows as at most two N tiles; vector epilogue only
    if (!ln->a || !ln->b || !ln->out || !out_f32 || out_bf16 || act != CN_ACT_NONE || ln->ld < N) return CN_ERR_BAD_ARG;
    if (N > 512 || N < 2 || (N & 3) || (ldc & 3) || (ln->ld & 3) || (residual && (ldr & 3)) ||
        (reinterpret_cast<uintptr_t>(out_f32) & 15) || (reinterpret_cast<uintptr_t>(ln->out) & 7) ||
        (reinterpret_cast<uintptr_t>(ln->a) & 15) || (reinterpret_cast<uintptr_t>(ln->b) & 15) ||
        (residual && (reinterpret_cast<uintptr_t>(residual) & 15)))
      return CN_ERR_UNSUPPORTED;
    bn = N > 256 ? 256 : 128;
  }
  CUtensorMap ma, mw;
  if (!make_map(&ma, A, M, acols, lda, BM) || !make_map(&mw, W, N, kcols, ldw, bn)) return CN_ERR_DEVICE;
  Epilogue ep;
  ep.bias = bias; ep.residual = residual; ep.ldr = ldr; ep.res_rows = res_rows; ep.alpha = alpha; ep.act = act;
  ep.out_f32 = out_f32; ep.out_bf16 = (__nv_bfloat16*)out_bf16; ep.ldc = ldc; ep.lo_off = out_split ? N : 0;
  ep.ln_a = ln ? ln->a : nullptr; ep.ln_b = ln ? ln->b : nullptr; ep.ln_eps = ln ? ln->eps : 0.f;
  ep.ln_out = ln ? (__nv_bfloat16*)ln->out : nullptr; ep.ld_ln = ln ? ln->ld : 0;
  ep.part = nullptr; ep.n_rowparts = 0; ep.fold_c = nullptr; ep.fold_eps = 0.f; ep.fold_d = 0;
  ep.st_max = nullptr; ep.st_idx = nullptr; ep.st_sum = nullptr; ep.st_max2 = nullptr; ep.st_idx2 = nullptr; ep.n_parts = 0;
  ep.st_target = nullptr; ep.st_target_logit = nullptr;
  ep.ce_lse = nullptr; ep.ce_target = nullptr; ep.ce_rowscale = nullptr; ep.ce_conf = 0.f; ep.ce_eps = 0.f;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int m = (int)M, n = (int)N, k = (int)K;
  if (fold && fold->part_out) {
    ep.part = fold->part_out; ep.n_rowparts = 2 * (int)((N + bn - 1) / bn);
    if (bn == 256) {
      if (residual) return launch_tc<256, CN_ACT_NONE, true, false, false, false, LNM_PART>(ma, mw, ep, m, n, k, stream, split);
      return launch_tc<256, CN_ACT_NONE, false, false, false, false, LNM_PART>(ma, mw, ep, m, n, k, stream, split);
    }
    if (residual) return launch_tc<128, CN_ACT_NONE, true, false, false, false, LNM_PART>(ma, mw, ep, m, n, k, stream, split);
    return launch_tc<128, CN_ACT_NONE, false, false, false, false, LNM_PART>(ma, mw, ep, m, n, k, stream, split);
  }
  if (fold && fold->part_in) {
    ep.part = const_cast<float*>(fold->part_in); ep.n_rowparts = fold->n_parts; ep.fold_c = fold->c; ep.fold_eps = fold->eps;
    ep.fold_d = fold->d;
#define CN_TC_FOLD(BN_)                                                                                                        \
    do {                                                                                                                        \
      if (act == CN_ACT_RELU) return launch_tc<BN_, CN_ACT_RELU, false, false, false, false, LNM_FOLD>(ma, mw, ep, m, n, k, stream, split);   \
      if (act == CN_ACT_SWISH) return launch_tc<BN_, CN_ACT_SWISH, false, false, false, false, LNM_FOLD>(ma, mw, ep, m, n, k, stream, split); \
      return launch_tc<BN_, CN_ACT_NONE, false, false, false, false, LNM_FOLD>(ma, mw, ep, m, n, k, stream, split);                           \
    } while (0)
    if (bn == 256) CN_TC_FOLD(256);
    CN_TC_FOLD(128);
#undef CN_TC_FOLD
  }
  if (ln) {
    if (bn == 256) {
      if (residual) return launch_tc<256, CN_ACT_NONE, true, false, false, false, LNM_FUSED>(ma, mw, ep, m, n, k, stream, split);
      return launch_tc<256, CN_ACT_NONE, false, false, false, false, LNM_FUSED>(ma, mw, ep, m, n, k, stream, split);
    }
    if (residual) return launch_tc<128, CN_ACT_NONE, true, false, false, false, LNM_FUSED>(ma, mw, ep, m, n, k, stream, split);
    return launch_tc<128, CN_ACT_NONE, false, false, false, false, LNM_FUSED>(ma, mw, ep, m, n, k, stream, split);
  }
#define CN_TC_DISPATCH(BN_)                                                                                     \
  do {                                                                                                          \
    if (residual) {                                                                                             \
      if (act == CN_ACT_RELU) return launch_tc<BN_, CN_ACT_RELU, true>(ma, mw, ep, m, n, k, stream, split);    \
      if (act == CN_ACT_SWISH) return launch_tc<BN_, CN_ACT_SWISH, true>(ma, mw, ep, m, n, k, stream, split);  \
      return launch_tc<BN_, CN_ACT_NONE, true>(ma, mw, ep, m, n, k, stream, split);                            \
    }                                                                                                           \
    if (act == CN_ACT_RELU) return launch_tc<BN_, CN_ACT_RELU, false>(ma, mw, ep, m, n, k, stream, split);     \
    if (act == CN_ACT_SWISH) return launch_tc<BN_, CN_ACT_SWISH, false>(ma, mw, ep, m, n, k, stream, split);   \
    return launch_tc<BN_, CN_ACT_NONE, false>(ma, mw, ep, m, n, k, stream, split);                             \
  } while (0)
  if (bn == 256) CN_TC_DISPATCH(256);
  if (bn == 64) CN_TC_DISPATCH(64);
  CN_TC_DISPATCH(128);
#undef CN_TC_DISPATCH
}

extern "C" int cn_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                            const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act,
                            float* out_f32, void* out_bf16, int64_t ldc, int64_t M, int64_t N, int64_t K,
                            cn_stream_t stream_) {
  return gemm_bf16_impl(A, lda, W, ldw, bias, residual, ldr, res_rows, alpha, act, out_f32, out_bf16, ldc, M, N, K, 0, 0,
                        stream_);
}

// Backward half of the fused linear + label-smoothed cross entropy: recompute the logits tile by
// tile and write ONLY their loss gradient, in bf16 -- the operand of the dX / dW GEMMs.  The forward half is
// cn_gemm_bf16_rowstats (log-sum-exp and target logit per row, no logits written) plus the row sums of the logits,
// which are a matrix-vector product with the column sums of W.
extern "C" int cn_gemm_bf16_ce_grad(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* lse,
                                    const int32_t* target, const float* rowscale, float conf, float eps, void* dz_bf16,
                                    int64_t ldc, int64_t M, int64_t N, int64_t K, cn_stream_t stream_) {
  if (!A || !W || !lse || !target || !rowscale || !dz_bf16 || M < 0 || N < 0 || K <= 0 || lda < K || ldw < K || ldc < N)
    return CN_ERR_BAD_ARG;
  if ((lda & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return CN_ERR_UNSUPPORTED;
  if (M > 0x7fffff00 || N > 0x7fffff00 || K > 0x3fffff00) return CN_ERR_UNSUPPORTED;
  if (M == 0 || N == 0) return CN_OK;
  CUtensorMap ma, mw;
  if (!make_map(&ma, A, M, K, lda, BM) || !make_map(&mw, W, N, K, ldw, 256)) return CN_ERR_DEVICE;
  Epilogue ep;
  ep.bias = bias; ep.residual = nullptr; ep.ldr = 0; ep.res_rows = 0; ep.alpha = 1.f; ep.act = CN_ACT_NONE;
  ep.out_f32 = nullptr; ep.out_bf16 = (__nv_bfloat16*)dz_bf16; ep.ldc = ldc; ep.lo_off = 0;
  ep.ln_a = nullptr; ep.ln_b = nullptr; ep.ln_eps = 0.f; ep.ln_out = nullptr; ep.ld_ln = 0;
  ep.part = nullptr; ep.n_rowparts = 0; ep.fold_c = nullptr; ep.fold_eps = 0.f; ep.fold_d = 0;
  ep.st_max = nullptr; ep.st_idx = nullptr; ep.st_sum = nullptr; ep.st_max2 = nullptr; ep.st_idx2 = nullptr; ep.n_parts = 0;
  ep.st_target = nullptr; ep.st_target_logit = nullptr;
  ep.ce_lse = lse; ep.ce_target = target; ep.ce_rowscale = rowscale; ep.ce_conf = conf; ep.ce_eps = eps;
  return launch_tc<256, CN_ACT_NONE, false, false, false, false, LNM_CE>(ma, mw, ep, (int)M, (int)N, (int)K, (cudaStream_t)stream_, 0);
}

extern "C" int cn_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                               const float* residual, int64_t ldr, int64_t res_rows, float alpha, float* out_f32,
                               int64_t ldc, const float* ln_a, const float* ln_b, float ln_eps, void* ln_out,
                               int64_t ld_ln, int64_t M, int64_t N, int64_t K, cn_stream_t stream_) {
  LnArgs ln{ln_a, ln_b, ln_eps, ln_out, ld_ln};
  return gemm_bf16_impl(A, lda, W, ldw, bias, residual, ldr, res_rows, alpha, CN_ACT_NONE, out_f32, nullptr, ldc, M, N, K, 0, 0,
                        stream_, &ln);
}

extern "C" int64_t cn_gemm_rowpart_count(int64_t N, int64_t K) {
  const int bn = gemm_tile_n(N, K, 0);
  return 2 * ((N + bn - 1) / bn);
}

extern "C" int cn_gemm_bf16_rowpart(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                    const float* residual, int64_t ldr, int64_t res_rows, float alpha, float* out_f32,
                                    void* out_bf16, int64_t ldc, float* part, int64_t M, int64_t N, int64_t K,
                                    cn_stream_t stream_) {
  if (!part) return CN_ERR_BAD_ARG;
  FoldArgs f{part, nullptr, 0, nullptr, 0.f, 0};
  return gemm_bf16_impl(A, lda, W, ldw, bias, residual, ldr, res_rows, alpha, CN_ACT_NONE, out_f32, out_bf16, ldc, M, N, K, 0, 0,
                        stream_, nullptr, &f);
}

extern "C" int cn_gemm_bf16_folded(const void* A_raw, int64_t lda, const void* W_folded, int64_t ldw, const float* bias_folded,
                                   const float* colsum, const float* part, int64_t n_parts, int64_t ln_width, float ln_eps,
                                   float alpha, int act, float* out_f32, void* out_bf16, int64_t ldc, int64_t M, int64_t N,
                                   int64_t K, cn_stream_t stream_) {
  if (!part || !colsum) return CN_ERR_BAD_ARG;
  FoldArgs f{nullptr, part, (int)n_parts, colsum, ln_eps, (int)ln_width};
  return gemm_bf16_impl(A_raw, lda, W_folded, ldw, bias_folded, nullptr, 0, 0, alpha, act, out_f32, out_bf16, ldc, M, N, K, 0, 0,
                        stream_, nullptr, &f);
}

extern "C" int cn_gemm_bf16x3(const void* A, int64_t lda, int a_split, const void* W, int64_t ldw, const float* bias,
                              const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act,
                              float* out_f32, void* out_bf16, int out_split, int64_t ldc, int64_t M, int64_t N, int64_t K,
                              cn_stream_t stream_) {
  return gemm_bf16_impl(A, lda, W, ldw, bias, residual, ldr, res_rows, alpha, act, out_f32, out_bf16, ldc, M, N, K,
                        a_split ? 1 : 2, out_split, stream_);
}

// ---------------------------------------------------------------- backward GEMMs without transposed copies
// C[M,N] = op(A) op(W)^T with either operand given in its "contraction-major" storage:
//   a_mn = 0: A is [M, K] (row pitch lda)        a_mn = 1: A is [K, M] (row pitch lda)  -> C = A^T ...
//   w_mn = 0: W is [N, K] (row pitch ldw)        w_mn = 1: W is [K, N] (row pitch ldw)
// dX = dY W      -> A = dY [M,N'] (a_mn 0), W = weight [N',K'] stored [contraction, out] (w_mn 1)
// dW = dY^T X    -> A = dY [M',N] stored [contraction, rows] (a_mn 1), W = X [M',K] (w_mn 1)
namespace {
bool make_map_mn(CUtensorMap* map, const void* base, int64_t kc, int64_t mn, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)kc};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" int cn_gemm_bf16_mn(const void* A, int64_t lda, int a_mn, const void* W, int64_t ldw, int w_mn, float* out_f32,
                               void* out_bf16, int64_t ldc, int64_t M, int64_t N, int64_t K, cn_stream_t stream_) {
  if (!A || !W || (!out_f32 && !out_bf16) || M < 0 || N < 0 || K <= 0 || ldc < N) return CN_ERR_BAD_ARG;
  if (lda < (a_mn ? M : K) || ldw < (w_mn ? N : K)) return CN_ERR_BAD_ARG;
  if (!w_mn) return CN_ERR_UNSUPPORTED;                        // (K-major W: cn_gemm_bf16 covers it when a_mn == 0)
  if ((lda & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return CN_ERR_UNSUPPORTED;
  if (M > 0x7fffff00 || N > 0x7fffff00 || K > 0x7fffff00) return CN_ERR_UNSUPPORTED;
  if (M == 0 || N == 0) return CN_OK;
  int bn = N > 128 ? 256 : 128;
  {
    const int64_t mblocks = (M + BM - 1) / BM;              // (few output rows: narrower tiles, see gemm_bf16_impl)
    while (bn > 64 && mblocks * ((N + bn - 1) / bn) < sm_count()) bn /= 2;
  }
  CUtensorMap ma, mw;
  const bool ok_a = a_mn ? make_map_mn(&ma, A, K, M, lda) : make_map(&ma, A, M, K, lda, BM);
  if (!ok_a || !make_map_mn(&mw, W, K, N, ldw)) return CN_ERR_DEVICE;
  Epilogue ep;
  ep.bias = nullptr; ep.residual = nullptr; ep.ldr = 0; ep.res_rows = 0; ep.alpha = 1.f; ep.act = CN_ACT_NONE;
  ep.out_f32 = out_f32; ep.out_bf16 = (__nv_bfloat16*)out_bf16; ep.ldc = ldc; ep.lo_off = 0;
  ep.ln_a = nullptr; ep.ln_b = nullptr; ep.ln_eps = 0.f; ep.ln_out = nullptr; ep.ld_ln = 0;
  ep.part = nullptr; ep.n_rowparts = 0; ep.fold_c = nullptr; ep.fold_eps = 0.f; ep.fold_d = 0;
  ep.st_max = nullptr; ep.st_idx = nullptr; ep.st_sum = nullptr; ep.st_max2 = nullptr; ep.st_idx2 = nullptr; ep.n_parts = 0;
  ep.st_target = nullptr; ep.st_target_logit = nullptr;
  ep.ce_lse = nullptr; ep.ce_target = nullptr; ep.ce_rowscale = nullptr; ep.ce_conf = 0.f; ep.ce_eps = 0.f;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int m = (int)M, n = (int)N, k = (int)K;
  if (a_mn) {
    if (bn == 256) return launch_tc<256, CN_ACT_NONE, false, false, true, true>(ma, mw, ep, m, n, k, stream);
    if (bn == 64) return launch_tc<64, CN_ACT_NONE, false, false, true, true>(ma, mw, ep, m, n, k, stream);
    return launch_tc<128, CN_ACT_NONE, false, false, true, true>(ma, mw, ep, m, n, k, stream);
  }
  if (bn == 256) return launch_tc<256, CN_ACT_NONE, false, false, false, true>(ma, mw, ep, m, n, k, stream);
  if (bn == 64) return launch_tc<64, CN_ACT_NONE, false, false, false, true>(ma, mw, ep, m, n, k, stream);
  return launch_tc<128, CN_ACT_NONE, false, false, false, true>(ma, mw, ep, m, n, k, stream);
}

// ================================================================ conv front-end (fast mode)
// ConvEmbedding's two 3x3 / stride-2 / pad-1 convolutions (src/models/modules/embedding.py:103-119) with
// channels-last bf16 activations:
//   conv1 (1 -> C channels): direct, memory-bound (its output is the large tensor);
//   conv2 (C -> C): implicit GEMM on tcgen05.  M = (utterance, 32 output frames, 4 output bins) = 128 rows,
//   K = (tap, input channel) = 9*C, N = C.  The A tile of one (tap, 64-channel) k-block is ONE 4-D TMA box over
//   the [B, T1, F1, C] input with traversal strides (1, 2, 2, 1): the stride-2 sampling and the zero padding
//   (out-of-bounds fill) are done by the TMA unit, no im2col buffer exists.
namespace {
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int CV_TT = 32, CV_FT = 4;      // output frames x output bins per 128-row tile

struct ConvArgs {
  const float* bias; __nv_bfloat16* out;  // out [B, T2, F2, C] channels-last
  int B, T2, F2, C;                       // C = input = output channels
};

template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv2_igemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_w, ConvArgs cv) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tt = (cv.T2 + CV_TT - 1) / CV_TT, ft = (cv.F2 + CV_FT - 1) / CV_FT;
  const int tiles_n = cv.C / BN, tiles_m = cv.B * tt * ft;
  const int n_tiles = tiles_m * tiles_n;
  const int cblk = cv.C / BK, k_blocks = 9 * cblk;
  auto decode = [&](int tile, int& b, int& t0, int& f0, int& n0) {
    const int mt = tile / tiles_n;
    n0 = (tile % tiles_n) * BN;
    f0 = (mt % ft) * CV_FT;
    t0 = ((mt / ft) % tt) * CV_TT;
    b = mt / (ft * tt);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int b, t0, f0, n0;
        decode(tile, b, t0, f0, n0);
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int tap = kb / cblk, c0 = (kb % cblk) * BK;
          const int dt = tap / 3, df = tap % 3;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          // input frame 2*t + dt - 1, input bin 2*f + df - 1 (pad 1): negative / past-the-end coordinates read zeros
          tma_load_4d(&map_a, &full[stage], sa, c0, 2 * f0 + df - 1, 2 * t0 + dt - 1, b);
          tma_load_2d(&map_w, &full[stage], sa + A_BYTES, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_bf16(tmem_d, da + (uint64_t)((k * UK * 2) >> 4), db + (uint64_t)((k * UK * 2) >> 4), idesc, (kb | k) ? 1u : 0u);
          tcgen05_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // epilogue: thread = one output position (TMEM lane r -> frame t0 + r/4, bin f0 + r%4); bias + ReLU, bf16,
    // 64 contiguous bytes per 32-column chunk into the channels-last output
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int b, t0, f0, n0;
      decode(tile, b, t0, f0, n0);
      const int t = t0 + (r >> 2), f = f0 + (r & 3);
      const bool ok = t < cv.T2 && f < cv.F2;
      __nv_bfloat16* orow = cv.out + (((int64_t)b * cv.T2 + t) * cv.F2 + f) * cv.C + n0;
      mbar_wait(&acc_full[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c0, v);
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x0 = fmaxf(__uint_as_float(v[j + 2 * q]) + __ldg(cv.bias + n0 + c0 + j + 2 * q), 0.f);
              const float x1 = fmaxf(__uint_as_float(v[j + 2 * q + 1]) + __ldg(cv.bias + n0 + c0 + j + 2 * q + 1), 0.f);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
              pk[q] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(orow + c0 + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// conv1: x f32 [B, Tin, Fin] (one input channel) -> y bf16 [B, T1, F1, C] = relu(conv3x3 s2 p1).
// A warp produces 256 channels of one output position per step: lane = 8 channels whose 72 weights + 8 biases
// stay in registers, the 9 input taps are warp-uniform (broadcast) loads, the store is one coalesced 512-byte
// row segment.  Two positions per iteration for memory-level parallelism.  Bound: the 2*B*T1*F1*C output bytes.
__global__ void __launch_bounds__(256) conv1_direct_kernel(const float* __restrict__ x, int64_t sbx, int64_t stx,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           int Tin, int Fin, int T1, int F1, int C, int64_t n_pos,
                                                           __nv_bfloat16* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int warp_g = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int n_warps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  const int n_rows = (int)(n_pos / F1);                    // (utterance, output frame) rows; one warp walks a row's bins
  for (int c0 = lane * 8; c0 < C; c0 += 256) {
    float wr[8][9], br[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      br[q] = bias[c0 + q];
#pragma unroll
      for (int k = 0; k < 9; ++k) wr[q][k] = w[(c0 + q) * 9 + k];
    }
    for (int row = warp_g; row < n_rows; row += n_warps) {
      const int b = row / T1, t = row - b * T1;
      const float* xb = x + (int64_t)b * sbx;
      const float* xr[3];
      bool okr[3];
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int ti = 2 * t + dt - 1;
        okr[dt] = ti >= 0 && ti < Tin;
        xr[dt] = xb + (int64_t)(okr[dt] ? ti : 0) * stx;
      }
      __nv_bfloat16* yrow = y + (int64_t)row * F1 * C + c0;
      // sliding window over the bins: the tap column 2f+1 of bin f is the column 2(f+1)-1 of bin f+1
      float prev[3];
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) prev[dt] = 0.f;       // column -1 (left padding)
      for (int f = 0; f < F1; ++f) {
        float in[9];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
          const int f1 = 2 * f, f2 = 2 * f + 1;
          in[dt * 3 + 0] = prev[dt];
          in[dt * 3 + 1] = (okr[dt] && f1 < Fin) ? __ldg(xr[dt] + f1) : 0.f;
          in[dt * 3 + 2] = (okr[dt] && f2 < Fin) ? __ldg(xr[dt] + f2) : 0.f;
          prev[dt] = in[dt * 3 + 2];
        }
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a0 = br[2 * q], a1 = br[2 * q + 1];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            a0 = fmaf(in[k], wr[2 * q][k], a0);
            a1 = fmaf(in[k], wr[2 * q + 1][k], a1);
          }
          __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
          pk[q] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(yrow + (int64_t)f * C) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

bool make_map_conv(CUtensorMap* map, const void* base, int64_t B, int64_t T1, int64_t F1, int64_t C) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)F1, (cuuint64_t)T1, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)F1 * C * 2, (cuuint64_t)T1 * F1 * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, 2 * CV_FT, 2 * CV_TT, 1};       // traversed extent; every 2nd bin / frame is taken
  cuuint32_t estr[4] = {1, 2, 2, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" int cn_conv_frontend_bf16(const float* x, int64_t sbx, int64_t stx, int64_t B, int64_t Tin, int64_t Fin,
                                     const float* w1, const float* b1, const void* w2_taps, const float* b2, int64_t C,
                                     void* mid, void* out, cn_stream_t stream_) {
  if (!x || !w1 || !b1 || !w2_taps || !b2 || !mid || !out || B < 0 || Tin <= 0 || Fin <= 0 || C <= 0) return CN_ERR_BAD_ARG;
  if (C % 256 != 0 || C > 1024 || (reinterpret_cast<uintptr_t>(mid) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(w2_taps) & 15))
    return CN_ERR_UNSUPPORTED;
  if (B == 0) return CN_OK;
  const int64_t T1 = (Tin - 1) / 2 + 1, F1 = (Fin - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  cudaStream_t stream = (cudaStream_t)stream_;
  {
    const int64_t n_pos = B * T1 * F1;
    int64_t blocks = (B * T1 + 7) / 8;                     // 8 warps per block, one (utterance, frame) row per warp step
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    conv1_direct_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, sbx, stx, w1, b1, (int)Tin, (int)Fin, (int)T1, (int)F1, (int)C,
                                                                  n_pos, (__nv_bfloat16*)mid);
    CN_CHECK_LAUNCH();
  }
  CUtensorMap ma, mw;
  if (!make_map_conv(&ma, mid, B, T1, F1, C) || !make_map(&mw, w2_taps, C, 9 * C, 9 * C, 256)) return CN_ERR_DEVICE;
  ConvArgs cv;
  cv.bias = b2; cv.out = (__nv_bfloat16*)out; cv.B = (int)B; cv.T2 = (int)T2; cv.F2 = (int)F2; cv.C = (int)C;
  constexpr uint32_t smem2 = 1024 + STAGES * (BM * BK * 2 + 256 * BK * 2) + 256;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv2_igemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess)
      return CN_ERR_LAUNCH;
    configured = true;
  }
  const int64_t tiles = B * ((T2 + CV_TT - 1) / CV_TT) * ((F2 + CV_FT - 1) / CV_FT) * (C / 256);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  conv2_igemm_kernel<256><<<grid, CONV_THREADS, smem2, stream>>>(ma, mw, cv);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// ---------------------------------------------------------------- fused head + row statistics
namespace {
__global__ void __launch_bounds__(256) merge_row_parts_kernel(const float* __restrict__ pmax, const int32_t* __restrict__ pidx,
                                                              const float* __restrict__ psum, const float* __restrict__ pmax2,
                                                              const int32_t* __restrict__ pidx2, int n_parts, int64_t rows,
                                                              int32_t* __restrict__ top1, int32_t* __restrict__ top2,
                                                              float* __restrict__ peak, float* __restrict__ lse) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t o = r * n_parts;
  float m1 = -INFINITY, m2 = -INFINITY;
  int i1 = 0x7fffffff, i2 = 0x7fffffff;
  for (int q = 0; q < n_parts; ++q) {                     // parts are in increasing column order
    const float v = pmax[o + q];
    const int i = pidx[o + q];
    if (v > m1) { m2 = m1; i2 = i1; m1 = v; i1 = i; }
    else if (v > m2) { m2 = v; i2 = i; }
    if (pidx2) {
      const float w = pmax2[o + q];
      const int k = pidx2[o + q];
      if (cn_better(w, k, m2, i2) && k != i1) { m2 = w; i2 = k; }
    }
  }
  float s = 0.f;
  for (int q = 0; q < n_parts; ++q) s += psum[o + q] * expf(pmax[o + q] - m1);
  top1[r] = i1;
  if (top2) top2[r] = i2;
  if (peak) peak[r] = m1;
  if (lse) lse[r] = m1 + logf(s);
}
}  // namespace

extern "C" int64_t cn_gemm_rowstats_workspace_bytes(int64_t M, int64_t N) {
  const int64_t parts = ((N + 255) / 256) * 2;
  return M * parts * 5 * 4 + 256;
}

static int gemm_rowstats_impl(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             float alpha, int64_t M, int64_t N, int64_t K, int32_t* top1, int32_t* top2,
                             float* peak, float* lse, const int32_t* target, float* target_logit,
                             void* workspace, int64_t workspace_bytes, int split, cn_stream_t stream_) {
  const int64_t kcols = split ? 2 * K : K;
  if (!A || !W || !top1 || !workspace || M < 0 || N <= 0 || K <= 0 || lda < kcols || ldw < kcols) return CN_ERR_BAD_ARG;
  if ((target == nullptr) != (target_logit == nullptr)) return CN_ERR_BAD_ARG;
  if (split && (K % BK)) return CN_ERR_UNSUPPORTED;
  if ((lda & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return CN_ERR_UNSUPPORTED;
  if (M > 0x7fffff00 || N > 0x7fffff00 || K > 0x7fffff00) return CN_ERR_UNSUPPORTED;
  if (workspace_bytes < cn_gemm_rowstats_workspace_bytes(M, N)) return CN_ERR_WORKSPACE;
  if (M == 0) return CN_OK;
  CUtensorMap ma, mw;
  if (!make_map(&ma, A, M, kcols, lda, BM) || !make_map(&mw, W, N, kcols, ldw, 256)) return CN_ERR_DEVICE;
  const int parts = (int)(((N + 255) / 256) * 2);
  float* pmax = (float*)workspace;
  int32_t* pidx = (int32_t*)(pmax + M * parts);
  float* psum = (float*)(pidx + M * parts);
  float* pmax2 = psum + M * parts;
  int32_t* pidx2 = (int32_t*)(pmax2 + M * parts);
  Epilogue ep;
  ep.bias = bias; ep.residual = nullptr; ep.ldr = 0; ep.res_rows = 0; ep.alpha = alpha; ep.act = CN_ACT_NONE;
  ep.out_f32 = nullptr; ep.out_bf16 = nullptr; ep.ldc = 0; ep.lo_off = 0;
  ep.ln_a = nullptr; ep.ln_b = nullptr; ep.ln_eps = 0.f; ep.ln_out = nullptr; ep.ld_ln = 0;
  ep.part = nullptr; ep.n_rowparts = 0; ep.fold_c = nullptr; ep.fold_eps = 0.f; ep.fold_d = 0;
  ep.st_max = pmax; ep.st_idx = pidx; ep.st_sum = psum; ep.st_max2 = top2 ? pmax2 : nullptr;
  ep.st_idx2 = top2 ? pidx2 : nullptr; ep.n_parts = parts;
  ep.st_target = target; ep.st_target_logit = target_logit;
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = launch_tc<256, CN_ACT_NONE, false, true>(ma, mw, ep, (int)M, (int)N, (int)K, stream, split);
  if (rc != CN_OK) return rc;
  merge_row_parts_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(pmax, pidx, psum, ep.st_max2, ep.st_idx2, parts, M,
                                                                          top1, top2, peak, lse);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_gemm_bf16_rowstats(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                     float alpha, int64_t M, int64_t N, int64_t K, int32_t* top1, int32_t* top2,
                                     float* peak, float* lse, const int32_t* target, float* target_logit,
                                     void* workspace, int64_t workspace_bytes, cn_stream_t stream_) {
  return gemm_rowstats_impl(A, lda, W, ldw, bias, alpha, M, N, K, top1, top2, peak, lse, target, target_logit, workspace,
                            workspace_bytes, 0, stream_);
}

extern "C" int cn_gemm_bf16x3_rowstats(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                       float alpha, int64_t M, int64_t N, int64_t K, int32_t* top1, int32_t* top2,
                                       float* peak, float* lse, const int32_t* target, float* target_logit,
                                       void* workspace, int64_t workspace_bytes, cn_stream_t stream_) {
  return gemm_rowstats_impl(A, lda, W, ldw, bias, alpha, M, N, K, top1, top2, peak, lse, target, target_logit, workspace,
                            workspace_bytes, 1, stream_);
}

// f32 -> bf16 cast (operands of the tensor-core GEMM that arrive in fp32, e.g. enc_h)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = cn_ld_stream_f32x4(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[i] = pk;
  }
}

// f32 [rows, K] -> split bf16 operand [rows, 2K] = [hi | lo], hi = bf16(x), lo = bf16(x - hi): hi + lo carries 16
// mantissa bits, which is what the bf16x3 GEMM consumes.  The hi half alone is the plain bf16 cast.
__global__ void __launch_bounds__(256) cast_bf16_split_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                                              int64_t ldy, int64_t rows, int K4) {
  const int64_t total = rows * K4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K4;
    const int c = (int)(i - r * K4) * 4;
    const float4 v = cn_ld_stream_f32x4(reinterpret_cast<const float4*>(x + r * ldx + c));
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(y + r * ldy + c) = pk;
    pk.x = *reinterpret_cast<uint32_t*>(&l0); pk.y = *reinterpret_cast<uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(y + r * ldy + 4 * (int64_t)K4 + c) = pk;
  }
}

extern "C" int cn_cast_bf16_split(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int64_t K,
                                  cn_stream_t stream_) {
  if (!x || !y || rows < 0 || K <= 0 || (K & 3) || ldx < K || ldy < 2 * K || (ldx & 3) || (ldy & 3) ||
      (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 7))
    return CN_ERR_BAD_ARG;
  if (rows == 0) return CN_OK;
  int64_t blocks = (rows * (K / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_bf16_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, ldx, (__nv_bfloat16*)y, ldy, rows, (int)(K / 4));
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_cast_bf16(const float* x, void* y, int64_t n, cn_stream_t stream_) {
  if (!x || !y || n < 0 || (n & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 7))
    return CN_ERR_BAD_ARG;
  if (n == 0) return CN_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, (__nv_bfloat16*)y, n / 4);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_set_sm_limit(int n_ctas) {
  const int before = g_sm_limit;
  g_sm_limit = n_ctas > 0 ? n_ctas : 0;
  return before;
}
