// Masked multi-head attention on the 5th-gen tensor cores (fast mode), d_k = 64, Tk <= 512.
//
// Persistent kernel, one CTA per SM looping over (utterance, head) items; per item the K/V tiles
// are loaded once (3-D TMA maps over [B, T, row], 128-byte swizzle, per-tile mbarriers) and shared
// by all query tiles.  Two softmax warpgroups ping-pong on two query tiles (A/B) so that the
// tensor pipe computes S_B while group A is in its softmax and vice versa:
//   * S_g = Q_g K_j^T  (128 x 128 x 64, K-major operands) into TMEM;
//   * the softmax group reads its S_g row ONCE (tcgen05.ld, 128 fp32 values per thread stay in registers), applies
//     1/sqrt(d_k) and the mask with the reference's finfo.min semantics (fully masked rows stay
//     uniform over all Tk keys, modules/attention.py:19-21), and writes P_g as packed bf16 back into TMEM
//     (tcgen05.st): P never touches shared memory;
//   * O_g += P_g V_j (128 x 64 x 128; A operand from TMEM, the V tile used in place as an MN-major shared-memory
//     operand) accumulates in TMEM across the key tiles.  The running maximum is only advanced when a tile exceeds
//     it by more than 2^8 (lazy rescale): then, and only then, a thread rescales its O row in TMEM; the exact
//     normalisation by the row sum happens once, at the end;
//   * a dedicated TMA warp runs ahead through the schedule (next Q tiles, next item's K/V as soon
//     as the last MMA that reads a slot has retired).
// Replaces (fast mode) attention() of src/models/modules/attention.py:13-24.
#include <cstdlib>
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int QT = 128;        // queries per CTA
constexpr int KT = 128;        // keys per tile
constexpr int DK = 64;
constexpr int MAX_TILES = 4;   // K/V ring slots (all of Tk <= 512 stays resident)
constexpr int MAX_TK = 2048;   // key-mask words kept in shared memory

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
// wait with a suspend-time hint: the single-lane producer / MMA warps share their schedulers with softmax
// warps, so they should sleep in hardware rather than spin through issue slots
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    if (!ok && ++spins > 100000000u) __trap();
  }
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
// 32 lanes x 32 consecutive 32-bit columns <- 32 registers per thread (thread i = TMEM lane base+i)
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
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (P, bf16 pairs packed along K in 32-bit columns) comes from TMEM
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
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

struct RelTable { const void* data; int bf16; int64_t ld, sb, hs; };      // see AttnTcArgs::reltab

struct AttnTcArgs {
  const uint8_t* mask; int64_t mb;          // key mask u8 [B,Tk] (row pitch mb), shared by all query rows
  const uint32_t* mask_bits;                // or: per-row mask, bit-packed [B,Tq,16 words] (bit t of word t/32 = keep)
  int q_bcast;                              // Q is [1,Tq,d], shared by every utterance (extractor queries)
  int B, h, Tq, Tk;
  __nv_bfloat16* out; int64_t ldo, sbo;
  // REL: table of (q_i + v_h) . P[r,h] for the 2m+1 clamped offsets, f32 or bf16, element (b, i, head, r) at
  // reltab[b*rel_sb + i*rel_ld + head*rel_hs + r]; rel_slab = floats of shared memory per softmax group for it
  const void* reltab; int rel_bf16; int64_t rel_ld, rel_sb, rel_hs; int max_rel; int rel_slab;
  // optional per-utterance hints (key masks only): keys >= kv_len[b] are all masked (their tiles are
  // skipped: exactly zero weight), query rows >= q_len[b] are not needed by the caller (zero-filled)
  const int32_t* kv_len; const int32_t* q_len;
  // optional processing order: utterance order[i] is the i-th to be scheduled (heaviest first), so that the
  // static round-robin over the persistent CTAs hands every CTA a similar amount of work
  const int32_t* order;
  int causal;                               // hint: key j > query i is masked for every row -> tiles above the diagonal are skipped
};

// Persistent, ping-pong design (v2):
//   warps 0-3  softmax group A (query tile 2p of the current (utterance, head) item)
//   warps 4-7  softmax group B (query tile 2p+1)          -> two softmax warps per scheduler
//   warp 8     tcgen05.mma issuer (one lane)
//   warp 9     TMA producer (one lane), runs ahead through the whole schedule
// smem: Q_A | Q_B | K 4x16K | V 4x16K | barriers ; TMEM: S_A S_B | O_A O_B | P_A P_B
constexpr uint32_t OFF_Q = 0, OFF_K = 2 * 16384, OFF_V = OFF_K + MAX_TILES * 16384,
                   OFF_BAR = OFF_V + MAX_TILES * 16384, SMEM_BYTES = OFF_BAR + 1024 + 1024;   // barriers + key-mask words, alignment slack
constexpr int NTHREADS = 320;
// barrier indices
constexpr int B_QFULL = 0, B_QFREE = 2, B_KFULL = 4, B_VFULL = 8, B_KFREE = 12, B_VFREE = 16, B_SFULL = 20,
              B_PREADY = 22, B_PFREE = 24, N_BARS = 26;
constexpr float RESCALE_LOG2 = 8.f;     // the running maximum is advanced only when a tile exceeds it by 2^8

template <bool REL, bool BITS, bool RING>
__global__ void __launch_bounds__(NTHREADS, 1) attention_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                   const __grid_constant__ CUtensorMap map_k,
                                                                   const __grid_constant__ CUtensorMap map_v,
                                                                   AttnTcArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 256);
  uint32_t* kbits_all = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 288);   // [2 groups][MAX_TK / 32 words]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_max = (p.Tk + KT - 1) / KT;        // key tiles (<= 4)
  const int nq_max = (p.Tq + QT - 1) / QT;             // query tiles
  // per-item tile counts; every role evaluates the same expressions, so they stay in lock step
  auto tiles_of = [&](int b, int pr, int& n_tiles, int& nq) {
    n_tiles = n_tiles_max;
    if (p.causal) n_tiles = min(n_tiles, 2 * pr + 2);   // rows of pair pr end at query 256*pr+255: keys beyond are masked
    nq = nq_max;
    if (p.kv_len) { int t = (min(max(p.kv_len[b], 1), p.Tk) + KT - 1) / KT; n_tiles = min(max(t, 1), n_tiles); }
    if (p.q_len) { int t = (min(max(p.q_len[b], 0), p.Tq) + QT - 1) / QT; nq = min(t, nq_max); }
  };
  // work item = (utterance, head, pair of 128-query tiles): finer than (utterance, head) so that the tail of
  // the static round-robin over 148 CTAs stays short (K/V of an item are re-fetched per pair, from L2)
  const int np_max = (nq_max + 1) / 2;
  const int n_items = p.B * p.h * np_max;
  auto decode_item = [&](int item, int& b, int& hh, int& pr) {
    const int u = item / (p.h * np_max), rem = item % (p.h * np_max);
    b = p.order ? p.order[u] : u;
    hh = rem / np_max;
    pr = rem % np_max;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_BARS; ++i) mbar_init(bar(i), (i == B_PREADY || i == B_PREADY + 1) ? 4 : 1);   // one arrival per softmax warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  cn_pdl_wait();                 // (prologue above overlapped the predecessor's tail)
  cn_pdl_launch();
  // TMEM columns: S_A 0..127 | S_B 128..255 | O_A 256..319 | O_B 320..383 | P_A 384..447 | P_B 448..511 (bf16 pairs)
  auto tm_s = [&](int g) { return tmem + g * 128; };
  auto tm_o = [&](int g) { return tmem + 256 + g * 64; };
  auto tm_p = [&](int g) { return tmem + 384 + g * 64; };

  if (warp == 9) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // K/V tiles go through a ring of MAX_TILES slots each (tile j -> slot j % MAX_TILES); up to MAX_TILES
      // tiles of the next item are prefetched before its Q tiles, later ones follow as slots are released
      uint32_t k_fill = 0, v_fill = 0, q_cnt[2] = {0, 0};      // phase bit per slot, kept in one register each
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, hh, pr0;
        decode_item(item, b, hh, pr0);
        int n_tiles, nq;
        tiles_of(b, pr0, n_tiles, nq);
        if (2 * pr0 >= nq) continue;                            // nothing to compute for this item
        auto load_k = [&](int j) {
          const int sl = RING ? j % MAX_TILES : j;          // RING: Tk > 512 streams through the slots
          mbar_wait_sleepy(bar(B_KFREE + sl), ((k_fill >> sl) & 1) ^ 1);
          mbar_expect_tx(bar(B_KFULL + sl), KT * DK * 2);
          tma_load_3d(&map_k, bar(B_KFULL + sl), sbase + OFF_K + sl * 16384, hh * DK, j * KT, b);
          k_fill ^= 1u << sl;
        };
        auto load_v = [&](int j) {
          const int sl = RING ? j % MAX_TILES : j;          // RING: Tk > 512 streams through the slots
          mbar_wait_sleepy(bar(B_VFREE + sl), ((v_fill >> sl) & 1) ^ 1);
          mbar_expect_tx(bar(B_VFULL + sl), KT * DK * 2);
          tma_load_3d(&map_v, bar(B_VFULL + sl), sbase + OFF_V + sl * 16384, hh * DK, j * KT, b);
          v_fill ^= 1u << sl;
        };
        const int n_first = n_tiles < MAX_TILES ? n_tiles : MAX_TILES;
        for (int j = 0; j < n_first; ++j) load_k(j);
        for (int j = 0; j < n_first; ++j) load_v(j);
        {
          const int pr = pr0;
          for (int g = 0; g < 2; ++g) {
            const int qt = 2 * pr + g;
            if (qt >= nq) continue;
            mbar_wait_sleepy(bar(B_QFREE + g), (q_cnt[g] & 1) ^ 1);
            mbar_expect_tx(bar(B_QFULL + g), QT * DK * 2);
            tma_load_3d(&map_q, bar(B_QFULL + g), sbase + OFF_Q + g * 16384, hh * DK, qt * QT, p.q_bcast ? 0 : b);
            ++q_cnt[g];
          }
        }
        if (RING)
          for (int j = n_first; j < n_tiles; ++j) { load_k(j); load_v(j); }   // Tk > 512: stream through the ring
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(QT, KT, false);
      constexpr uint32_t idesc_o = idesc_bf16(QT, DK, true);
      uint32_t k_seen = 0, v_seen = 0, q_cnt[2] = {0, 0}, p_cnt[2] = {0, 0};   // phase bit per K/V slot
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b_, hh_, pr;
        decode_item(item, b_, hh_, pr);
        int n_tiles, nq;
        tiles_of(b_, pr, n_tiles, nq);
        if (2 * pr >= nq) continue;
        {
          const bool has[2] = {2 * pr < nq, 2 * pr + 1 < nq};
          const bool last_pair = true;                          // one pair per item: K/V are released at its end
          const int last_g = has[1] ? 1 : 0;
          auto issue_s = [&](int g, int j, bool last_user) {     // last_user: the other group already issued S(j)
            const int sl = RING ? j % MAX_TILES : j;          // RING: Tk > 512 streams through the slots
            if (j == 0) mbar_wait_sleepy(bar(B_QFULL + g), q_cnt[g] & 1);
            mbar_wait_sleepy(bar(B_KFULL + sl), (k_seen >> sl) & 1);
            fence_after();
            const uint64_t dq = desc_kmajor(sbase + OFF_Q + g * 16384), dkk = desc_kmajor(sbase + OFF_K + sl * 16384);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k) umma(tm_s(g), dq + (uint64_t)(k * 2), dkk + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
            tcgen05_commit(bar(B_SFULL + g));
            if (j == n_tiles - 1) { tcgen05_commit(bar(B_QFREE + g)); ++q_cnt[g]; }   // Q tile no longer needed
            if (last_pair && last_user) { tcgen05_commit(bar(B_KFREE + sl)); k_seen ^= 1u << sl; }   // K_j consumed
          };
          for (int g = 0; g < 2; ++g)
            if (has[g]) issue_s(g, 0, g == last_g);
          for (int j = 0; j < n_tiles; ++j) {
            for (int g = 0; g < 2; ++g) {
              if (!has[g]) continue;
              mbar_wait_sleepy(bar(B_PREADY + g), p_cnt[g] & 1);       // P_g(j) written, S_g(j) consumed
              ++p_cnt[g];
              // S_g(j+1) goes to the tensor pipe BEFORE P_g(j) V_j: the group's next softmax is the critical
              // path, O_g(j) is only folded after that softmax
              const bool last_user = g == last_g;
              if (j + 1 < n_tiles) issue_s(g, j + 1, last_user);
              const int sl = RING ? j % MAX_TILES : j;          // RING: Tk > 512 streams through the slots
              mbar_wait_sleepy(bar(B_VFULL + sl), (v_seen >> sl) & 1);
              fence_after();
              const uint64_t dv = desc_mnmajor(sbase + OFF_V + sl * 16384);
              // O_g (+)= P_g V_j: P from TMEM (8 packed columns per 16 keys), accumulating over the key tiles
#pragma unroll
              for (int k = 0; k < KT / 16; ++k)
                umma_ts(tm_o(g), tm_p(g) + (uint32_t)(k * 8), dv + (uint64_t)((k * 16 * 128) >> 4), idesc_o, (j | k) ? 1u : 0u);
              tcgen05_commit(bar(B_PFREE + g));                 // P_g consumed, O_g updated (the softmax group waits on it)
              if (last_pair && last_user) { tcgen05_commit(bar(B_VFREE + sl)); v_seen ^= 1u << sl; }
            }
          }
        }
      }
    }
  } else {
    // ===================== softmax groups: thread = query row = TMEM lane =====================
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const float scale = 0.125f * 1.4426950408889634f;      // 1/sqrt(64) * log2(e)
    uint32_t* kbits = kbits_all + g * (MAX_TK / 32);
    uint32_t s_cnt = 0, pf_cnt = 0;                         // phases of S_full / P_free (one completion per key tile)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, hh, pr0;
      decode_item(item, b, hh, pr0);
      int n_tiles, nq;
      tiles_of(b, pr0, n_tiles, nq);
      if (2 * pr0 >= nq && !(pr0 == 0 && nq < nq_max)) continue;
      if (p.mask && 2 * pr0 < nq) {
        // key-mask bits of this utterance (one private copy per group)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");     // previous item's readers are done
        const uint8_t* mk = p.mask + (int64_t)b * p.mb;
        for (int k0 = 0; k0 < p.Tk; k0 += 128) {
          const int k = k0 + row;
          const uint32_t bits = __ballot_sync(0xffffffffu, k < p.Tk && mk[k] != 0);
          if (lane == 0) kbits[k >> 5] = bits;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      }
      if (pr0 == 0 && nq < nq_max) {
        // query rows the caller does not need: zero-filled (finite inputs for whatever reads them next)
        for (int r = nq * QT + g * 128 + row; r < p.Tq; r += 256) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (int64_t)b * p.sbo + (int64_t)r * p.ldo + hh * DK);
#pragma unroll
          for (int d = 0; d < DK / 8; ++d) dst[d] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      const int qt = 2 * pr0 + g;
      if (qt >= nq) continue;
      const int qi = qt * QT + row;
      const uint32_t* brow = BITS ? p.mask_bits + ((int64_t)b * p.Tq + (qi < p.Tq ? qi : 0)) * 16 : nullptr;
      // BITS: a row without any unmasked key is uniform over all Tk keys (scores all finfo.min,
      // modules/attention.py:19-21) = softmax of equal scores; every other row gives masked keys exactly zero
      // weight, so chunks without a kept key cost nothing
      bool uniform_row = false;
      if (BITS) {
        uint32_t any = 0;
        for (int w = 0; w < (p.Tk + 31) >> 5; ++w) {
          const int n = p.Tk - w * 32;
          any |= __ldg(brow + w) & (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
        }
        uniform_row = any == 0u;
      }
      // m_run: the maximum the exponentials are taken against (log2 domain).  It lags the true running maximum by
      // at most RESCALE_LOG2, so P stays below 2^8 and O / l_run need no per-tile correction.
      float m_run = -INFINITY, l_run = 0.f;
      // relative-position term (modules/attention.py:104-143): score += (q_i+v).P[clamp(j-i,-m,m)+m]
      // Each thread copies its own row's 2m+1 entries into its group's slab of shared memory, SKEWED: entry r of
      // local row il goes to [(il + r) * nrel + r], i.e. to [(j - j_first) * nrel + r] for the key j = i - m + r it
      // applies to -- the lanes of a warp (consecutive query rows) gathering the band entries of ONE key then read
      // consecutive words.  The loads fly while the first score tile is being computed.
      const int nrel = 2 * p.max_rel + 1;
      const int qic = qi < p.Tq ? qi : p.Tq - 1;
      float* slab = reinterpret_cast<float*>(smem + OFF_BAR + 1024) + g * p.rel_slab;
      float rel_left = 0.f, rel_right = 0.f;
      if (REL) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");       // the previous item's gathers are done
        const int64_t src = (int64_t)b * p.rel_sb + (int64_t)qic * p.rel_ld + (int64_t)hh * p.rel_hs;
        float* dst = slab + row * nrel;
        if (p.rel_bf16) {
          const __nv_bfloat16* t16 = reinterpret_cast<const __nv_bfloat16*>(p.reltab) + src;
#pragma unroll 4
          for (int r = 0; r < nrel; ++r) {
            const float v = __bfloat162float(t16[r]);
            dst[r * (nrel + 1)] = v;
            if (r == 0) rel_left = v;
            rel_right = v;
          }
        } else {
          const float* t32 = reinterpret_cast<const float*>(p.reltab) + src;
#pragma unroll 4
          for (int r = 0; r < nrel; ++r) {
            const float v = __ldg(t32 + r);
            dst[r * (nrel + 1)] = v;
            if (r == 0) rel_left = v;
            rel_right = v;
          }
        }
      }
      const int i_lo = min(qt * QT + (warp & 3) * 32, p.Tq - 1), i_hi = min(qt * QT + (warp & 3) * 32 + 31, p.Tq - 1);   // this warp's rows
      bool any_keep = false;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(bar(B_SFULL + g), s_cnt & 1);
        ++s_cnt;
        fence_after();
        const int k0 = j * KT;
        const uint32_t s_addr = tm_s(g) + lane_addr;
        if (REL && j == 0) asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");     // every row of the slab is written
        if (REL) {
          // Band pre-pass: the 32-key chunks in which some row of this warp is within +-max_rel of a key get the
          // gathered term added IN TENSOR MEMORY (load, add, store back) -- a chunk's TMEM column is a run-time
          // address, so ONE copy of the gather code serves all chunks (unrolled into the register-resident score
          // row below it quadrupled the loop body and the kernel became instruction-fetch bound).
#pragma unroll 1
          for (int c = 0; c < KT; c += 32) {
            if (k0 + c + 31 - i_lo <= -p.max_rel || k0 + c - i_hi >= p.max_rel || k0 + c >= p.Tk) continue;
            float sb[32];
            tmem_ld32(s_addr + c, sb);
            const int r0 = k0 + c - qic + p.max_rel;                        // table column of the chunk's first key
            const int lr0 = k0 + c + p.max_rel - qt * QT;                   // slab row of the chunk's first key
            // key q of the chunk uses table column r0 + q of slab row lr0 + q: one stride of nrel + 1 words per key;
            // columns outside [0, nrel) are the clamped ends (this row's first / last entry, already in registers)
            const float* band = slab + lr0 * nrel + r0;
            float bv[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const bool inside = (unsigned)(r0 + q) < (unsigned)nrel;
              bv[q] = inside ? band[q * (nrel + 1)] : (r0 + q < 0 ? rel_left : rel_right);
            }
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 32; ++q) sb[q] += bv[q];
            tmem_st32(s_addr + c, reinterpret_cast<uint32_t*>(sb));
            tmem_wait_st();
          }
        }
        float t[KT];                                               // this thread's score row of the tile
#pragma unroll
        for (int c = 0; c < KT; c += 32) tmem_ld32(s_addr + c, t + c);
        tmem_wait_ld();
        // ---- masks / relative term per 32-key chunk; afterwards either t holds raw scores (state RAW: the exponent is
        // fma(t, scale, off*scale - m)) or, for chunks that needed per-element work, final log2-domain values (DONE)
        float off[KT / 32];
        uint32_t done = 0;                                         // bit c: chunk c holds final log2-domain values
        float mx = -INFINITY;
#pragma unroll
        for (int ci = 0; ci < KT / 32; ++ci) {
          float* tc = t + ci * 32;
          const int c = ci * 32;
          const int n = p.Tk - (k0 + c);
          const uint32_t valid = n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
          uint32_t keep;
          if (BITS) keep = uniform_row ? 0xffffffffu : __ldg(brow + ((k0 + c) >> 5));   // per-row mask, bit-packed
          else if (p.mask) keep = kbits[(k0 + c) >> 5];                 // key mask: shared by all rows
          else keep = 0xffffffffu;
          off[ci] = 0.f;
          if (REL) {
            // away from the diagonal every key of the chunk clamps to the same table entry (ONE offset, folded into
            // the exponent for free); chunks that straddle the +-max_rel band got their terms in the pre-pass above
            if (k0 + c + 31 - i_lo <= -p.max_rel) off[ci] = rel_left;
            else if (k0 + c - i_hi >= p.max_rel) off[ci] = rel_right;
            any_keep = any_keep || ((keep & valid) != 0u);
          }
          if (BITS && uniform_row) {
#pragma unroll
            for (int q = 0; q < 32; ++q) tc[q] = 0.f;                      // equal scores
          }
          if ((keep & valid) == 0xffffffffu) {
            float cm[4] = {tc[0], tc[1], tc[2], tc[3]};
#pragma unroll
            for (int q = 4; q < 32; ++q) cm[q & 3] = fmaxf(cm[q & 3], tc[q]);
            mx = fmaxf(mx, (fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) + off[ci]) * scale);   // scale > 0
          } else if (BITS && (keep & valid) == 0u) {
            done |= 1u << ci;                                             // no weight in this chunk
#pragma unroll
            for (int q = 0; q < 32; ++q) tc[q] = -INFINITY;
          } else {
            done |= 1u << ci;
            float mm = -INFINITY;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              float v = (tc[q] + off[ci]) * scale;
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
        const float m_use = m_run == -INFINITY ? 0.f : m_run;      // a row with no weight so far: exponentials are 0
        // ---- exponentials, row sum, P as packed bf16 pairs (two keys per 32-bit TMEM column)
        float lsv[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[KT / 2];
#pragma unroll
        for (int ci = 0; ci < KT / 32; ++ci) {
          float* tc = t + ci * 32;
          if (!((done >> ci) & 1u)) {                                // raw scores: scale, offset and maximum in one FFMA
            const float addend = fmaf(off[ci], scale, -m_use);
#pragma unroll
            for (int q = 0; q < 32; ++q) tc[q] = fmaf(tc[q], scale, addend);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) tc[q] -= m_use;
          }
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float e0 = ex2(tc[q]), e1 = ex2(tc[q + 1]);
            lsv[(q >> 1) & 3] += e0 + e1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
            pk[(ci * 32 + q) >> 1] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
        // ---- P_g / O_g are free once PV(j-1) has retired
        if (j > 0) {
          mbar_wait(bar(B_PFREE + g), pf_cnt & 1);
          ++pf_cnt;
          fence_after();
          if (__any_sync(0xffffffffu, rescale)) {                  // rare: this warp's rows advance their maximum
            float o[DK];
            tmem_ld32(tm_o(g) + lane_addr, o);
            tmem_ld32(tm_o(g) + lane_addr + 32, o + 32);
            tmem_wait_ld();
#pragma unroll
            for (int d = 0; d < DK; ++d) o[d] *= corr;
            tmem_st32(tm_o(g) + lane_addr, reinterpret_cast<uint32_t*>(o));
            tmem_st32(tm_o(g) + lane_addr + 32, reinterpret_cast<uint32_t*>(o) + 32);
          }
        }
        tmem_st32(tm_p(g) + lane_addr, pk);
        tmem_st32(tm_p(g) + lane_addr + 32, pk + 32);
        tmem_wait_st();
        fence_before();                                                  // S_g read, P_g (and O_g) written: hand over
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PREADY + g));
        l_run = l_run * corr + ((lsv[0] + lsv[1]) + (lsv[2] + lsv[3]));
      }
      // ---- epilogue of the query tile: O_g / l once the last PV has retired
      mbar_wait(bar(B_PFREE + g), pf_cnt & 1);
      ++pf_cnt;
      fence_after();
      {
        float o[DK];
        tmem_ld32(tm_o(g) + lane_addr, o);
        tmem_ld32(tm_o(g) + lane_addr + 32, o + 32);
        tmem_wait_ld();
        fence_before();
        const float inv = (REL && !any_keep) ? 0.f : 1.f / l_run;      // REL: masked probabilities are re-zeroed
        if (qi < p.Tq) {
          __nv_bfloat16* dst = p.out + (int64_t)b * p.sbo + (int64_t)qi * p.ldo + hh * DK;
#pragma unroll
          for (int d = 0; d < DK; d += 8) {
            __nv_bfloat162 b0 = __floats2bfloat162_rn(o[d] * inv, o[d + 1] * inv), b1 = __floats2bfloat162_rn(o[d + 2] * inv, o[d + 3] * inv),
                           b2 = __floats2bfloat162_rn(o[d + 4] * inv, o[d + 5] * inv), b3 = __floats2bfloat162_rn(o[d + 6] * inv, o[d + 7] * inv);
            uint4 w4;
            w4.x = *reinterpret_cast<uint32_t*>(&b0); w4.y = *reinterpret_cast<uint32_t*>(&b1);
            w4.z = *reinterpret_cast<uint32_t*>(&b2); w4.w = *reinterpret_cast<uint32_t*>(&b3);
            *reinterpret_cast<uint4*>(dst + d) = w4;
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) {
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

// (q_i + v_h) . P[r, h, :] for every query row and the 2m+1 clamped offsets; Q holds q + u_h.
// Thread = one (row, head): its 64 query values stay in registers, the head's P rows are shared-memory
// broadcasts (all lanes of a warp read the same address).
__global__ void __launch_bounds__(128) rel_bias_kernel(const __nv_bfloat16* __restrict__ Q, int64_t ldq, int64_t sbq,
                                                       const float* __restrict__ rel_pos, const float* __restrict__ rel_u,
                                                       const float* __restrict__ rel_v, int B, int h, int Tq, int nrel,
                                                       float* __restrict__ out) {
  extern __shared__ float ps[];                      // [nrel][64] rows of this head, then (v - u)[64]
  const int hh = blockIdx.y;
  for (int i = threadIdx.x; i < nrel * 64; i += blockDim.x) ps[i] = rel_pos[(int64_t)(i >> 6) * h * 64 + hh * 64 + (i & 63)];
  float* du = ps + nrel * 64;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) du[i] = rel_v[hh * 64 + i] - rel_u[hh * 64 + i];
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= (int64_t)B * Tq) return;
  const int64_t b = row / Tq;
  const int i = (int)(row % Tq);
  const uint4* src = reinterpret_cast<const uint4*>(Q + b * sbq + (int64_t)i * ldq + hh * 64);
  float q[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 raw = src[c];
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h2[e]);
      q[c * 8 + 2 * e] = f.x + du[c * 8 + 2 * e];
      q[c * 8 + 2 * e + 1] = f.y + du[c * 8 + 2 * e + 1];
    }
  }
  float* o = out + (((int64_t)b * h + hh) * Tq + i) * nrel;
  for (int r = 0; r < nrel; ++r) {
    const float4* pr = reinterpret_cast<const float4*>(ps + r * 64);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 pv = pr[c];
      a0 = fmaf(q[4 * c], pv.x, a0); a1 = fmaf(q[4 * c + 1], pv.y, a1);
      a2 = fmaf(q[4 * c + 2], pv.z, a2); a3 = fmaf(q[4 * c + 3], pv.w, a3);
    }
    o[r] = (a0 + a1) + (a2 + a3);
  }
}

}  // namespace

static int attention_bf16_impl(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                               const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb, int64_t mi,
                               const uint32_t* mask_bits, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                               void* out, int64_t ldo, int64_t sbo, const RelTable* rel, int max_rel,
                               const int32_t* kv_len, const int32_t* q_len, const int32_t* order, int causal,
                               cn_stream_t stream_) {
  if (!Q || !K || !V || !out || B < 0 || h <= 0 || Tq < 0 || Tk <= 0) return CN_ERR_BAD_ARG;
  if (dk != DK || Tk > MAX_TK || h > 65535 || B > 65535) return CN_ERR_UNSUPPORTED;
  if (mask_bits && Tk > 512) return CN_ERR_UNSUPPORTED;  // bit-packed per-row masks: 16 words per row
  if (mask && mi != 0) return CN_ERR_UNSUPPORTED;      // per-row byte masks go through cn_pack_mask_bits + *_bits
  const bool q_bcast = sbq == 0;
  // TMA: 16-byte aligned bases and pitches; 16-byte vector stores of the output rows
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (sbq & 7) || (sbk & 7) || (sbv & 7) || (ldo & 7) || (sbo & 7) ||
      ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V) |
        reinterpret_cast<uintptr_t>(out)) & 15))
    return CN_ERR_UNSUPPORTED;
  if (B == 0 || Tq == 0) return CN_OK;
  CUtensorMap mq, mk, mv;
  const int64_t width = h * DK;
  if (!make_map3(&mq, Q, q_bcast ? 1 : B, Tq, width, ldq, q_bcast ? Tq * ldq : sbq) || !make_map3(&mk, K, B, Tk, width, ldk, sbk) ||
      !make_map3(&mv, V, B, Tk, width, ldv, sbv))
    return CN_ERR_DEVICE;
  AttnTcArgs p;
  p.mask = mask_bits ? nullptr : mask; p.mb = mb; p.mask_bits = mask_bits; p.q_bcast = q_bcast ? 1 : 0;
  p.B = (int)B; p.h = (int)h; p.Tq = (int)Tq; p.Tk = (int)Tk;
  p.out = (__nv_bfloat16*)out; p.ldo = ldo; p.sbo = sbo; p.max_rel = max_rel;
  p.reltab = nullptr; p.rel_bf16 = 0; p.rel_ld = p.rel_sb = p.rel_hs = 0; p.rel_slab = 0;
  size_t smem_bytes = SMEM_BYTES;
  const bool relbias = rel != nullptr;
  if (relbias) {
    const int64_t slab = (int64_t)(QT + 2 * max_rel) * (2 * max_rel + 1);      // floats per softmax group
    if (SMEM_BYTES + 2 * slab * 4 > 227 * 1024) return CN_ERR_UNSUPPORTED;      // max_rel <= 22
    p.reltab = rel->data; p.rel_bf16 = rel->bf16; p.rel_ld = rel->ld; p.rel_sb = rel->sb; p.rel_hs = rel->hs;
    p.rel_slab = (int)slab;
    smem_bytes += 2 * slab * 4;
  }
  p.kv_len = (mask || mask_bits) ? kv_len : nullptr; p.q_len = q_len; p.order = order; p.causal = causal;
  static bool configured = false;
  if (!configured) {
    const int sm = 227 * 1024;
    const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
    if (cudaFuncSetAttribute(attention_tc_kernel<false, false, false>, at, sm) != cudaSuccess ||
        cudaFuncSetAttribute(attention_tc_kernel<false, false, true>, at, sm) != cudaSuccess ||
        cudaFuncSetAttribute(attention_tc_kernel<true, false, false>, at, sm) != cudaSuccess ||
        cudaFuncSetAttribute(attention_tc_kernel<true, false, true>, at, sm) != cudaSuccess ||
        cudaFuncSetAttribute(attention_tc_kernel<false, true, false>, at, sm) != cudaSuccess)
      return CN_ERR_LAUNCH;
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = B * h * ((Tq + 2 * QT - 1) / (2 * QT));          // (utterance, head, query-tile pair)
  const unsigned grid = (unsigned)(items < sms ? items : sms);           // persistent: one CTA per SM
  if (relbias && mask_bits) return CN_ERR_UNSUPPORTED;
  const bool ring = Tk > MAX_TILES * KT;
  cudaStream_t st = (cudaStream_t)stream_;
  if (relbias && ring) cn_launch_pdl(attention_tc_kernel<true, false, true>, dim3(grid), dim3(NTHREADS), smem_bytes, st, mq, mk, mv, p);
  else if (relbias) cn_launch_pdl(attention_tc_kernel<true, false, false>, dim3(grid), dim3(NTHREADS), smem_bytes, st, mq, mk, mv, p);
  else if (mask_bits) cn_launch_pdl(attention_tc_kernel<false, true, false>, dim3(grid), dim3(NTHREADS), SMEM_BYTES, st, mq, mk, mv, p);
  else if (ring) cn_launch_pdl(attention_tc_kernel<false, false, true>, dim3(grid), dim3(NTHREADS), SMEM_BYTES, st, mq, mk, mv, p);
  else cn_launch_pdl(attention_tc_kernel<false, false, false>, dim3(grid), dim3(NTHREADS), SMEM_BYTES, st, mq, mk, mv, p);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

extern "C" int cn_attention_bf16(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                                 const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb, int64_t mi,
                                 int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk, void* out, int64_t ldo,
                                 int64_t sbo, const int32_t* kv_len, const int32_t* q_len, const int32_t* order,
                                 cn_stream_t stream_) {
  return attention_bf16_impl(Q, ldq, sbq, K, ldk, sbk, V, ldv, sbv, mask, mb, mi, nullptr, B, h, Tq, Tk, dk, out, ldo, sbo,
                             nullptr, 0, kv_len, q_len, order, 0, stream_);
}

extern "C" int cn_attention_bf16_bits(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                                      const void* V, int64_t ldv, int64_t sbv, const uint32_t* mask_bits, int64_t B,
                                      int64_t h, int64_t Tq, int64_t Tk, int64_t dk, void* out, int64_t ldo, int64_t sbo,
                                      const int32_t* kv_len, const int32_t* q_len, int causal_hint, cn_stream_t stream_) {
  if (!mask_bits) return CN_ERR_BAD_ARG;
  return attention_bf16_impl(Q, ldq, sbq, K, ldk, sbk, V, ldv, sbv, nullptr, 0, 0, mask_bits, B, h, Tq, Tk, dk, out, ldo,
                             sbo, nullptr, 0, kv_len, q_len, nullptr, causal_hint ? 1 : 0, stream_);
}

extern "C" int64_t cn_rel_attention_workspace_bytes(int64_t B, int64_t h, int64_t Tq, int64_t max_rel) {
  return B * h * Tq * (2 * max_rel + 1) * 4 + 256;
}

extern "C" int cn_rel_attention_bf16(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                                     const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb,
                                     int64_t mi, const float* rel_pos, const float* rel_bias_u, const float* rel_bias_v,
                                     int64_t max_rel, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                                     void* out, int64_t ldo, int64_t sbo, void* workspace, int64_t workspace_bytes,
                                     const int32_t* kv_len, cn_stream_t stream_) {
  if (!rel_pos || !rel_bias_u || !rel_bias_v || !workspace || max_rel < 0 || !Q) return CN_ERR_BAD_ARG;
  if (h > 8 || max_rel > 22 || dk != DK) return CN_ERR_UNSUPPORTED;
  if (workspace_bytes < cn_rel_attention_workspace_bytes(B, h, Tq, max_rel)) return CN_ERR_WORKSPACE;
  if (B == 0 || Tq == 0) return CN_OK;
  float* relbias = (float*)workspace;
  const int nrel = (int)(2 * max_rel + 1);
  rel_bias_kernel<<<dim3((unsigned)((B * Tq + 127) / 128), (unsigned)h), 128, (size_t)(nrel + 1) * 64 * sizeof(float),
                    (cudaStream_t)stream_>>>((const __nv_bfloat16*)Q, ldq, sbq, rel_pos, rel_bias_u, rel_bias_v, (int)B, (int)h,
                                             (int)Tq, nrel, relbias);
  CN_CHECK_LAUNCH();
  const RelTable rel = {relbias, 0, nrel, h * Tq * nrel, Tq * nrel};            // [B,h,Tq,2m+1] f32
  return attention_bf16_impl(Q, ldq, sbq, K, ldk, sbk, V, ldv, sbv, mask, mb, mi, nullptr, B, h, Tq, Tk, dk, out, ldo, sbo,
                             &rel, (int)max_rel, kv_len, nullptr, nullptr, 0, stream_);
}

extern "C" int cn_rel_attention_bf16_table(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                                           const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb,
                                           int64_t mi, const void* table, int table_is_bf16, int64_t table_ld,
                                           int64_t table_sb, int64_t table_hs, int64_t max_rel, int64_t B, int64_t h,
                                           int64_t Tq, int64_t Tk, int64_t dk, void* out, int64_t ldo, int64_t sbo,
                                           const int32_t* kv_len, cn_stream_t stream_) {
  if (!table || max_rel < 0 || !Q) return CN_ERR_BAD_ARG;
  if (h > 8 || max_rel > 22 || dk != DK) return CN_ERR_UNSUPPORTED;
  if (B == 0 || Tq == 0) return CN_OK;
  const RelTable rel = {table, table_is_bf16 ? 1 : 0, table_ld, table_sb, table_hs};
  return attention_bf16_impl(Q, ldq, sbq, K, ldk, sbk, V, ldv, sbv, mask, mb, mi, nullptr, B, h, Tq, Tk, dk, out, ldo, sbo,
                             &rel, (int)max_rel, kv_len, nullptr, nullptr, 0, stream_);
}
