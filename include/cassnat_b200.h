/* cassnat_b200 — C ABI of the B200-native CASS-NAT alignment-conditioned decoding path.
 *
 * The reference (balaji1312/cassnat_asr_public) has no FFI layer: its boundary for this
 * path is the Python object protocol of `CassNAT` (src/models/cassnat.py) and the FaNAT
 * blocks (src/models/blocks/fanat_blocks.py, fanat_conformer_blocks.py).  This header is
 * what a ctypes binding of that path binds instead; every entry point cites the reference
 * code (file:line, relative to the reference root) whose torch-op sequence it replaces.
 * The Python mirror in cassnat_asr_public_b200/ is the only in-tree caller.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (tensor.data_ptr()) unless it says "host";
 *   - sizes/strides are in ELEMENTS, int64_t; innermost dimension contiguous unless a
 *     stride is given for it;
 *   - the caller allocates outputs and workspaces (no cudaMalloc inside);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never
 *     synchronises the host;
 *   - return value: CN_OK (0) or a negative cn_status; nothing is thrown;
 *   - sm_100a only.  There is no CPU fallback: without a Blackwell GPU every compute call
 *     returns CN_ERR_DEVICE / CN_ERR_LAUNCH.
 *   - masks are uint8 (0 = masked), integer tensors the reference returns as int64 stay int64.
 */
#ifndef CASSNAT_B200_H_
#define CASSNAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cn_stream_t;

enum cn_status {
  CN_OK = 0,
  CN_ERR_BAD_ARG = -1,      /* null pointer, negative size, inconsistent shapes            */
  CN_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels implement (says which in doc) */
  CN_ERR_LAUNCH = -3,       /* cudaGetLastError() after a launch                            */
  CN_ERR_WORKSPACE = -4,    /* workspace too small                                          */
  CN_ERR_DEVICE = -5        /* no sm_100 device / driver entry point missing                */
};

enum cn_act { CN_ACT_NONE = 0, CN_ACT_RELU = 1, CN_ACT_SWISH = 2 };

int cn_version(void);
const char* cn_status_string(int status);
/* 0 when the current device is compute capability 10.x, CN_ERR_DEVICE otherwise. */
int cn_device_check(void);

/* ------------------------------------------------------------------------------------
 * K1/K2/K9 row statistics over a [rows, V] f32 matrix (V contiguous, row_stride elements
 * between rows): first-index argmax, runner-up index, max value, log-sum-exp and an
 * optional log-softmax write.
 *   replaces: ctc_out.argmax(-1) / .max(-1) / .topk(2,-1)   src/models/cassnat.py:371-378
 *             F.log_softmax in Generator.forward             src/models/cassnat.py:110-113
 *             att_out.argmax / topk(beam_width=1)            src/models/cassnat.py:501,611
 * Ties: lower index first (torch.argmax order; torch.topk's order among equal values is
 * implementation-defined).  top2/peak/lse/logp_out may be NULL.
 * logp_out[r, v] = x[r, v] - lse[r]  (out_stride elements between rows; may alias x).
 */
int cn_row_stats(const float* x, int64_t rows, int64_t V, int64_t row_stride,
                 int32_t* top1, int32_t* top2, float* peak, float* lse,
                 float* logp_out, int64_t out_stride, cn_stream_t stream);

/* K2  greedy / sampled CTC best path -> aligned_seq_shift, ylen, ymax.
 *   replaces: CassNAT.best_path_align                        src/models/cassnat.py:367-389
 * top1/top2/peak [B*T] come from cn_row_stats on the [B,T,V] log-posteriors.
 * src_mask u8 [B,T].  select u8 [B,T] = the host-drawn torch.randint(0,2,(B,T,1)) of
 * cassnat.py:373 (NULL when sample_num <= 1; the draw stays on the host so the CPU default
 * generator is consumed exactly like the reference).  Outputs: aligned i64 [B,T],
 * ylen i64 [B], ymax i32 [1] (device; max over b of ylen). */
int cn_bestpath_collapse(const int32_t* top1, const int32_t* top2, const float* peak,
                         const uint8_t* select, const uint8_t* src_mask,
                         int64_t B, int64_t T, int64_t blank, int64_t sample_num,
                         float threshold, int include_best,
                         int64_t* aligned, int64_t* ylen, int32_t* ymax, cn_stream_t stream);

/* K2 convenience: cn_row_stats + cn_bestpath_collapse on log-posteriors logp [B,T,V]
 * (strides sb, st in elements).  workspace: cn_bestpath_workspace_bytes(B,T). */
int64_t cn_bestpath_workspace_bytes(int64_t B, int64_t T);
int cn_bestpath_from_logp(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                          const uint8_t* src_mask, const uint8_t* select, int64_t blank,
                          int64_t sample_num, float threshold, int include_best,
                          int64_t* aligned, int64_t* ylen, int32_t* ymax,
                          void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* K3  CTC forced alignment (max-plus DP over the (2U+1) x T lattice + back-trace +
 * collapse + shift), one CTA per utterance: one DP warp, three emission-gather warps,
 * lattice staged through shared memory, 2-bit back-pointers in the workspace.
 *   replaces: CassNAT.viterbi_align                          src/models/cassnat.py:272-353
 * logp f32 [B,T,V] (strides sb, st); src_mask u8 [B,T]; src_size i64 [B] (frames);
 * ys i64 [B,U] (row stride ys_stride); ylens i64 [B]; t_sample i32 [n_sample] = frames
 * where the 2nd-best predecessor is taken (cassnat.py:311-326; host-drawn), may be NULL.
 * Output aligned i64 [B,T].  Supported: 1 <= U <= 511 (2U+1 <= 1024 lattice states).
 * Tie-breaks as the reference: stay > s-1 > s-2; end state needs strict '>'. */
int64_t cn_viterbi_workspace_bytes(int64_t B, int64_t T, int64_t U);
int cn_viterbi_align(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                     const uint8_t* src_mask, const int64_t* src_size,
                     const int64_t* ys, int64_t U, int64_t ys_stride, const int64_t* ylens,
                     int64_t blank, const int32_t* t_sample, int64_t n_sample,
                     int64_t* aligned, void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* K3 on compact emissions.  The forced alignment reads only U+1 of the V posteriors of a frame; with the dumped
 * [B,T,V] tensor that gather costs a 64-byte DRAM atom per 4-byte value (cfg2: 3.7 GB moved for 0.2 GB used).
 * cn_gather_emissions extracts emis[b,t,0] = logp[b,t,blank], emis[b,t,u] = logp[b,t,ys[b,u-1]] (u = 1..U) once --
 * fully parallel, and shareable with the CTC loss, which needs the same values -- and cn_viterbi_align_compact runs
 * the identical dynamic program (same kernel, bit-identical alignments) on 4*(U+1) coalesced bytes per frame.
 *   replaces: the emission indexing of CassNAT.viterbi_align      src/models/cassnat.py:296-331
 * emis f32 [B,T,U+1] (batch / frame pitches eb, et); the other arguments as cn_viterbi_align (V: vocabulary size,
 * used to clamp label ids exactly as there). */
int cn_gather_emissions(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* ys,
                        int64_t U, int64_t ys_stride, int64_t blank, float* emis, cn_stream_t stream);
int cn_viterbi_align_compact(const float* emis, int64_t B, int64_t T, int64_t V, int64_t eb, int64_t et,
                             const uint8_t* src_mask, const int64_t* src_size, const int64_t* ys, int64_t U,
                             int64_t ys_stride, const int64_t* ylens, int64_t blank, const int32_t* t_sample,
                             int64_t n_sample, int64_t* aligned, void* workspace, int64_t workspace_bytes,
                             cn_stream_t stream);

/* K4/K5  alignment -> per-token trigger mask (+ optional +-1 frame dilation and src_mask AND).
 *   replaces: CassNAT.align_to_mask                          src/models/cassnat.py:355-365
 *             CassNAT.expand_trigger_mask + "& src_mask"     src/models/cassnat.py:259-270,:175,:468
 * aligned i64 [B,T]; ylens i64 [B]; src_mask u8 [B,T]; src_size i64 [B].
 * mask_out u8 [B, ymax+1, T]; ylen_out i64 [B] = ylens + 1 (may be NULL).
 * left/right: 0 or 1 (the reference only tests > 0).  and_src_mask != 0 applies the final
 * "& src_mask" of cassnat.py:468.  With left=right=0 and and_src_mask=0 the result is
 * exactly align_to_mask's.  seg_out i32 [B,T] (may be NULL) receives cumsum(aligned != blank),
 * the compact form the fused extractor consumes. */
int cn_align_to_mask(const int64_t* aligned, const int64_t* ylens, const uint8_t* src_mask,
                     const int64_t* src_size, int64_t B, int64_t T, int64_t ymax, int64_t blank,
                     int left, int right, int and_src_mask,
                     uint8_t* mask_out, int64_t* ylen_out, int32_t* seg_out, cn_stream_t stream);

/*   replaces: CassNAT.expand_trigger_mask on an arbitrary mask   src/models/cassnat.py:259-270
 * mask u8 [B,U,T] -> out u8 [B,U,T] (out must not alias mask). */
int cn_expand_trigger_mask(const uint8_t* mask, int64_t B, int64_t U, int64_t T,
                           int left, int right, uint8_t* out, cn_stream_t stream);

/*   replaces: tgt_mask scatter/cumprod/scatter                 src/models/cassnat.py:185-187,476-478
 * out_i64 i64 [B,1,ymax] and/or out_u8 u8 [B,ymax] (either may be NULL): 1 where u < ylen[b]. */
int cn_tgt_length_mask(const int64_t* ylen, int64_t B, int64_t ymax,
                       int64_t* out_i64, uint8_t* out_u8, cn_stream_t stream);

/* ------------------------------------------------------------------------------------
 * fp32 building blocks of the extractor / SAD / MAD / heads (parity mode: plain FFMA,
 * one rounding per multiply-add, no tensor cores).
 */

/* C[M,N] = residual + alpha * act(A[M,K] * W[N,K]^T + bias)   (bias, residual may be NULL);
 * residual row for output row m is (res_rows > 0 ? m % res_rows : m), which lets the
 * batch-independent sinusoid rows of cassnat.py:481 act as the residual without a repeat.
 *   replaces: nn.Linear / PositionwiseFeedForward / SublayerConnection residual
 *             src/models/modules/positionff.py:6-16, modules/utils.py:25-32,
 *             modules/attention.py:55-66 (projections), cassnat.py:105,113 (Generator.proj) */
int cn_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act,
                float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, cn_stream_t stream);

/* y = a * (x - mean) / (std_unbiased + eps) + b over the last dim (d <= 2048)
 *   replaces: LayerNorm.forward                               src/models/modules/norm.py:15-18
 * out_f32 and/or out_bf16 (uint16 storage) may be NULL. */
int cn_layernorm(const float* x, int64_t rows, int64_t d, int64_t ldx, const float* a, const float* b,
                 float eps, float* out_f32, void* out_bf16, int64_t ldo, cn_stream_t stream);

/* Masked multi-head attention, exact reference semantics: scores/sqrt(dk), masked entries
 * set to finfo(f32).min (fully masked rows -> uniform over ALL Tk keys), softmax, P*V.
 *   replaces: attention() + head split/merge                  src/models/modules/attention.py:13-24,55-65
 * Q [B,Tq,h*dk] row stride ldq; K,V [B,Tk,h*dk] row strides ldk, ldv; batch strides sbq/sbk/sbv.
 * mask u8, element (b,i,j) at mask[b*mb + i*mi + j] (mi = 0 broadcasts a [B,1,Tk] mask;
 * mask may be NULL).  rel_* (all or none NULL) add Transformer-XL relative-position scores:
 *   replaces: RelMultiHeadedAttention.forward                 src/models/modules/attention.py:104-143
 *   score(i,j) = ((q_i+u_h).k_j + (q_i+v_h).P[clamp(j-i,-m,m)+m, h]) / sqrt(dk),
 *   rel_pos f32 [2m+1, h*dk] = linear_pos(embedding rows), and masked probabilities are
 *   re-zeroed after the softmax (attention.py:138) so fully masked rows give 0.
 * out [B,Tq,h*dk] row stride ldo.  dk must be 64. */
int cn_attention_f32(const float* Q, int64_t ldq, int64_t sbq, const float* K, int64_t ldk, int64_t sbk,
                     const float* V, int64_t ldv, int64_t sbv,
                     const uint8_t* mask, int64_t mb, int64_t mi,
                     const float* rel_pos, const float* rel_bias_u, const float* rel_bias_v, int64_t max_rel,
                     int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                     float* out, int64_t ldo, int64_t sbo, cn_stream_t stream);

/* K6  token acoustic extractor attention with the trigger mask evaluated on the fly from
 * the alignment (never materialised): each token row u pools only the frames of its CTC
 * segment (+-1 dilation), rows whose mask is empty reproduce the reference's uniform
 * attention over all T frames.
 *   replaces: align_to_mask + expand_trigger_mask + MultiHeadedAttention on the trigger mask
 *             src/models/cassnat.py:465-468,482; modules/attention.py:13-24
 * Q f32 [U, h*dk] (batch-independent: projected sinusoid rows) row stride ldq;
 * K,V [B,T,h*dk]; seg i32 [B,T] from cn_align_to_mask; ylens i64 [B] (token counts BEFORE
 * the +1 of align_to_mask); src_mask u8 [B,T]; src_size i64 [B]; out [B,U,h*dk]. dk == 64. */
int cn_segment_attention_f32(const float* Q, int64_t ldq, const float* K, int64_t ldk, int64_t sbk,
                             const float* V, int64_t ldv, int64_t sbv,
                             const int32_t* seg, const int64_t* ylens, const uint8_t* src_mask,
                             const int64_t* src_size, int left, int right,
                             int64_t B, int64_t h, int64_t U, int64_t T, int64_t dk,
                             float* out, int64_t ldo, int64_t sbo,
                             void* workspace, int64_t workspace_bytes, cn_stream_t stream);
int64_t cn_segment_attention_workspace_bytes(int64_t B, int64_t U, int64_t d /* = h*dk */);

/* Conformer convolution-module core between the two pointwise convolutions (which are GEMMs):
 * GLU over channels -> depthwise conv (odd kernel <= 63, zero padding) along the token axis ->
 * GroupNorm(1,C) over the whole padded [C,L] slab of each utterance -> activation.
 *   replaces: F.glu / depthwise_conv / GroupNorm / Swish      src/models/modules/conformer_related.py:33-38
 * y f32 or bf16 [B,L,2C] (row stride ldy, batch stride sby) = pointwise_conv1 output; dw_weight [C,ksz];
 * out f32 [B,L,C] contiguous; out_bf16 (optional) receives the result as bf16 -- with bf16 y and kernel sizes
 * 3/7 the convolution is then evaluated twice from y (statistics, then normalise-and-store) instead of being
 * written to `out`, whose contents are unspecified in that case.  workspace: cn_conv_module_workspace_bytes(B,L). */
int64_t cn_conv_module_workspace_bytes(int64_t B, int64_t L);
int cn_glu_dwconv_groupnorm_act(const void* y, int y_is_bf16, int64_t ldy, int64_t sby, const float* dw_weight,
                                const float* dw_bias, const float* gamma, const float* beta, float eps, int act,
                                int64_t B, int64_t L, int64_t C, int64_t ksz, float* out /* f32 scratch + result */,
                                void* out_bf16 /* optional bf16 result (then `out` only holds the pre-norm values) */,
                                void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Fast mode: bf16 tensor-core GEMM (tcgen05.mma, fp32 accumulation in TMEM, TMA-fed, persistent).
 * C[M,N] = residual + alpha * act(A[M,K] * W[N,K]^T + bias), A and W bf16 (uint16 storage),
 * K-major; lda, ldw multiples of 8 and 16-byte aligned bases (TMA).  out_f32 and/or out_bf16.
 *   replaces (fast mode): the same call sites as cn_gemm_f32. */
int cn_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                 const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act,
                 float* out_f32, void* out_bf16, int64_t ldc, int64_t M, int64_t N, int64_t K,
                 cn_stream_t stream);
/* y(bf16) = x(f32), n elements (n % 4 == 0, 16-byte aligned x). */
int cn_cast_bf16(const float* x, void* y, int64_t n, cn_stream_t stream);

/* The pre-norm residual step and the LayerNorm that consumes it, in one kernel:
 *   x_new = residual + alpha * (A W^T + bias)      -> out_f32 [M,N]   (the fp32 residual stream)
 *   ln_out = ln_a * (x_new - mean) / (std_unbiased + ln_eps) + ln_b   -> bf16 [M,N] (operand of the next GEMM)
 *   replaces: SublayerConnection + the next sublayer's LayerNorm   src/models/modules/utils.py:25-32, norm.py:15-18
 * A CTA owns complete 128-row blocks (its <= 2 column tiles back to back); the row statistics are accumulated in
 * the epilogue, exchanged through shared memory, and the normalised copy is produced from the L2-hot fp32 rows
 * while the tensor pipe already works on the next block.  N <= 512, N % 4 == 0, 16-byte aligned rows. */
int cn_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                    const float* residual, int64_t ldr, int64_t res_rows, float alpha, float* out_f32, int64_t ldc,
                    const float* ln_a, const float* ln_b, float ln_eps, void* ln_out, int64_t ld_ln,
                    int64_t M, int64_t N, int64_t K, cn_stream_t stream);

/* LayerNorm folding (fast mode): the pre-norm LayerNorm between a residual GEMM and the GEMM that consumes it
 * disappears as a kernel.  LN(x) W^T = rinv * (x (W diag(a))^T - mean * c) + (bias + W b), c_n = sum_k (W diag(a))_nk:
 *   producer  cn_gemm_bf16_rowpart: the residual GEMM also writes the raw bf16 copy of its fp32 rows and per-row
 *             partial sums (sum, sum of squares over each 128-column half): part f32 [M][n_parts][2],
 *             n_parts = cn_gemm_rowpart_count(N, K);
 *   consumer  cn_gemm_bf16_folded: A = the raw bf16 rows, W_folded = bf16(W diag(a)), colsum = c (of the ROUNDED
 *             folded weight, so that a constant row cancels exactly), bias_folded = bias + W b; its epilogue
 *             finalises mean / 1/(std_unbiased + eps) of its rows from the partial sums and applies the formula.
 *   replaces: LayerNorm.forward between SublayerConnection and the next nn.Linear
 *             src/models/modules/norm.py:15-18, modules/utils.py:25-32 */
int64_t cn_gemm_rowpart_count(int64_t N, int64_t K);
int cn_gemm_bf16_rowpart(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                         const float* residual, int64_t ldr, int64_t res_rows, float alpha, float* out_f32,
                         void* out_bf16, int64_t ldc, float* part, int64_t M, int64_t N, int64_t K, cn_stream_t stream);
int cn_gemm_bf16_folded(const void* A_raw, int64_t lda, const void* W_folded, int64_t ldw, const float* bias_folded,
                        const float* colsum, const float* part, int64_t n_parts, int64_t ln_width, float ln_eps,
                        float alpha, int act, float* out_f32, void* out_bf16, int64_t ldc, int64_t M, int64_t N,
                        int64_t K, cn_stream_t stream);

/* ------------------------------------------------------------------------------------
 * bf16x3 ("split") arithmetic: fp32-grade products on the bf16 tensor cores.  A value v is carried as the pair
 * hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits); a split operand matrix stores [hi | lo] along the contraction
 * axis ([rows, 2K], row pitch ld).  The GEMM accumulates A_hi W_hi + A_lo W_hi + A_hi W_lo in fp32 (TMEM): three
 * passes of the same tcgen05 mainloop with remapped TMA coordinates; relative error ~2^-17 of |a||w| instead of
 * 2^-9.  Used for the two decision heads (K1 CTC argmax -> alignment, K9 token pick: cassnat.py:367-389, :611),
 * where a bf16 rounding flip would change integers, and for the whole path in the 'bf16x3' precision mode.
 * K % 64 == 0.  a_split == 0: A is a plain bf16 [M,K] operand (an attention context), only W is split (two passes).
 * out_split != 0: out_bf16 is itself written as a split operand [M, 2N] (ldc >= 2N, out_f32 NULL). */
int cn_gemm_bf16x3(const void* A, int64_t lda, int a_split, const void* W_split, int64_t ldw, const float* bias,
                   const float* residual, int64_t ldr, int64_t res_rows, float alpha, int act,
                   float* out_f32, void* out_bf16, int out_split, int64_t ldc, int64_t M, int64_t N, int64_t K,
                   cn_stream_t stream);
/* x f32 [rows, K] (row pitch ldx) -> split operand bf16 [rows, 2K] (row pitch ldy); K % 4 == 0.  The hi half is
 * the plain bf16 cast (single-pass GEMMs may read it in place with row pitch ldy). */
int cn_cast_bf16_split(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int64_t K, cn_stream_t stream);
/* LayerNorm (as cn_layernorm) emitting the split operand bf16 [rows, 2d] (row pitch ldo) and, optionally, the fp32
 * result as well (out_f32 may be NULL; row pitch ldo_f32): the decoder's final norm feeding the K9 head. */
int cn_layernorm_bf16x2(const float* x, int64_t rows, int64_t d, int64_t ldx, const float* a, const float* b,
                        float eps, float* out_f32, int64_t ldo_f32, void* out_split, int64_t ldo, cn_stream_t stream);

/* Fast mode: masked multi-head attention on tcgen05 (bf16 Q/K/V in, bf16 context out, fp32 softmax).
 * Same semantics as cn_attention_f32 without the relative-position terms; d_k == 64, Tk <= 512,
 * row/batch pitches multiples of 8 elements and 16-byte aligned bases (TMA).
 *   replaces (fast mode): attention()                          src/models/modules/attention.py:13-24 */
int cn_attention_bf16(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                      const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb, int64_t mi,
                      int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                      void* out, int64_t ldo, int64_t sbo,
                      const int32_t* kv_len /* i32 [B] or NULL: keys >= kv_len[b] are all masked (key masks only) */,
                      const int32_t* q_len  /* i32 [B] or NULL: rows >= q_len[b] are not needed -> zero-filled */,
                      const int32_t* order  /* i32 [B] or NULL: a permutation of the utterances, heaviest first,
                                               so that the persistent CTAs get balanced work */,
                      cn_stream_t stream);

/* K10  CTC loss (alpha-beta in log space) fused with the log-softmax backward.
 *   replaces: torch.nn.CTCLoss(reduction='mean', zero_infinity=True)   src/models/cassnat.py:126,232
 * logp f32 [B,T,V] log-posteriors (strides sb, st; the reference passes the [T,B,V] transposed view
 * of the same storage); targets i64 [B,L] padded (row stride tgt_stride); input_lengths,
 * target_lengths i64 [B]; blank id.  Outputs: loss f32 [1] = mean_b(nll_b / max(L_b,1)) with inf -> 0;
 * nll f32 [B]; grad_logits f32 [B,T,V] (strides gsb, gst; may be NULL) = grad_out * d loss / d logits
 * where logp = log_softmax(logits) (what autograd delivers after log_softmax.backward).
 * Supported: L <= 511.  workspace: cn_ctc_loss_workspace_bytes(B,T,L). */
int64_t cn_ctc_loss_workspace_bytes(int64_t B, int64_t T, int64_t L);
int cn_ctc_loss(const float* logp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                const int64_t* targets, int64_t L, int64_t tgt_stride, const int64_t* input_lengths,
                const int64_t* target_lengths, int64_t blank, float grad_out, float* loss, float* nll,
                float* grad_logits, int64_t gsb, int64_t gst, void* workspace, int64_t workspace_bytes,
                cn_stream_t stream);

/* K1/K9 fused: head GEMM (bf16 tcgen05, fp32 accumulate) whose epilogue keeps only the row
 * statistics of  z = alpha * (A W^T + bias):  first-index argmax, runner-up, max and log-sum-exp.
 * The [M,N] logits never reach HBM.
 *   replaces: Generator.forward + argmax/max/topk(2)       src/models/cassnat.py:110-113,371-378,501,611
 * top2 / peak / lse may be NULL.  workspace: cn_gemm_rowstats_workspace_bytes(M,N). */
int64_t cn_gemm_rowstats_workspace_bytes(int64_t M, int64_t N);
int cn_gemm_bf16_rowstats(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float alpha,
                          int64_t M, int64_t N, int64_t K, int32_t* top1, int32_t* top2, float* peak, float* lse,
                          const int32_t* target /* i32 [M] or NULL: one column per row (ids outside [0,N) are ignored) */,
                          float* target_logit   /* f32 [M]: z[row, target[row]]; log-prob = target_logit - lse
                                                   (rescoring, src/models/cassnat.py:517: gather(lm_out, att_pred)) */,
                          void* workspace, int64_t workspace_bytes, cn_stream_t stream);
/* Fused linear + label-smoothed cross entropy, backward half: recomputes z = A W^T + bias tile
 * by tile and writes only dz = rowscale[m] * (exp(z - lse[m]) - w), w = conf at column target[m] and eps elsewhere,
 * in bf16 (the operand of the dX / dW GEMMs).  The forward half is cn_gemm_bf16_rowstats (lse, target logit, argmax;
 * no logits) plus the row sums of the logits = A . colsum(W) + sum(bias).  No fp32 [rows, V] tensor exists in either
 * direction.
 *   replaces: Generator + LabelSmoothing + their autograd      src/models/cassnat.py:226,248, src/utils/loss.py:21-35 */
int cn_gemm_bf16_ce_grad(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* lse,
                         const int32_t* target, const float* rowscale, float conf, float eps, void* dz_bf16, int64_t ldc,
                         int64_t M, int64_t N, int64_t K, cn_stream_t stream);

/* Attention in split-bf16 arithmetic (precise mode): Q/K/V as (hi, lo) bf16 pairs with common pitches, scores and
 * context accumulated from three tcgen05 passes each (relative error ~2^-17).  Context out: fp32 `out`, or (out NULL)
 * as (hi, lo) bf16 pairs at out16 / out16 + lo_off -- the operand of the three-pass output projection -- with the same
 * pitches ldo / sbo (elements).  Key masks and the kv_len / q_len / order hints as cn_attention_bf16; d_k == 64,
 * Tk <= 2048.
 *   replaces (precise mode): attention()                       src/models/modules/attention.py:13-24 */
int cn_attention_bf16x3(const void* Qhi, const void* Qlo, int64_t ldq, int64_t sbq, const void* Khi, const void* Klo,
                        int64_t ldk, int64_t sbk, const void* Vhi, const void* Vlo, int64_t ldv, int64_t sbv,
                        const uint8_t* mask, int64_t mb, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                        float* out, void* out16, int64_t lo_off, int64_t ldo, int64_t sbo, const int32_t* kv_len,
                        const int32_t* q_len, const int32_t* order, cn_stream_t stream);

/* The same head with split (bf16x3) operands A [M,2K], W [N,2K]: argmax / runner-up equal the fp32 result wherever
 * the fp32 top-2 margin exceeds ~1e-4 of the logit scale (decision-safe heads of the fast mode). */
int cn_gemm_bf16x3_rowstats(const void* A_split, int64_t lda, const void* W_split, int64_t ldw, const float* bias,
                            float alpha, int64_t M, int64_t N, int64_t K, int32_t* top1, int32_t* top2, float* peak,
                            float* lse, const int32_t* target, float* target_logit, void* workspace,
                            int64_t workspace_bytes, cn_stream_t stream);

/* Fast mode, encoder front-end (upstream of the accelerated path): the two 3x3 / stride-2 / pad-1
 * convolutions + ReLU of ConvEmbedding with channels-last bf16 activations.
 *   replaces: ConvEmbedding.conv                              src/models/modules/embedding.py:103-104,114-116
 * x f32 [B,Tin,Fin] (batch / frame pitches sbx, stx); w1 f32 [C,9], b1 [C]; w2_taps bf16 [C, 9*C] with
 * k = (3*dt+df)*C + c_in; b2 f32 [C]; mid bf16 [B,T1,F1,C] (scratch), out bf16 [B,T2,F2,C],
 * T1 = (Tin-1)/2+1, F1 = (Fin-1)/2+1, T2 = (T1-1)/2+1, F2 = (F1-1)/2+1.  conv2 is an implicit GEMM on tcgen05
 * whose A tiles are 4-D TMA boxes with traversal stride 2 (no im2col buffer).  C % 256 == 0. */
int cn_conv_frontend_bf16(const float* x, int64_t sbx, int64_t stx, int64_t B, int64_t Tin, int64_t Fin,
                          const float* w1, const float* b1, const void* w2_taps, const float* b2, int64_t C,
                          void* mid, void* out, cn_stream_t stream);

/* Length hints of one decode step for the tcgen05 attention in one launch (B <= 1024):
 *   need_rows[b] = min(ylen[b]+1, ymax)   token rows the pick loop can read (cassnat.py:587)
 *   kv_tgt[b]    = min(ylen[b], ymax)     valid tokens (keys of the self-attention)
 *   last_key[b]  = 1 + last frame with src_mask != 0 (keys of the source attention)
 *   order_tgt / order_src: utterances sorted by (key tiles) x (query tiles), heaviest first (stable);
 *   use_q = 0 leaves the query tiles out (conformer family: every row is needed). */
int cn_decode_hints(const int64_t* ylen, int64_t ymax, const uint8_t* src_mask, int64_t B, int64_t T, int use_q,
                    int32_t* need_rows, int32_t* kv_tgt, int32_t* last_key, int32_t* order_tgt, int32_t* order_src,
                    cn_stream_t stream);

/* Fast mode, per-row masks: the mask arrives bit-packed, bits[b][i][w] (16 uint32 words per query row,
 * bit (t&31) of word (t>>5) = keep(i,t)); cn_attention_bf16 itself only takes key masks (mi == 0).
 *   cn_pack_mask_bits      u8 [B,Tq,Tk] (strides mb, mi) -> bits        (dense masks: src_trigger, unimask)
 *   cn_trigger_mask_bits   segment ids + lengths -> bits of the trigger mask of cassnat.py:355-365,259-270,:468
 *                          (the [B,U+1,T] byte mask is never materialised; eos fix-up, +-1 dilation, & src_mask)
 *   cn_attention_bf16_bits the tcgen05 attention with that mask; sbq == 0 shares one [Tq, h*64] query block
 *                          between all utterances (the extractor's positional queries).
 *   replaces (fast mode): AcEmbedExtractor's masked cross-attention  fanat_blocks.py:24-37,101-113 */
int cn_pack_mask_bits(const uint8_t* mask, int64_t mb, int64_t mi, int64_t B, int64_t Tq, int64_t Tk, uint32_t* bits,
                      cn_stream_t stream);
int cn_trigger_mask_bits(const int32_t* seg, const int64_t* ylens, const uint8_t* src_mask, const int64_t* src_size,
                         int64_t B, int64_t T, int64_t n_rows, int left, int right, uint32_t* bits, cn_stream_t stream);
int cn_attention_bf16_bits(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                           const void* V, int64_t ldv, int64_t sbv, const uint32_t* mask_bits, int64_t B, int64_t h,
                           int64_t Tq, int64_t Tk, int64_t dk, void* out, int64_t ldo, int64_t sbo,
                           const int32_t* kv_len, const int32_t* q_len,
                           int causal_hint /* 1: the mask is zero for every key j > query i (LM ranker): key tiles above
                                              the diagonal are not computed; the bits are still applied inside */,
                           cn_stream_t stream);

/* Fast mode: relative-position (Transformer-XL style) self-attention on tcgen05.
 *   replaces (fast mode): RelMultiHeadedAttention.forward      src/models/modules/attention.py:104-143
 * Q must hold q + pos_bias_u (fold u into the query bias of the projection); rel_pos f32
 * [2m+1, h*64] = linear_pos(embedding rows); rel_bias_u / rel_bias_v f32 [h,64].  A pre-kernel
 * tabulates (q_i + v_h).P[r,h] into the workspace; the attention kernel adds
 * table[clamp(j-i,-m,m)+m] to the scores and re-zeroes fully masked rows (attention.py:138).
 * h <= 8, max_rel <= 22 (the table rows of a query tile are staged in shared memory).
 * workspace: cn_rel_attention_workspace_bytes(B,h,Tq,max_rel). */
int64_t cn_rel_attention_workspace_bytes(int64_t B, int64_t h, int64_t Tq, int64_t max_rel);
int cn_rel_attention_bf16(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                          const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb, int64_t mi,
                          const float* rel_pos, const float* rel_bias_u, const float* rel_bias_v, int64_t max_rel,
                          int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk,
                          void* out, int64_t ldo, int64_t sbo, void* workspace, int64_t workspace_bytes,
                          const int32_t* kv_len, cn_stream_t stream);
/* Same attention on a table the caller already has: element (b, i, head, r) = (q_i + v_head).P[r,head] at
 * table[b*table_sb + i*table_ld + head*table_hs + r], f32 or bf16 (table_is_bf16).  The host side folds
 * linear_q, linear_pos and pos_bias_v into h*(2m+1) extra output columns of the fused QKV projection
 * (modules.RelMultiHeadedAttention.project_qkv_rel), so the table costs no launch of its own. */
int cn_rel_attention_bf16_table(const void* Q, int64_t ldq, int64_t sbq, const void* K, int64_t ldk, int64_t sbk,
                                const void* V, int64_t ldv, int64_t sbv, const uint8_t* mask, int64_t mb, int64_t mi,
                                const void* table, int table_is_bf16, int64_t table_ld, int64_t table_sb,
                                int64_t table_hs, int64_t max_rel, int64_t B, int64_t h, int64_t Tq, int64_t Tk,
                                int64_t dk, void* out, int64_t ldo, int64_t sbo, const int32_t* kv_len,
                                cn_stream_t stream);

/* K11  sampled-alignment rescoring: per sample the mean ranker log-probability of the predicted ids
 * over the valid tokens (exact zeros do not count), then the first arg-max over the S samples.
 *   replaces: gather / masked_fill / sum / count / max         src/models/cassnat.py:517-520
 * lm_out f32 [B*S, U, V] (row_stride = elements between utterance-samples, tok_stride between tokens);
 * ids i32 [B*S,U]; tgt_mask u8 [B*S,U]; mean_out f32 [B*S]; best i64 [B].  S <= 32. */
int cn_rescore_select(const float* lm_out, int64_t row_stride, int64_t tok_stride, const int32_t* ids,
                      const uint8_t* tgt_mask, int64_t B, int64_t S, int64_t U, float* mean_out, int64_t* best,
                      cn_stream_t stream);

/* ============================ backward pass of the training step (cfg5) ============================
 * What loss.backward() (src/tasks/cassnat_task.py:291) runs through ATen autograd for the blocks on
 * the path.  Every reduction is order-deterministic (no floating-point atomics). */

/* Batched strided fp32 GEMM:  C[z](m,n) = alpha * sum_k A[z](m,k) * B[z](n,k) + beta * C[z](m,n),
 * z = (z1,z2) with independent element strides; C rows have stride c_m and unit column stride.
 * Expresses dX = dY W, dW = dY^T X and the five contractions of attention backward without copies.
 *   replaces: autograd of F.linear / torch.matmul   src/models/modules/attention.py:15-24,55-65,104-143 */
int cn_gemm_f32_strided(const float* A, int64_t a_z1, int64_t a_z2, int64_t a_m, int64_t a_k,
                        const float* B, int64_t b_z1, int64_t b_z2, int64_t b_n, int64_t b_k,
                        float* C, int64_t c_z1, int64_t c_z2, int64_t c_m,
                        int64_t nz1, int64_t nz2, int64_t M, int64_t N, int64_t K,
                        float alpha, float beta, cn_stream_t stream);

/* dpre = alpha * dy * act'(saved) (ReLU: saved = activation output; Swish: saved = pre-activation;
 * CN_ACT_NONE: saved unused) and colsum[n] (+)= sum_m dpre[m,n]  (the bias gradient).
 * dpre or colsum may be NULL.  workspace: cn_colsum_workspace_bytes(M,N) when colsum != NULL.
 *   replaces: autograd of bias add / ReLU / Swish    src/models/modules/positionff.py:15-16 */
int64_t cn_colsum_workspace_bytes(int64_t M, int64_t N);
int cn_act_bwd_colsum(const float* dy, int64_t lddy, const float* saved, int64_t lds, int act, float alpha,
                      float* dpre, int64_t ldp, int64_t M, int64_t N, float* colsum, int accumulate,
                      void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* The same for the tensor-core backward of a linear layer: dy / saved in f32 or bf16 (flags), dpre written in bf16
 * (the operand of the dX and dW GEMMs); dpre_bf16 or colsum may be NULL. */
int cn_act_bwd_colsum_bf16(const void* dy, int dy_is_bf16, int64_t lddy, const void* saved, int saved_is_bf16, int64_t lds,
                           int act, float alpha, void* dpre_bf16, int64_t ldp, int64_t M, int64_t N, float* colsum,
                           int accumulate, void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* Backward of y = a*(x-mean)/(std_unbiased+eps)+b: dx [rows,d], da [d], db [d].
 *   replaces: autograd of LayerNorm.forward          src/models/modules/norm.py:15-18 */
int64_t cn_layernorm_bwd_workspace_bytes(int64_t rows, int64_t d);
int cn_layernorm_bwd(const float* x, int64_t ldx, const float* dy, int64_t lddy, const float* a, float eps,
                     int64_t rows, int64_t d, float* dx, int64_t lddx, float* da, float* db,
                     void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* Softmax backward of masked attention, in place: S [B,h,Tq,Tk] (scaled raw scores) <- probabilities
 * as the forward used them; dP [B,h,Tq,Tk] (= dO V^T) <- gradient at the scaled scores.  Fully masked
 * rows: uniform probabilities and zero score gradient (attention.py:19-21); zero_masked = 1 applies the
 * relative variant's re-zeroing (attention.py:138).  mask as in cn_attention_f32.  dP may be NULL: forward only
 * (the materialised attention of the training path with dropout on the probabilities). */
int cn_attn_softmax_bwd(float* S, float* dP, const uint8_t* mask, int64_t mb, int64_t mi, int64_t B, int64_t h,
                        int64_t Tq, int64_t Tk, int zero_masked, cn_stream_t stream);

/* Attention backward for few query rows (Tq <= 128, d_k == 64: the label-token side of the training step), ONE launch:
 * recomputes the probabilities of attention() (masking as cn_attn_softmax_bwd, plain variant; dropout on the
 * probabilities regenerated from (seed, site) exactly as cn_dropout applied it to the [B,h,Tq,Tk] probabilities in
 * the forward), and writes dq / dk / dv.  q, k, v, o (the forward's context), go (its gradient) and the three
 * gradients share one element type (is_bf16 ? bf16 : f32); all are [B,T,h*64] views with row / batch pitches in
 * elements.  mask u8 [B,1|Tq,Tk] or NULL (mb / mi as cn_attention_f32).
 *   replaces (training): autograd through attention()           src/models/modules/attention.py:13-24 */
int cn_attention_bwd_small(const void* q, int64_t ldq, int64_t sbq, const void* k, int64_t ldk, int64_t sbk,
                           const void* v, int64_t ldv, int64_t sbv, const void* o, int64_t ldo, int64_t sbo,
                           const void* go, int64_t ldg, int64_t sbg, int is_bf16, const uint8_t* mask, int64_t mb,
                           int64_t mi, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk, float p_drop,
                           const int64_t* seed_dev, int64_t seed, int64_t site, void* dq, int64_t lddq, int64_t sbdq,
                           void* dk_out, int64_t lddk, int64_t sbdk, void* dv, int64_t lddv, int64_t sbdv,
                           cn_stream_t stream);

/* The forward of the same attentions with dropout on the probabilities (attention.py:22-23), one launch: online
 * softmax over the key tiles, the keep mask a function of (seed, site, index into the [B,h,Tq,Tk] probabilities) --
 * the mask cn_attention_bwd_small regenerates.  Tq <= 128, d_k == 64; element type as above. */
int cn_attention_fwd_small(const void* q, int64_t ldq, int64_t sbq, const void* k, int64_t ldk, int64_t sbk,
                           const void* v, int64_t ldv, int64_t sbv, int is_bf16, const uint8_t* mask, int64_t mb,
                           int64_t mi, int64_t B, int64_t h, int64_t Tq, int64_t Tk, int64_t dk, float p_drop,
                           const int64_t* seed_dev, int64_t seed, int64_t site, void* out, int64_t ldo, int64_t sbo,
                           cn_stream_t stream);

/* Dropout (training): out[i] = keep(i) ? x[i] / (1 - p) : 0 with keep a counter-based hash of (seed, site, i):
 * stateless, so the backward pass regenerates the mask from the same (seed, site) instead of storing it.
 *   replaces: nn.Dropout                 src/models/modules/utils.py:21,32; attention.py:22-23,140-141;
 *                                        positionff.py:12,16; embedding.py:17,31,49,60
 * x / out f32 or bf16 (flags), n elements, in place allowed.  seed_dev (i64 [1], may be NULL) overrides `seed`: a
 * step captured in a CUDA graph advances the device-side seed between replays.  site: a distinct id per dropout
 * call site of a step. */
int cn_dropout(const void* x, int x_is_bf16, void* out, int out_is_bf16, int64_t n, float p, const int64_t* seed_dev,
               int64_t seed, int64_t site, cn_stream_t stream);
/* residual + scale * dropout(x) in one pass (SublayerConnection, src/models/modules/utils.py:32); residual f32 [n] or NULL
 * (then: scale * dropout(x), the branch gradient of the same expression), out f32 [n]; same mask as cn_dropout for the
 * same (seed, site). */
int cn_dropout_add(const void* x, int x_is_bf16, const float* residual, float* out, int64_t n, float p, float scale,
                   const int64_t* seed_dev, int64_t seed, int64_t site, cn_stream_t stream);

/* Label-smoothing KL loss over log-softmax rows and its gradient at the logits:
 *   loss = sum_{rows with target != pad} KL(w || p) / #such rows,  w = (1-s) one-hot + s/(V-1) elsewhere
 *   dlogits[r,:] = grad_scale * (p - w) / #rows   (0 for pad rows);  dlogits may be NULL.
 *   replaces: LabelSmoothing.forward + autograd      src/utils/loss.py:21-35
 * workspace: (R + 4) * 4 bytes. */
int cn_label_smoothing(const float* logp, int64_t ld, int64_t R, int64_t V, const int64_t* target, int64_t pad,
                       float smoothing, float grad_scale, float* loss, float* dlogits, int64_t ldg,
                       void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* Relative-position attention backward helpers (attention.py:104-143):
 *   cn_rel_bias_add:  S[z,i,j] += scale * Tb[z,i,clamp(j-i,-m,m)+m],  Tb [Z,Tq,2m+1] = (q+v) P^T
 *   cn_rel_bucket:    G[z,i,r]  = sum over keys j with clamp(j-i,-m,m)+m == r of dS[z,i,j]
 * (Z = B*h; the remaining contractions are cn_gemm_f32_strided calls.) */
int cn_rel_bias_add(float* S, const float* Tb, int64_t Z, int64_t Tq, int64_t Tk, int64_t max_rel, float scale,
                    cn_stream_t stream);
int cn_rel_bucket(const float* dS, float* G, int64_t Z, int64_t Tq, int64_t Tk, int64_t max_rel, cn_stream_t stream);

/* y = act(x), n elements, into f32 and/or bf16 (training keeps the pre-activation x for backward). */
int cn_act_fwd(const float* x, int act, float* y_f32, void* y_bf16, int64_t n, cn_stream_t stream);

/* Backward of cn_glu_dwconv_groupnorm_act: recomputes GLU / depthwise conv / GroupNorm statistics from y,
 * then dy [B,L,2C] (contiguous f32), d_dw_weight [C,ksz], d_dw_bias [C], dgamma [C], dbeta [C]; dout [B,L,C].
 *   replaces: autograd of ConvModule.forward core      src/models/modules/conformer_related.py:33-38 */
int64_t cn_conv_module_bwd_workspace_bytes(int64_t B, int64_t L, int64_t C, int64_t ksz);
int cn_glu_dwconv_groupnorm_act_bwd(const void* y, int y_is_bf16, int64_t ldy, int64_t sby, const float* dw_weight,
                                    const float* dw_bias, const float* gamma, const float* beta, float eps, int act,
                                    int64_t B, int64_t L, int64_t C, int64_t ksz, const float* dout, float* dy,
                                    float* d_dw_weight, float* d_dw_bias, float* dgamma, float* dbeta,
                                    void* workspace, int64_t workspace_bytes, cn_stream_t stream);

/* Backward GEMMs on tcgen05 without transposed copies: C[M,N] = op(A) op(W)^T, bf16 operands, fp32 accumulate.
 *   a_mn = 0: A is [M,K] row-major (pitch lda);  a_mn = 1: A is stored [K,M] (the contraction index is the row)
 *   w_mn = 1: W is stored [K,N] (required);      K-major W with a_mn = 0 is cn_gemm_bf16.
 * The MN-major operands are fed to the MMA as they are (TMA boxes of 64 contraction rows x 64 elements, MN-major
 * shared-memory descriptors):  dX = dY W -> (A = dY, a_mn 0; W = weight [out,in], w_mn 1);
 *                              dW = dY^T X -> (A = dY, a_mn 1; W = X, w_mn 1).
 *   replaces: autograd of F.linear                            src/models/modules/positionff.py:15-16 etc. */
int cn_gemm_bf16_mn(const void* A, int64_t lda, int a_mn, const void* W, int64_t ldw, int w_mn, float* out_f32,
                    void* out_bf16, int64_t ldc, int64_t M, int64_t N, int64_t K, cn_stream_t stream);

/* out(bf16)[C, ldo] = in^T, in [R,C] f32 or bf16 with row stride ldi: operand staging of the
 * tensor-core backward GEMMs (dX = dY W needs W^T rows, dW = dY^T X needs dY^T and X^T rows). */
int cn_transpose_bf16(const void* in, int in_is_bf16, int64_t ldi, int64_t R, int64_t C, void* out, int64_t ldo,
                      cn_stream_t stream);

/* Data-parallel training: caps the CTA count of the persistent tensor-core GEMMs (0 = one per SM, the default) so
 * that a collective running on another stream (the bucketed gradient all_reduce of tasks/base_task.py:82-83's DDP)
 * finds free SMs while backward is still running.  Returns the previous value.  Takes effect at the next launch
 * (a captured CUDA graph keeps the grid it was captured with). */
int cn_set_sm_limit(int n_ctas);

#ifdef __cplusplus
}
#endif
#endif /* CASSNAT_B200_H_ */
