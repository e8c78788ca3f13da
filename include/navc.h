/*
 * navc.h -- C ABI of libnavc.so: hand-written sm_100a kernels for the hot path of
 * yangbang18/Non-Autoregressive-Video-Captioning (encoder, BERT-style decoder, vocabulary
 * projection, iterative-refinement decode step).
 *
 * The reference has no FFI layer: its boundary is the Python object API
 * (models.get_model -> Seq2Seq, models.Translator, decoding.generate).  The Python host package
 * mirrors that API and calls these entry points through ctypes; each entry point names the
 * reference code (file:line under the reference repo) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; plain C types only.
 *   - the callee never allocates, frees or synchronises; the caller owns all buffers.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = ok, non-zero = error; navc_last_error() returns a message (thread local).
 *   - row-major everywhere; "ld" = leading dimension in elements.
 *   - token ids / categories are int64 (torch.long), as in the reference.
 *   - bf16 "hi/lo" pairs: hi = bf16_rn(x), lo = bf16_rn(x - hi); x ~= hi + lo to ~2^-17 relative.
 */
#ifndef NAVC_H
#define NAVC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAVC_VERSION 4

/* token ids, config/Constants.py:1-6 */
#define NAVC_PAD 0
#define NAVC_UNK 1
#define NAVC_BOS 2
#define NAVC_EOS 3
#define NAVC_MASK 4
#define NAVC_VIS 5

/* activation codes (models/bert.py:9-19 ACT2FN, plus tanh/sigmoid for the highway gate) */
enum { NAVC_ACT_NONE = 0, NAVC_ACT_GELU_NEW = 1, NAVC_ACT_GELU = 2, NAVC_ACT_RELU = 3, NAVC_ACT_SWISH = 4 };

/* self-attention mask kinds, models/Decoder.py:105-124 */
enum { NAVC_MASK_KEYPAD = 0 /* NARFormer */, NAVC_MASK_CAUSAL = 1 /* ARFormer */, NAVC_MASK_SELF = 2 /* SelfMask */ };

/* tensor-core operand modes of the tcgen05 GEMMs */
enum { NAVC_TC_BF16 = 1 /* one product: hi*hi */, NAVC_TC_BF16X3 = 3 /* hi*hi + hi*lo + lo*hi */ };

/* Epilogue applied to a GEMM tile before it is stored:
 *   v = acc + bias[col]; v = act(v); v += residual[row,col] (fp32, or res_hi + res_lo); if (row_tokens[row]==PAD) v = 0;
 * (bias / activation: nn.Linear + ACT2FN, models/bert.py:227-230; residual: bert.py:196-197, 243;
 *  row mask: `* non_pad_mask`, bert.py:271-272, 293-294, 298-299).  Any output may be NULL. */
typedef struct {
    const float* bias;          /* [N] or NULL */
    const float* residual;      /* [M, ld_res] or NULL */
    const int64_t* row_tokens;  /* [M] or NULL */
    int32_t act;                /* NAVC_ACT_* */
    int32_t ld_res;
    float* out_f32;             /* [M, ld_out] or NULL */
    uint16_t* out_hi;           /* bf16 bits [M, ld_out] or NULL */
    uint16_t* out_lo;           /* bf16 bits [M, ld_out] or NULL */
    int32_t ld_out;
    int32_t reserved;
    int32_t split_k;            /* tcgen05 path: > 1 splits the K loop over that many CTAs per output tile;
                                   implies accumulate (out_f32 only, no bias/act/residual/mask) */
    int32_t accumulate;         /* != 0: out_f32 += tile (atomic float adds; caller zero-fills first) */
    const uint16_t* res_hi;     /* tcgen05 path: residual given as a bf16 hi/lo pair [M, ld_res] (res_lo may be */
    const uint16_t* res_lo;     /* NULL); needs bf16-only outputs (out_f32 == NULL, residual == NULL), N % 8 == 0 */
    const int32_t* m_dev;       /* tcgen05 path: optional DEVICE scalar; only the first min(M, *m_dev) rows are computed
                                   (packed-row decoding: the row count is known on the device only) */
    int32_t m_hint;             /* optional HOST estimate of *m_dev (0 = none): only steers the tile shape / cluster choice
                                   of the launch, never the result */
    int32_t pad_;
} navc_epilogue_t;

int navc_version(void);
const char* navc_last_error(void);
/* One-time per-process setup for `device`: opt-in shared memory sizes, driver entry points. */
int navc_init(int device);
/* Number of SMs of the current device (grid sizing of the persistent kernels). */
int navc_sm_count(void);

/* ---- dense layers -------------------------------------------------------------------------- */
/* Y[M,N] = epilogue(X[M,K] * W[N,K]^T), fp32 CUDA-core path (reference-exact mode).
 * Replaces nn.Linear call sites: models/bert.py:146-152,193,228,241; models/Encoder.py:13-24;
 * models/Predictor.py:16-21; models/__init__.py:83.  K % 4 == 0, ldx % 4 == 0, ldw % 4 == 0. */
int navc_linear_f32(const float* x, int ldx, const float* w, int ldw, int M, int N, int K,
                    const navc_epilogue_t* epi, void* stream);

/* Same contract on the tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
 * Operands are bf16 (hi) or split bf16 (hi+lo); accumulation fp32.  K % 8 == 0, ld % 8 == 0 (the
 * K tail of the last 64-wide block is zero-filled by TMA).  x_lo / w_lo may be NULL when
 * mode == NAVC_TC_BF16.  Gradient GEMMs of the training path are this same entry point on
 * transposed operands (navc_transpose_pack) with split_k / accumulate. */
int navc_linear_tc(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx,
                   const uint16_t* w_hi, const uint16_t* w_lo, int ldw, int M, int N, int K,
                   const navc_epilogue_t* epi, void* stream);

/* navc_cross_attention_bwd_tc with dK / dV leaving as the bf16 hi (/ lo) operand pair of the K|V projection's gradient
 * GEMMs -- d_kv_hi / d_kv_lo [N * E, ld_dkv] (K at column 0, V at column D of the given pointers; the values
 * navc_transpose_pack would make of the fp32 result) -- plus their column sums accumulated into kv_colsum [2 D] (the bias
 * gradient; caller zeroes or passes the gradient buffer).  Saves the fp32 d_kv round trip (3 GB at 1024 videos). */
int navc_cross_attention_bwd_tc_split(int mode, const float* q, int ldq, const float* kv, int ldkv, const int32_t* seq_off,
                                      int N, int S, int E, int D, int H, const float* d_ctx, const float* ctx, float* d_q,
                                      int ld_dq, uint16_t* d_kv_hi, uint16_t* d_kv_lo, int ld_dkv, float* kv_colsum,
                                      void* stream);
/* Two chained linear layers in ONE persistent launch (csrc/gemm2_tc.cu, G2Chain): y0 = epilogue0(x w0^T), then
 * y1 = epilogue1(y0 w1^T) -- e.g. the attention out-projection (+ residual) followed by the text -> video query projection
 * (models/bert.py:192-200 then :81-92 of the next block).  N == K (square layers), bf16 hi (/ lo) outputs in e0->out_hi/lo
 * and e1->out_hi/lo (y0 is also problem 1's A operand), the same device-side row count in both epilogues.  A tile of
 * problem 1 waits for the row block of y0 through per-row-block completion counters; one pipeline fill / drain and one
 * tile quantisation instead of two. */
int navc_linear_chain_tc(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w0_hi,
                         const uint16_t* w0_lo, int ldw0, const navc_epilogue_t* e0, const uint16_t* w1_hi,
                         const uint16_t* w1_lo, int ldw1, const navc_epilogue_t* e1, int M, int N, int K, void* stream);
/* Y = epilogue(X W^T) with fp32 operands consumed by the tensor cores as TF32 (tcgen05 kind::tf32: 10-bit mantissa, two
 * bf16-MMA time units per product against split-bf16's three): x [M, K] and w [N, K] fp32 row-major (K % 4 == 0, 16-byte
 * aligned), generic epilogue -- bias, activation, fp32 residual, row mask; fp32 and / or bf16 hi (/ lo) outputs; a
 * device-side row count in e->m_dev.  The "1e-3 logits" mode of the engine (precision 'tf32') uses it wherever the A
 * operand exists in fp32 (models/bert.py:81-92 query/key/value, :227-247 FFN; Encoder.py:9-38). */
int navc_linear_tf32(const float* x, int ldx, const float* w, int ldw, int M, int N, int K, const navc_epilogue_t* e,
                     void* stream);
/* Weight gradient on the tensor cores, straight from row-major operands (no transposed copies):
 *   out_f32[n, k] += sum_r dY[r, n] * X[r, k]      dY [rows, n_out] (ld_dy), X [rows, k_in] (ld_x), bf16 hi(/lo).
 * Both operands are consumed MN-major (TMA boxes of 64 rows x 64 columns).  The epilogue only honours
 * out_f32 / ld_out / split_k (atomic accumulate: the caller zero-fills out_f32 first). */
int navc_wgrad_tc(int mode, const uint16_t* dy_hi, const uint16_t* dy_lo, int ld_dy, const uint16_t* x_hi,
                  const uint16_t* x_lo, int ld_x, int rows, int n_out, int k_in, const navc_epilogue_t* epi,
                  void* stream);
/* Data gradient dX[rows, k_in] = dY[rows, n_out] W[n_out, k_in] (+ e->residual) on the tensor cores, straight from
 * the forward's row-major bf16 weight copies (B operand consumed MN-major: no transposed weights).  fp32 output
 * (e->out_f32, e->ld_out), optional fp32 residual; dY pad columns [n_out, ld_dy) must be finite. */
/* Tail split of navc_linear_tc (pair epilogue, K >= 1024): the tiles of the last, partly filled wave of the
 * persistent grid are split along K across the idle SMs; the finishing CTA sums the partner CTAs' partial
 * accumulators (handed over through an L2-resident workspace) before its epilogue.  Measured slower than leaving the
 * SMs idle, so it is OFF unless $NAVC_STREAMK=1 or navc_set_streamk(1) (returns the previous
 * setting).  navc_streamk_error: 1 if a finishing CTA ever gave up waiting for its partners (never expected), 0
 * otherwise, -1 before navc_init; synchronises the device. */
int navc_set_streamk(int on);
int navc_streamk_error(void);
int navc_dgrad_tc(int mode, const uint16_t* dy_hi, const uint16_t* dy_lo, int ld_dy, const uint16_t* w_hi,
                  const uint16_t* w_lo, int ld_w, int rows, int n_out, int k_in, const navc_epilogue_t* e,
                  void* stream);

/* fp32 -> bf16 hi/lo split of a contiguous buffer (weights are split once when packed). */
int navc_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, int64_t n, void* stream);
/* The inverse: out = hi + lo (lo may be NULL).  Used at the API boundary when a caller asks for the
 * fp32 hidden states of a residual stream that the tensor-core path keeps as bf16 hi/lo pairs. */
int navc_join_bf16(const uint16_t* hi, const uint16_t* lo, float* out, int64_t n, void* stream);

/* ---- vocabulary projection with on-the-fly softmax statistics ------------------------------- */
/* For logits[M,V] = H[M,K] * Wv[V,K]^T (+bias), never written to memory, emit per row and per
 * column tile t (tile width = navc_vocab_tile(tc)) the partial triple
 *   part_max[row,t] = max_c logit, part_sum[row,t] = sum_c exp(logit - max), part_idx[row,t] = argmax
 * (lowest column on ties) and, if `target` != NULL, target_logit[row] = logit[row, target[row]].
 * Replaces model.tgt_word_prj + F.softmax + max (decoding/algorithms.py:7-15, 149, 197-200). */
int navc_vocab_tile(int tc);
int navc_vocab_partials_f32(const float* h, int ldh, const float* w, int ldw, const float* bias,
                            int M, int V, int K, float* part_max, float* part_sum, int32_t* part_idx,
                            const int64_t* target, float* target_logit, void* stream);
int navc_vocab_partials_tc(int mode, const uint16_t* h_hi, const uint16_t* h_lo, int ldh,
                           const uint16_t* w_hi, const uint16_t* w_lo, int ldw, const float* bias,
                           int M, int V, int K, float* part_max, float* part_sum, int32_t* part_idx,
                           const int64_t* target, float* target_logit, void* stream);
/* log_softmax over rows of a materialised logits matrix, in place allowed (seq2seq.py:102-103). */
int navc_log_softmax(const float* logits, float* out, int M, int V, int ld, void* stream);
/* Same with separate leading dimensions (training: logits come from a GEMM with a padded ld). */
int navc_log_softmax_ld(const float* logits, int ld_in, float* out, int ld_out, int M, int V, void* stream);

/* ---- encoder ------------------------------------------------------------------------------- */
/* Highway gate + frame mean + (eval) BatchNorm + temporal concat for one modality
 * (models/Encoder.py:19-25; models/joint_representation.py:27, 43-51).
 *   x  [B*F, D]  = Linear0(feats) ;  yg [B*F, 2D] = [w1 x + b1 | w2 x + b2] (pre-activation)
 *   o = sigmoid(g)*x + (1-sigmoid(g))*tanh(y)        (gate==0: o = x + tanh(y), yg is [B*F, D])
 *   enc_hidden[b,:] (+)= mean_f(o) / n_modalities    (accumulate != 0 adds to the existing value)
 *   enc_out[b, slot*F + f, :] = bn ? (o - rm) / sqrt(rv + eps) * bw + bb : o
 * enc_out has E rows per video (ld = D); optional bf16 hi/lo copies of enc_out. */
int navc_highway_bn(const float* x, const float* yg, int gate, int B, int F, int D, int E, int slot,
                    int n_modalities, int accumulate, const float* bn_rm, const float* bn_rv,
                    const float* bn_w, const float* bn_b, float bn_eps, float* enc_hidden,
                    float* enc_out, uint16_t* enc_hi, uint16_t* enc_lo, void* stream);

/* Same with the norm_type='ln' encoder norm (models/joint_representation.py:20, 46-47): nn.LayerNorm(D) over the
 * features of every frame row (eps 1e-5, biased variance) instead of the BatchNorm affine.  D <= 1024. */
int navc_highway_ln(const float* x, const float* yg, int gate, int B, int F, int D, int E, int slot,
                    int n_modalities, int accumulate, const float* ln_w, const float* ln_b, float ln_eps,
                    float* enc_hidden, float* enc_out, uint16_t* enc_hi, uint16_t* enc_lo, void* stream);

/* Length head (models/Predictor.py:23-30) + the frame mean reused by enhance_input=2
 * (models/Decoder.py:137): enc_mean[b,:] = mean_e enc_out[b,e,:];
 * pred_length[b,:] = log_softmax(W2 relu(W1 enc_mean + b1) + b2).  w1 may be NULL (mean only). */
int navc_length_head(const float* enc_out, int B, int E, int D, const float* w1, const float* b1,
                     const float* w2, const float* b2, int max_len, float* enc_mean,
                     float* pred_length, void* stream);

/* ---- decoder ------------------------------------------------------------------------------- */
/* BertEmbeddings (models/bert.py:70-96): out[n,s,:] = LayerNorm(word[tok] + pos[s] +
 * cat[category[n / group]] + extra[n / group]); cat_emb / extra may be NULL.  `group` = decoder rows
 * per video (length candidates); category and extra are per video.  tokens [N,S]; outputs [N*S, D]. */
int navc_embed_ln(const int64_t* tokens, const int64_t* category, const float* word_emb,
                  const float* pos_emb, const float* cat_emb, const float* extra, int group,
                  const float* ln_w, const float* ln_b, float eps, int N, int S, int D,
                  float* out_f32, uint16_t* out_hi, uint16_t* out_lo, void* stream);

/* LayerNorm over rows (+ optional `* non_pad_mask`): with_layernorm=True variant of
 * BertSelfOutput / BertOutput (models/bert.py:198-199, 244-245).  In place allowed. */
int navc_layernorm(const float* x, const float* w, const float* b, float eps, const int64_t* row_tokens,
                   int M, int D, float* out_f32, uint16_t* out_hi, uint16_t* out_lo, void* stream);

/* Token self-attention core (models/bert.py:154-176): per (row n, head h)
 * softmax(mask_fill(Q K^T / sqrt(dk), -1e7)) V with the mask derived in-kernel from the tokens
 * (key j masked iff tokens[n,j]==PAD; + causal / diagonal per mask_kind; `watch` as
 * models/Decoder.py:23-39).  qkv [N*S, ld] holds Q | K | V at column offsets 0, D, 2D.
 * ctx outputs [N*S, D] (merged heads).  probs (optional) is [H, N, S, S]. */
int navc_self_attention(const float* qkv, int ld, const int64_t* tokens, int N, int S, int D, int H,
                        int mask_kind, int watch, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                        float* probs, void* stream);

/* Text-to-video cross-attention core (models/bert.py:282-290 with the all-False mask of
 * models/Decoder.py:127-128).  q [N*S, ldq]; kv [(N/group)*E, ldkv] holds K | V at column offsets
 * 0 and D for this layer (computed once per video, shared by its `group` length candidates --
 * replaces the physical x lbs repeat of misc/utils.py:205-213).  probs (optional) [H, N, S, E]. */
int navc_cross_attention(const float* q, int ldq, const float* kv, int ldkv, int N, int S, int E,
                         int D, int H, int group, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                         float* probs, void* stream);

/* tcgen05 versions of the two attention cores (dk == 64; S <= 32 / E <= 128), same arithmetic and
 * masks; operands are the bf16 hi (+lo when mode == NAVC_TC_BF16X3) copies written by the
 * projection GEMMs: qkv_* [N*S, ld] = Q | K | V at column offsets 0, D, 2D; kv_* [(N/group)*E, ldkv]
 * = K | V at 0, D.  Attention probabilities are not materialised (use the fp32 cores for those). */
int navc_self_attention_tc(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                           const int64_t* tokens, int N, int S, int D, int H, int mask_kind, int watch,
                           float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream);
int navc_cross_attention_tc(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                            const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv, int N, int S, int E,
                            int D, int H, int group, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                            void* stream);

/* ---- packed rows ---------------------------------------------------------------------------
 * The reference runs every decoder op over all N*S positions although positions >= len are PAD and
 * every one of their outputs is multiplied by non_pad_mask = 0 (models/bert.py:271-299).  In the
 * packed layout only the sum(len) real positions exist: row seq_off[n] + s holds position (n, s).
 * Row counts are device-side (m_dev / count pointers), launch shapes stay at the N*S maximum, so
 * the refinement loop remains one CUDA graph. */
/* seq_off[0..N] = exclusive prefix sum of lens (seq_off[N] = row count); rowmap[i] = n*S + s. */
int navc_pack_rows(const int32_t* lens, int N, int S, int32_t* seq_off, int32_t* rowmap, void* stream);
/* navc_embed_ln over packed rows: row i < seq_off[N] is position rowmap[i]; also emits the row's
 * token id (tok_out[i], int64) for the GEMM epilogues' `* non_pad_mask`. */
int navc_embed_ln_packed(const int64_t* tokens, const int64_t* category, const float* word_emb,
                         const float* pos_emb, const float* cat_emb, const float* extra, int group,
                         const float* ln_w, const float* ln_b, float eps, int N, int S, int D,
                         const int32_t* seq_off, const int32_t* rowmap, int64_t* tok_out,
                         float* out_f32, uint16_t* out_hi, uint16_t* out_lo, void* stream);
/* tcgen05 attention cores over packed rows (same arithmetic and masks as the padded entry points;
 * tokens stays the padded [N,S] canvas, qkv / q / ctx are packed). */
int navc_self_attention_tc_packed(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                  const int64_t* tokens, const int32_t* seq_off, int N, int S, int D, int H,
                                  int mask_kind, int watch, float* ctx_f32, uint16_t* ctx_hi,
                                  uint16_t* ctx_lo, void* stream);
int navc_cross_attention_tc_packed(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                   const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv,
                                   const int32_t* seq_off, int N, int S, int E, int D, int H, int group,
                                   float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream);
/* Second-generation attention cores over packed rows (csrc/attention2_tc.cu): ONE persistent, warp-specialised,
 * two-stage pipelined kernel per launch (TMA producer / MMA issuer / two alternating softmax + epilogue groups).
 * Self-attention tiles are 96-row windows of the packed row space: navc_pack_tiles fills tile_seq[0..n_tiles] (first
 * sequence of every window; n_tiles = ceil(max rows / navc_attention_window())) once per set of candidate lengths.
 * Same arithmetic contract as navc_self_attention / navc_cross_attention (models/bert.py:139-179); bf16 hi(/lo) outputs. */
int navc_attention_window(void);
int navc_pack_tiles(const int32_t* seq_off, int N, int32_t* tile_seq, int n_tiles, void* stream);
int navc_self_attention_tc_tiles(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                 const int64_t* tokens, const int32_t* seq_off, const int32_t* tile_seq,
                                 int n_tiles, int rows, int N, int S, int D, int H, int mask_kind, int watch,
                                 uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream);
int navc_cross_attention_tc_tiles(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                  const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv,
                                  const int32_t* seq_off, int rows, int N, int S, int E, int D, int H, int group,
                                  uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream);
/* The same with the row count of the packed buffers given explicitly (`rows` >= seq_off[N]; the _packed entry
 * points assume buffers of N*S rows): the training path sizes its buffers to the real positions. */
int navc_self_attention_tc_rows(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                const int64_t* tokens, const int32_t* seq_off, int rows, int N, int S, int D, int H,
                                int mask_kind, int watch, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                                void* stream);
int navc_cross_attention_tc_rows(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                 const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv, const int32_t* seq_off,
                                 int rows, int N, int S, int E, int D, int H, int group, float* ctx_f32,
                                 uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream);
/* out[k, :] = in[rows[k], :] for k < *count (bf16 hi / lo pairs, D % 8 == 0; lo may be NULL). */
int navc_gather_rows(const uint16_t* in_hi, const uint16_t* in_lo, int D, const int32_t* rows,
                     const int32_t* count, int max_rows, uint16_t* out_hi, uint16_t* out_lo, void* stream);
/* Weight refresh after an optimizer step (parameters keep their storage, values changed): one launch over a device
 * table items[n][5] of int64 {src fp32 pointer, dst fp32 pointer or 0, bf16 hi pointer or 0, lo pointer or 0, elements}:
 * dst[i] = src[i] (rows of a concatenated operand) and hi / lo = the split-bf16 copies of src. */
int navc_refresh_pack(const int64_t* items, int n_items, void* stream);
/* Two navc_gather_rows through the same row list in one launch: oa[k] = a[rows[k]], ob[k] = b[rows[k]]. */
int navc_gather_rows2(const uint16_t* a_hi, const uint16_t* a_lo, uint16_t* oa_hi, uint16_t* oa_lo, const uint16_t* b_hi,
                      const uint16_t* b_lo, uint16_t* ob_hi, uint16_t* ob_lo, int D, const int32_t* rows, const int32_t* count,
                      int max_rows, void* stream);
/* Ordered compaction of the positions navc_refine_step selected (its flag mode: sel_rows == NULL, sel_slot = 0/1 per real
 * packed row).  In place: slot[r] := number of selected rows before r (the compact index of a selected row; slot has
 * max_rows + 1 entries); rows[k] = k-th selected packed row, ascending; seq_off_c[n] = slot[seq_off[n]], n = 0..N: the
 * packed offsets of the compacted row space -- the selected rows of a sequence (and of a video's candidates) stay
 * contiguous, so the packed cross-attention core and the GEMMs run on them unchanged; *count = seq_off_c[N].
 * Used for the LAST decoder layer of a refinement pass: only re-masked positions are read from that pass
 * (decoding/algorithms.py:262-268 assigns tokens / probabilities at mask_ind only), so everything behind the last
 * self-attention core runs on those rows alone. */
int navc_compact_rows(int32_t* slot, const int32_t* seq_off, int N, int max_rows, int32_t* rows, int32_t* count,
                      int32_t* seq_off_c, void* stream);
/* navc_vocab_partials_tc over the first min(M, *m_dev) rows. */
int navc_vocab_partials_tc_dyn(int mode, const uint16_t* h_hi, const uint16_t* h_lo, int ldh,
                               const uint16_t* w_hi, const uint16_t* w_lo, int ldw, const float* bias,
                               int M, int V, int K, const int32_t* m_dev, float* part_max, float* part_sum,
                               int32_t* part_idx, void* stream);

/* ---- autoregressive beam search on the device (models/Translator.py:94-161, models/Beam.py) ----
 * One new position per beam row and step.  k_cache / v_cache: [T, N, D] fp32 per layer, written at step position
 * `pos` by the row that computed it; anc [N, T] int32: anc[r, j] = the row that cached position j of r's prefix
 * (beams are re-ordered by re-gathering this table, never the cache); hist [N, T] int64: r's tokens
 * (hist[r, 0] = BOS), used for the key padding mask.  qkv [N, ld >= 3D] fp32 = the new position's projections. */
int navc_self_attention_step(const float* qkv, int ld, float* k_cache, float* v_cache, const int32_t* anc,
                             const int64_t* hist, int N, int T, int D, int H, int pos, int watch, float* ctx_f32,
                             uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream);
/* Top-K over beam x vocab of (beam score + log_softmax(logits row)) per video, straight from the logits
 * [B*K, ld] (Translator.py:113-114, Beam.py:68-83): beams whose last token hist[r, pos] is EOS contribute -1e20,
 * at the first step (first != 0) only beam 0 counts.  best_scores / best_ids [B, K] descending; id = beam * V + word;
 * K <= 8; ties -> lowest id. */
int navc_beam_topk(const float* logits, int ld, int B, int K, int V, const float* scores, const int64_t* hist, int T,
                   int pos, int first, float* best_scores, int64_t* best_ids, void* stream);
/* Beam.advance (Beam.py:68-117) for all B videos after step t (1-based): best_scores / best_ids [B, K] = top-K of
 * (beam score + log-prob) over beam x vocab; re-gathers hist / anc (in -> out, distinct buffers), updates the
 * beam scores, records hypotheses that emitted EOS (fin_*: [B, cap] score / length, [B, cap, T] tokens without
 * BOS), marks a video done after `want` finished hypotheses or at t + 1 == max_len, counts them in n_done. */
int navc_beam_advance(const float* best_scores, const int64_t* best_ids, int B, int K, int V, int t, int max_len,
                      int want, int T, const int64_t* hist_in, int64_t* hist_out, const int32_t* anc_in,
                      int32_t* anc_out, float* scores, int32_t* done, int32_t* fin_count, float* fin_score,
                      int32_t* fin_len, int64_t* fin_tok, int cap, int32_t* n_done, void* stream);

/* ---- iterative refinement (decoding/na_generate.py, decoding/algorithms.py) ----------------- */
/* Length beam + canvas (na_generate.py:33-50, 116-135): beam[b,:] = clamp(top-lbs indices of
 * pred_length[b,:] + length_bias, 4, max_len-1) (descending value, lowest index on ties);
 * smax[0] = max over the batch (int32, atomicMax; caller zeroes it).  */
int navc_length_beam(const float* pred_length, int B, int max_len, int lbs, int length_bias,
                     int32_t* beam, int32_t* smax, void* stream);
/* canvas[n,s] = tokens[n,s] = s < beam[n] ? fill : PAD ; probs[n,s] = s < beam[n] ? 0 : 1
 * (algorithms.py:291-292, 363-364).  tokens / probs may be NULL. */
int navc_init_canvas(const int32_t* beam, int N, int S, int64_t fill, int64_t* canvas, int64_t* tokens,
                     float* probs, void* stream);

/* How the step kernel merges the pass result into the state and which positions it re-masks. */
enum {
    NAVC_MERGE_NONE = 0,     /* no pass result */
    NAVC_MERGE_ALL = 1,      /* first pass: take every position (algorithms.py:238-241) */
    NAVC_MERGE_MASKED = 2,   /* only where upd_mask (algorithms.py:264-265) */
    NAVC_MERGE_EF = 3,       /* easy-first commit: min(q, remaining) most confident masked (297-309) */
};
enum {
    NAVC_SELECT_NONE = 0,    /* last step: no re-mask, emit lprobs = log(prob * teacher) */
    NAVC_SELECT_WORST = 1,   /* k = max(1, trunc(float(len)*ratio)) smallest prob*teacher (206-215) */
    NAVC_SELECT_MASKTOK = 2, /* positions whose token is MASK (algorithms.py:253-254) */
    NAVC_SELECT_GIVEN = 3,   /* caller-provided mask `given` (visual_mask, algorithms.py:327, 399) */
    NAVC_SELECT_KEEP = 4,    /* no re-mask, canvas = tokens (easy-first growth loop, 381-396) */
    NAVC_SELECT_WINDOW = 5,  /* left-to-right: the win_lo..win_hi-th set positions of `given` (297-316) */
};
typedef struct {
    /* pass result (NULL when merge == NONE) */
    const float* part_max; const float* part_sum; const int32_t* part_idx; int32_t n_tiles;
    int32_t is_ct;             /* coarse-grained template pass: prob = 0 where prediction == MASK */
    int32_t merge;             /* NAVC_MERGE_* */
    int32_t select;            /* NAVC_SELECT_* */
    int32_t q;                 /* easy-first commit width */
    float ratio;               /* NAVC_SELECT_WORST */
    int32_t win_lo, win_hi;    /* NAVC_SELECT_WINDOW */
    const int32_t* lens;       /* [N] candidate lengths (positions >= len are PAD) */
    const float* teacher;      /* [N,S] teacher probabilities or NULL (= ones) */
    const uint8_t* given;      /* [N,S] for NAVC_SELECT_GIVEN / NAVC_SELECT_WINDOW */
    int64_t* tokens;           /* [N,S] state: current hypothesis */
    float* probs;              /* [N,S] state: its probabilities */
    uint8_t* upd_mask;         /* [N,S] in: positions masked for the pass just run; out: next */
    int64_t* canvas;           /* [N,S] out: decoder input of the next pass */
    float* lprobs;             /* [N,S] out (NAVC_SELECT_NONE): log(prob * teacher) */
    int32_t* counters;         /* [2] out, atomicAdd (caller zeroes): [0] MASK tokens left in `tokens`
                                  after the merge, [1] positions selected for re-masking */
    uint8_t* visual;           /* [N,S] out or NULL: token != MASK && != PAD after the merge */
    uint8_t* masked0;          /* [N,S] out or NULL: token == MASK && s < len after the merge */
    const int32_t* seq_off;    /* NULL, or [N+1] packed-row offsets: the partials of position (n, s < len) live in
                                  row seq_off[n] + s instead of n*S + s (navc_pack_rows) */
    /* second-level packing of the vocabulary projection (needs seq_off): the next pass only needs
     * logits at the positions this step re-masks (the MERGE_MASKED merge ignores all others) */
    const int32_t* part_slot;  /* in, or NULL: partials of packed row r live in row part_slot[r] (a previous step's sel_slot) */
    int32_t* sel_rows;         /* out, or NULL: packed rows of the positions selected for re-masking, in any order */
    int32_t* sel_count;        /* out: their number (device scalar, atomicAdd; caller zeroes) */
    int32_t* sel_slot;         /* out: [N*S + 1] packed row -> index in sel_rows; with sel_rows == NULL: 0/1 selection flag of
                                  every real packed row, to be ordered by navc_compact_rows */
} navc_step_t;
/* One launch per refinement iteration: combine the vocabulary partials into (argmax, max prob),
 * apply the pad rules, merge into the state, choose the next positions to re-mask, write the next
 * canvas.  N rows of S positions. */
int navc_refine_step(const navc_step_t* p, int N, int S, void* stream);

/* Teacher re-scoring tail (algorithms.py:197-203): teacher[n,s] = exp(target_logit - max)/sum from
 * combined partials; 1.0 at pads (s >= lens[n]). */
int navc_teacher_probs(const float* part_max, const float* part_sum, int n_tiles, const float* target_logit,
                       const int32_t* lens, int N, int S, float* teacher, void* stream);
/* Shifted teacher input (algorithms.py:189-190, optional id remap 169-173):
 * out[n,0] = BOS, out[n,s] = map(tokens[n,s-1]); mapped[n,s] = map(tokens[n,s]). map may be NULL. */
int navc_teacher_inputs(const int64_t* tokens, const int64_t* map, int N, int S, int64_t* shifted,
                        int64_t* mapped, void* stream);

/* Candidate selection (na_generate.py:66-77): score[b,j] = sum_s lprobs / len^alpha; first argmax
 * over j; hyp[b,:] = tokens[b*lbs + j*, :]. */
int navc_select_best(const int64_t* tokens, const float* lprobs, const int32_t* lens, int B, int lbs,
                     int S, float alpha, int64_t* hyp, float* score, void* stream);


/* ==== training path (forward in train mode + hand-written backward; misc/run.py:254-261) ======
 * Dropout masks are never stored: keep(seed, i) is a counter-based hash of the element index, so
 * the backward kernels regenerate the forward mask from the same seed.  Scale 1/(1-p), p in [0,1). */

/* out = rowmask( drop(seed2,p2)( drop(seed1,p1)(y) + res ) ) over [M, D] (ld = D):
 * BertSelfOutput dense->dropout->+residual (models/bert.py:193-197), BertOutput
 * dense->dropout->+input->dropout (bert.py:241-247), `* non_pad_mask` (bert.py:271-299), and the
 * embedding dropout (bert.py:96) with res == NULL.  res / row_tokens may be NULL. */
int navc_drop_add(const float* y, const float* res, uint64_t seed1, float p1, uint64_t seed2, float p2,
                  const int64_t* row_tokens, int M, int D, float* out_f32, uint16_t* out_hi,
                  uint16_t* out_lo, void* stream);
/* g = dout * rowmask * m2/(1-p2);  d_res = g (may be NULL);  d_y = g * m1/(1-p1). */
int navc_drop_add_bwd(const float* dout, uint64_t seed1, float p1, uint64_t seed2, float p2,
                      const int64_t* row_tokens, int M, int D, float* d_y, float* d_res, void* stream);

/* out = drop(seed,p)(act(u)) elementwise over n values (BertIntermediate gelu, bert.py:227-230;
 * length head ReLU->Dropout, models/Predictor.py:16-21).  du = dout * m/(1-p) * act'(u). */
int navc_act_drop(const float* u, int act, uint64_t seed, float p, int64_t n, float* out_f32,
                  uint16_t* out_hi, uint16_t* out_lo, void* stream);
int navc_act_drop_bwd(const float* dout, const float* u, int act, uint64_t seed, float p, int64_t n,
                      float* du, void* stream);

/* navc_drop_add_bwd / navc_act_drop_bwd whose dY output leaves directly as the bf16 hi (/ lo) operand [M, ld_s] of the next
 * linear layer's gradient GEMMs, with its column sums (that layer's bias gradient) accumulated into colsum [D] resp. [N]
 * (or NULL) -- instead of an fp32 dY that navc_transpose_pack would re-read.  d_res (fp32, or NULL) as navc_drop_add_bwd. */
int navc_drop_add_bwd_split(const float* dout, uint64_t seed1, float p1, uint64_t seed2, float p2, const int64_t* row_tokens,
                            int M, int D, float* d_res, uint16_t* dy_hi, uint16_t* dy_lo, int ld_s, float* colsum, void* stream);
int navc_act_drop_bwd_split(const float* dout, const float* u, int act, uint64_t seed, float p, int M, int N,
                            uint16_t* dy_hi, uint16_t* dy_lo, int ld_s, float* colsum, void* stream);
/* x [M,N] (ld) fp32 -> any of: bf16 hi/lo copy [M, ld_s]; transposed fp32 / bf16 hi/lo [N, ld_t]
 * (columns M..ld_t-1 zero-filled); colsum[n] += sum_m x[m,n] (atomic; caller zero-fills).  Feeds
 * the gradient GEMMs dX = dY W and dW = dY^T X and the bias gradients. */
int navc_transpose_pack(const float* x, int M, int N, int ld, uint16_t* hi, uint16_t* lo, int ld_s,
                        float* t_f32, uint16_t* t_hi, uint16_t* t_lo, int ld_t, float* colsum,
                        void* stream);

/* Encoder, train mode (models/Encoder.py:19-25, 62-66): o = drop(seed,p)(highway(x, yg)) [BF, D]. */
int navc_highway_fwd_train(const float* x, const float* yg, int gate, int BF, int D, uint64_t seed,
                           float p, float* o, void* stream);
/* d_x (direct term) [BF,D] and d_yg [BF, gate?2D:D] from d_o. */
int navc_highway_bwd(const float* d_o, const float* x, const float* yg, int gate, int BF, int D,
                     uint64_t seed, float p, float* d_x, float* d_yg, void* stream);
/* BatchNorm1d batch statistics over the M rows of o [M,D] (joint_representation.py:43-45):
 * mean[d], var[d] (biased); if running_mean != NULL the running statistics are updated in place
 * with `momentum` and the unbiased variance, as nn.BatchNorm1d does in train mode. */
int navc_bn_stats(const float* o, int M, int D, float* mean, float* var, float momentum,
                  float* running_mean, float* running_var, void* stream);
/* enc_out[b, slot*F+f, :] = mean ? (o - mean)*rsqrt(var+eps)*w + b : o ; enc_hidden as navc_highway_bn. */
int navc_bn_apply_concat(const float* o, const float* mean, const float* var, const float* w,
                         const float* b, float eps, int B, int F, int D, int E, int slot, int n_modalities,
                         int accumulate, float* enc_hidden, float* enc_out, uint16_t* enc_hi,
                         uint16_t* enc_lo, void* stream);
/* Backward of the above for one modality slot: d_o [B*F, D]; d_w, d_b [D] (overwritten; NULL when
 * mean == NULL i.e. no norm); d_enc_hidden may be NULL. */
int navc_bn_bwd(const float* d_enc_out, const float* d_enc_hidden, const float* o, const float* mean,
                const float* var, const float* w, float eps, int B, int F, int D, int E, int slot,
                int n_modalities, float* d_w, float* d_b, float* d_o, void* stream);
/* d_enc_out[b,e,:] += d_mean[b,:] / E  (backward of enc_output.mean(1): Predictor.py:29,
 * Decoder.py:137). */
int navc_mean_bwd(const float* d_mean, int B, int E, int D, float* d_enc_out, void* stream);

/* log_softmax backward over rows: dlogits = g - exp(logp) * sum(g); columns V..ld_out-1 of the
 * output are zero-filled (seq2seq.py:102-103 under autograd). */
int navc_log_softmax_bwd(const float* g, const float* logp, int M, int V, int ld_in, float* dlogits,
                         int ld_out, void* stream);

/* LayerNorm backward (rows whose token is PAD were zeroed by the forward -> zero gradient):
 * dx [M,D]; dw[d] += , db[d] += (atomic; caller zero-fills). */
int navc_layernorm_bwd(const float* dy, const float* x, const float* w, float eps,
                       const int64_t* row_tokens, int M, int D, float* dx, float* dw, float* db,
                       void* stream);
/* Backward of navc_embed_ln given d(out) [N*S, D]: LayerNorm backward (d_ln_w, d_ln_b += ) and
 * scatter-add of the embedding-sum gradient into d_word[tok] (not for PAD: padding_idx),
 * d_pos[s], d_cat[category[n/group]], d_extra[n/group] (all atomic; caller zero-fills). */
int navc_embed_ln_bwd(const float* dout, const int64_t* tokens, const int64_t* category,
                      const float* word_emb, const float* pos_emb, const float* cat_emb,
                      const float* extra, int group, const float* ln_w, const float* ln_b, float eps,
                      int N, int S, int D, float* d_word, float* d_pos, float* d_cat, float* d_extra,
                      float* d_ln_w, float* d_ln_b, void* stream);

/* Packed-row variants for the training path (row counts are host-side there): the query rows of sequence n are
 * rows [seq_off[n], seq_off[n+1]) of qkv / q / d_ctx / d_q(kv); tokens stays the padded [N,S] tensor. */
int navc_self_attention_bwd_packed(const float* qkv, int ld, const int64_t* tokens, const int32_t* seq_off, int N,
                                   int S, int D, int H, int mask_kind, int watch, const float* d_ctx, float* d_qkv,
                                   void* stream);
int navc_cross_attention_bwd_packed(const float* q, int ldq, const float* kv, int ldkv, const int32_t* seq_off, int N,
                                    int S, int E, int D, int H, const float* d_ctx, float* d_q, int ld_dq, float* d_kv,
                                    int ld_dkv, void* stream);
/* The same two gradients on the tensor cores (csrc/attention_bwd_tc.cu; dk == 64, S <= 32, E <= 128, packed rows only): one
 * CTA per (owner, head), five tcgen05 products (S, dP, dV, dK, dQ) on bf16 (mode NAVC_TC_BF16) or split-bf16
 * (NAVC_TC_BF16X3) operands made in-kernel from the fp32 activations; softmax statistics, P and dS in fp32.
 * Same arguments / results as navc_self_attention_bwd_packed / navc_cross_attention_bwd_packed (self: the PAD mask is
 * implied by the packing, so no token ids are read), plus `ctx` [rows, D] fp32, the forward pass's context (or NULL):
 * with it the softmax-gradient row term sum_j P_ij dP_ij is taken as dO_i . ctx_i, which saves a pass over dP. */
int navc_self_attention_bwd_tc(int mode, const float* qkv, int ld, const int32_t* seq_off, int N, int S, int D, int H,
                               int mask_kind, int watch, const float* d_ctx, const float* ctx, float* d_qkv, void* stream);
int navc_cross_attention_bwd_tc(int mode, const float* q, int ldq, const float* kv, int ldkv, const int32_t* seq_off,
                                int N, int S, int E, int D, int H, const float* d_ctx, const float* ctx, float* d_q,
                                int ld_dq, float* d_kv, int ld_dkv, void* stream);
int navc_embed_ln_bwd_packed(const float* dout, const int64_t* tokens, const int64_t* category, const float* word_emb,
                             const float* pos_emb, const float* cat_emb, const float* extra, int group,
                             const float* ln_w, float eps, int S, int D, const int32_t* rowmap, int rows,
                             float* d_word, float* d_pos, float* d_cat, float* d_extra, float* d_ln_w, float* d_ln_b,
                             void* stream);
/* log_softmax of packed logits rows written to padded rows (out row = rowmap[i]) and its backward reading the
 * padded g / logp rows; fp32 row gather (scatter = 0: out[i] = in[rowmap[i]]) / scatter (out[rowmap[i]] = in[i]). */
int navc_log_softmax_rows(const float* logits, int ld_in, float* out, int ld_out, const int32_t* rowmap, int rows,
                          int V, void* stream);
int navc_log_softmax_bwd_rows(const float* g, const float* logp, const int32_t* rowmap, int rows, int V, int ld_in,
                              float* dlogits, int ld_out, void* stream);
/* Bias gradient of the rows a packed vocabulary projection skipped (PAD positions: hidden == 0, log-probs ==
 * const_logp = log_softmax(bias)): db[v] += sum over pad rows r of g[r,v] - exp(const_logp[v]) * sum_v g[r,:].
 * All-zero rows of g (a loss that ignores PAD) cost one read. */
int navc_log_softmax_bwd_padrows(const float* g, int ld, const int32_t* pad_rows, int n_pad, const float* const_logp,
                                 int V, float* db, void* stream);
int navc_rows_f32(const float* in, float* out, int D, const int32_t* rowmap, int rows, int scatter, void* stream);

/* Fused cross-entropy (projection + log-softmax + masked NLL, seq2seq.py:102-103 + misc/crit.py:62-84)
 * without a [rows, V] log-prob tensor.  Forward: navc_vocab_partials_* with target = labels, then
 * navc_ce_stats: lse[r], nll[r] = lse - logit[label] (0 where label == PAD), argmax[r].  Backward, per
 * row chunk: logits chunk (GEMM) -> navc_ce_grad in place: (softmax - onehot) * scale[0] * (label != PAD),
 * pad columns zeroed -> the usual gradient GEMMs.  `scale` is a device scalar (upstream gradient). */
int navc_ce_stats(const float* part_max, const float* part_sum, const int32_t* part_idx, int n_tiles,
                  const float* target_logit, const int64_t* labels, int R, float* lse, float* nll,
                  int32_t* argmax, void* stream);
int navc_ce_grad(float* logits, const float* lse, const int64_t* labels, const float* scale, int rows,
                 int V, int ld, void* stream);

/* Fused clip_grad_value_ + Adam with L2 weight decay over flat fp32 buffers (misc/run.py:260-261,
 * misc/optim.py:61-62; torch.optim.Adam semantics, bias correction by `step` >= 1):
 *   g = clamp(g, -clip, clip) (clip <= 0: off); g += wd*p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).   One launch per optimizer step. */
int navc_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float clip, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

/* Attention backward (softmax recomputed from Q, K; masks as the forward).  Self: d_qkv [N*S, ld]
 * receives dQ | dK | dV at column offsets 0, D, 2D.  Cross: d_q [N*S, ld_dq]; d_kv [(N/group)*E,
 * ld_dkv] receives dK | dV at 0, D summed over the `group` rows of each video. */
int navc_self_attention_bwd(const float* qkv, int ld, const int64_t* tokens, int N, int S, int D, int H,
                            int mask_kind, int watch, const float* d_ctx, float* d_qkv, void* stream);
int navc_cross_attention_bwd(const float* q, int ldq, const float* kv, int ldkv, int N, int S, int E,
                             int D, int H, int group, const float* d_ctx, float* d_q, int ld_dq,
                             float* d_kv, int ld_dkv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NAVC_H */
