// Autoregressive beam search on the device (reference models/Translator.py:94-161 + models/Beam.py):
// one new position per beam row and step, self-attention over a K/V cache, beam bookkeeping in a kernel.
//
//  * K/V cache: per layer kc / vc [T, N, D] fp32, written at step position p by the row that computed it.
//    Beams are re-ordered every step; instead of moving the cache, row r keeps its ANCESTRY anc[r, j] = the
//    row that wrote position j of r's prefix (a [N, T] int32 table, re-gathered per step: 4 bytes per entry
//    instead of 2 * L * D * 4).
//  * the token history hist [N, T] (int64, hist[r, 0] = BOS) gives the key padding mask (a PAD chosen as a
//    word masks that key, models/Decoder.py:26-39) and the finished hypotheses.
#include "common.cuh"

namespace navc {

constexpr float kMaskFillStep = -10e6f;  // models/bert.py:161

// grid N, block 32 * H: warp h of block r = head h of beam row r.  Query = the row's new position p.
template <int DK>
__global__ void self_attention_step_kernel(const float* __restrict__ qkv, int ld, float* __restrict__ kc, float* __restrict__ vc,
                                           const int32_t* __restrict__ anc, const int64_t* __restrict__ hist, int N, int T,
                                           int D, int p, int watch, float* __restrict__ ctx_f32, uint16_t* __restrict__ ctx_hi,
                                           uint16_t* __restrict__ ctx_lo) {
    constexpr int PER = (DK + 31) / 32;  // head columns per lane
    extern __shared__ float sm[];
    const int r = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* qs = sm + (size_t)h * (DK + 64);  // [DK] query, then [<= 64] probabilities
    float* ps = qs + DK;
    const float* row = qkv + (size_t)r * ld + h * DK;
    // append this row's key / value to the cache, stage the query
    const size_t slot = ((size_t)p * N + r) * D + h * DK;
#pragma unroll
    for (int c = 0; c < PER; ++c) {
        const int d = lane + 32 * c;
        if (d < DK) {
            qs[d] = row[d];
            kc[slot + d] = row[D + d];
            vc[slot + d] = row[2 * D + d];
        }
    }
    __syncwarp();
    const float inv_sqrt = 1.0f / sqrtf((float)DK);
    const int nk = p + 1;
    const int32_t* arow = anc + (size_t)r * T;
    const int64_t* hrow = hist + (size_t)r * T;
    // scores: lane j (and j + 32) owns key j
    float sc[2];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        sc[t] = -INFINITY;
        if (j < nk) {
            const float* kr = (j == p) ? row + D : kc + ((size_t)j * N + arow[j]) * D + h * DK;
            float s = 0.f;
#pragma unroll 4
            for (int d = 0; d < DK; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(kr + d);
                s = fmaf(qs[d], kk.x, s); s = fmaf(qs[d + 1], kk.y, s);
                s = fmaf(qs[d + 2], kk.z, s); s = fmaf(qs[d + 3], kk.w, s);
            }
            s *= inv_sqrt;
            const bool masked = hrow[j] == NAVC_PAD || (watch != 0 && j <= p - watch);
            sc[t] = masked ? kMaskFillStep : s;
            m = fmaxf(m, sc[t]);
        }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        if (j < nk) {
            sc[t] = expf(sc[t] - m);
            sum += sc[t];
        }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        if (j < nk) ps[j] = sc[t] * inv;
    }
    __syncwarp();
    // context: lane owns columns lane + 32 c
    float acc[PER];
#pragma unroll
    for (int c = 0; c < PER; ++c) acc[c] = 0.f;
    for (int j = 0; j < nk; ++j) {
        const float pj = ps[j];
        const float* vr = (j == p) ? row + 2 * D : vc + ((size_t)j * N + arow[j]) * D + h * DK;
#pragma unroll
        for (int c = 0; c < PER; ++c)
            if (lane + 32 * c < DK) acc[c] = fmaf(pj, vr[lane + 32 * c], acc[c]);
    }
    const size_t o = (size_t)r * D + h * DK;
#pragma unroll
    for (int c = 0; c < PER; ++c) {
        const int d = lane + 32 * c;
        if (d >= DK) continue;
        if (ctx_f32) ctx_f32[o + d] = acc[c];
        if (ctx_hi) {
            uint16_t hi, lo;
            split_bf16(acc[c], hi, lo);
            ctx_hi[o + d] = hi;
            if (ctx_lo) ctx_lo[o + d] = lo;
        }
    }
}

// Beam.advance (models/Beam.py:68-117) for every video: one block per video, thread 0 does the (K-long) sequential
// part, all threads copy the history / ancestry rows.  best_* come from a top-K over beam x vocab of
// (beam score + log-prob), EOS-terminated beams filled with -1e20 (Beam.py:71-74).
__global__ void beam_advance_kernel(const float* __restrict__ best_scores, const int64_t* __restrict__ best_ids, int B, int K,
                                    int V, int t, int max_len, int want, int T, const int64_t* __restrict__ hist_in,
                                    int64_t* __restrict__ hist_out, const int32_t* __restrict__ anc_in,
                                    int32_t* __restrict__ anc_out, float* __restrict__ scores, int32_t* __restrict__ done,
                                    int32_t* __restrict__ fin_count, float* __restrict__ fin_score, int32_t* __restrict__ fin_len,
                                    int64_t* __restrict__ fin_tok, int cap, int32_t* __restrict__ n_done) {
    const int b = blockIdx.x;
    const int p = t - 1;  // position whose K/V the step just cached
    const bool was_done = done[b] != 0;
    for (int i = 0; i < K; ++i) {
        const int r = b * K + i;
        int parent = i;
        int64_t tok = 0;
        if (!was_done) {
            const int64_t id = best_ids[(size_t)b * K + i];
            parent = (int)(id / V);
            tok = id - (int64_t)parent * V;
        }
        const int pr = b * K + parent;
        for (int j = threadIdx.x; j < T; j += blockDim.x) {
            int64_t hv = hist_in[(size_t)pr * T + j];
            int32_t av = anc_in[(size_t)pr * T + j];
            if (!was_done) {
                if (j == t) hv = tok;
                if (j == p) av = pr;  // position p of the child's prefix was cached by the parent row
            }
            hist_out[(size_t)r * T + j] = hv;
            anc_out[(size_t)r * T + j] = av;
        }
        if (threadIdx.x == 0 && !was_done) scores[(size_t)b * K + i] = best_scores[(size_t)b * K + i];
    }
    __syncthreads();
    if (threadIdx.x != 0 || was_done) return;
    int cnt = fin_count[b];
    bool now_done = false;
    auto append = [&](int i) {
        if (cnt >= cap) return;
        const size_t e = (size_t)b * cap + cnt;
        fin_score[e] = best_scores[(size_t)b * K + i];
        fin_len[e] = t;
        for (int j = 0; j < t; ++j) fin_tok[e * T + j] = hist_out[(size_t)(b * K + i) * T + j + 1];
        ++cnt;
    };
    for (int i = 0; i < K && !now_done; ++i) {
        if (hist_out[(size_t)(b * K + i) * T + t] == NAVC_EOS) {
            append(i);
            if (cnt >= want) now_done = true;
        }
    }
    if (!now_done && t + 1 == max_len) {
        now_done = true;
        if (cnt == 0)
            for (int i = 0; i < K; ++i) append(i);
    }
    fin_count[b] = cnt;
    if (now_done) {
        done[b] = 1;
        atomicAdd(n_done, 1);
    }
}

// Top-K over beam x vocab of (beam score + log_softmax(logits)) for every video, straight from the logits:
// replaces log_softmax + add + masked_fill + topk (Translator.py:113-114, Beam.py:68-83).  One block per video.
// Beams whose last token is EOS contribute -1e20 (Beam.py:71-74); at the first step only beam 0 counts
// (Beam.py:75-76).  Ties: lowest flat index first.
constexpr int kBeamMaxK = 8;
constexpr int kTopkThreads = 1024;  // one block per video: the whole SM works on its beam x vocab candidates

__global__ void __launch_bounds__(kTopkThreads)
beam_topk_kernel(const float* __restrict__ logits, int ld, int V, int K, const float* __restrict__ scores,
                 const int64_t* __restrict__ hist, int T, int pos, int first, float* __restrict__ best_scores,
                 int64_t* __restrict__ best_ids) {
    __shared__ float red[kTopkThreads / 32];
    __shared__ float s_lse[kBeamMaxK];
    __shared__ float s_val[kTopkThreads];
    __shared__ int s_idx[kTopkThreads];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = first ? 1 : K;
    // log-sum-exp of every beam row
    for (int k = 0; k < rows; ++k) {
        const float* lr = logits + (size_t)(b * K + k) * ld;
        float m = -INFINITY;
        for (int v = tid; v < V; v += kTopkThreads) m = fmaxf(m, lr[v]);
        m = warp_max(m);
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = red[0];
        for (int w = 1; w < kTopkThreads / 32; ++w) m = fmaxf(m, red[w]);
        __syncthreads();
        float sum = 0.f;
        for (int v = tid; v < V; v += kTopkThreads) sum += expf(lr[v] - m);
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < kTopkThreads / 32; ++w) t += red[w];
            s_lse[k] = m + logf(t);
        }
        __syncthreads();
    }
    // per-thread sorted top-K of its candidates (value desc, index asc)
    float lv[kBeamMaxK];
    int li[kBeamMaxK];
#pragma unroll
    for (int i = 0; i < kBeamMaxK; ++i) { lv[i] = -INFINITY; li[i] = 0x7fffffff; }
    for (int k = 0; k < rows; ++k) {
        const float* lr = logits + (size_t)(b * K + k) * ld;
        const bool dead = !first && hist[(size_t)(b * K + k) * T + pos] == NAVC_EOS;
        const float base = first ? 0.f : scores[(size_t)b * K + k];
        const float lse = s_lse[k];
        for (int v = tid; v < V; v += kTopkThreads) {
            const float val = dead ? -1e20f : (lr[v] - lse) + base;   // log_softmax first, then + score (torch's order)
            const int idx = k * V + v;
            if (val > lv[K - 1] || (val == lv[K - 1] && idx < li[K - 1])) {
                int j = K - 1;
#pragma unroll
                for (int q = kBeamMaxK - 1; q > 0; --q) {
                    if (q <= j && (val > lv[q - 1] || (val == lv[q - 1] && idx < li[q - 1]))) {
                        lv[q] = lv[q - 1]; li[q] = li[q - 1];
                        j = q - 1;
                    }
                }
#pragma unroll
                for (int q = 0; q < kBeamMaxK; ++q)
                    if (q == j) { lv[q] = val; li[q] = idx; }   // static indices: the lists stay in registers
            }
        }
    }
    // K rounds of block arg-max over the heads of the per-thread lists
    int head = 0;
    for (int r = 0; r < K; ++r) {
        float hv = -INFINITY;
        int hi = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < kBeamMaxK; ++q)
            if (q == head) { hv = lv[q]; hi = li[q]; }
        s_val[tid] = hv;
        s_idx[tid] = hi;
        __syncthreads();
        for (int o = kTopkThreads / 2; o > 0; o >>= 1) {
            if (tid < o) {
                const float ov = s_val[tid + o];
                const int oi = s_idx[tid + o];
                if (ov > s_val[tid] || (ov == s_val[tid] && oi < s_idx[tid])) { s_val[tid] = ov; s_idx[tid] = oi; }
            }
            __syncthreads();
        }
        const int win = s_idx[0];
        if (tid == 0) {
            best_scores[(size_t)b * K + r] = s_val[0];
            best_ids[(size_t)b * K + r] = (int64_t)win;
        }
        if (hi == win && head < K) ++head;  // flat indices are unique: exactly one thread advances
        __syncthreads();
    }
}

}  // namespace navc

using namespace navc;

extern "C" int navc_beam_topk(const float* logits, int ld, int B, int K, int V, const float* scores, const int64_t* hist,
                              int T, int pos, int first, float* best_scores, int64_t* best_ids, void* stream) {
    NAVC_REQUIRE(logits && scores && hist && best_scores && best_ids, "navc_beam_topk: null pointer");
    NAVC_REQUIRE(B > 0 && K >= 1 && K <= kBeamMaxK && V >= K && ld >= V && pos >= 0 && pos < T,
                 "navc_beam_topk: bad arguments (beam size <= %d)", kBeamMaxK);
    beam_topk_kernel<<<B, kTopkThreads, 0, as_stream(stream)>>>(logits, ld, V, K, scores, hist, T, pos, first, best_scores, best_ids);
    return check_launch("navc_beam_topk");
}

extern "C" int navc_self_attention_step(const float* qkv, int ld, float* k_cache, float* v_cache, const int32_t* anc,
                                        const int64_t* hist, int N, int T, int D, int H, int pos, int watch,
                                        float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream) {
    NAVC_REQUIRE(qkv && k_cache && v_cache && anc && hist && (ctx_f32 || ctx_hi), "navc_self_attention_step: null pointer");
    NAVC_REQUIRE(N > 0 && H > 0 && H <= 32 && D % H == 0 && ld >= 3 * D && ld % 4 == 0 && D % 4 == 0 && pos >= 0 && pos < T && T <= 64,
                 "navc_self_attention_step: bad shape (N=%d T=%d D=%d H=%d pos=%d; T <= 64)", N, T, D, H, pos);
    NAVC_REQUIRE(((((uintptr_t)qkv) | ((uintptr_t)k_cache) | ((uintptr_t)v_cache)) & 15) == 0,
                 "navc_self_attention_step: operands must be 16-byte aligned");
    const int dk = D / H;
    const size_t smem = (size_t)H * (dk + 64) * sizeof(float);
    cudaStream_t st = as_stream(stream);
#define NAVC_STEP(DKV)                                                                                                        \
    if (dk == DKV) {                                                                                                          \
        self_attention_step_kernel<DKV><<<N, 32 * H, smem, st>>>(qkv, ld, k_cache, v_cache, anc, hist, N, T, D, pos, watch,   \
                                                                 ctx_f32, ctx_hi, ctx_lo);                                   \
        return check_launch("navc_self_attention_step");                                                                      \
    }
    NAVC_STEP(64)
    NAVC_STEP(32)
    NAVC_STEP(16)
    NAVC_STEP(128)
#undef NAVC_STEP
    NAVC_REQUIRE(false, "navc_self_attention_step: head size %d unsupported (16, 32, 64, 128)", dk);
    return 1;
}

extern "C" int navc_beam_advance(const float* best_scores, const int64_t* best_ids, int B, int K, int V, int t, int max_len,
                                 int want, int T, const int64_t* hist_in, int64_t* hist_out, const int32_t* anc_in,
                                 int32_t* anc_out, float* scores, int32_t* done, int32_t* fin_count, float* fin_score,
                                 int32_t* fin_len, int64_t* fin_tok, int cap, int32_t* n_done, void* stream) {
    NAVC_REQUIRE(best_scores && best_ids && hist_in && hist_out && anc_in && anc_out && scores && done && fin_count &&
                     fin_score && fin_len && fin_tok && n_done,
                 "navc_beam_advance: null pointer");
    NAVC_REQUIRE(B > 0 && K > 0 && V > 0 && t >= 1 && t < T && T >= max_len && cap >= 1 && want >= 1 && hist_in != hist_out &&
                     anc_in != anc_out,
                 "navc_beam_advance: bad arguments");
    beam_advance_kernel<<<B, 64, 0, as_stream(stream)>>>(best_scores, best_ids, B, K, V, t, max_len, want, T, hist_in, hist_out,
                                                        anc_in, anc_out, scores, done, fin_count, fin_score, fin_len, fin_tok,
                                                        cap, n_done);
    return check_launch("navc_beam_advance");
}
