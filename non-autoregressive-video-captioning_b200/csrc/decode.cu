// Iterative-refinement decode kernels: length beam, canvas init, the fused per-iteration step
// (combine vocabulary partials -> argmax/prob -> pad rules -> merge -> re-mask selection -> next
// canvas), teacher probabilities and candidate selection.  All tiny, latency-bound, int/float work.
#include "common.cuh"

namespace navc {

// one warp per video: top-`lbs` of max_len values (descending, lowest index first on ties)
__global__ void length_beam_kernel(const float* __restrict__ pred, int B, int max_len, int lbs, int bias,
                                   int32_t* __restrict__ beam, int32_t* __restrict__ smax) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= B) return;
    float* vals = sm + (size_t)warp * max_len;
    for (int i = lane; i < max_len; i += 32) vals[i] = pred[(size_t)b * max_len + i];
    __syncwarp();
    int local_max = 0;
    for (int r = 0; r < lbs; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = lane; i < max_len; i += 32) {
            float v = vals[i];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        int len = bi + bias;
        len = len < 4 ? 4 : len;
        len = len > max_len - 1 ? max_len - 1 : len;
        if (lane == 0) {
            beam[(size_t)b * lbs + r] = len;
            vals[bi] = -INFINITY;  // NaN-free inputs assumed (log-probabilities)
        }
        local_max = max(local_max, len);
        __syncwarp();
    }
    if (lane == 0) atomicMax(smax, local_max);
}

__global__ void init_canvas_kernel(const int32_t* __restrict__ beam, int N, int S, int64_t fill,
                                   int64_t* canvas, int64_t* tokens, float* probs) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    int n = idx / S, s = idx - n * S;
    bool inside = s < beam[n];
    int64_t t = inside ? fill : (int64_t)NAVC_PAD;
    if (canvas) canvas[idx] = t;
    if (tokens) tokens[idx] = t;
    if (probs) probs[idx] = inside ? 0.f : 1.f;
}

struct StepArgs {
    navc_step_t p;
};

// one block per candidate row; thread i owns position i.
__global__ void refine_step_kernel(StepArgs a, int S) {
    extern __shared__ float sm[];
    float* key = sm;                                   // [S] ranking keys
    int* flag = reinterpret_cast<int*>(sm + S);        // [S] 0/1 flags
    const navc_step_t& p = a.p;
    const int n = blockIdx.x, i = threadIdx.x;
    const bool active = i < S;
    const size_t o = (size_t)n * S + (active ? i : 0);
    const int len = p.lens[n];
    const bool pad = active && i >= len;

    int64_t tok = active ? p.tokens[o] : (int64_t)NAVC_PAD;
    float prob = active ? p.probs[o] : 1.f;

    if (p.merge != NAVC_MERGE_NONE) {
        // phase A: the warps of the block combine the vocabulary partials of the row's positions (lane t takes
        // tiles t, t+32, ...; coalesced loads, shuffle reduction) into shared memory
        int* s_tok = flag + S;
        float* s_prob = reinterpret_cast<float*>(s_tok + S);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int pos = warp; pos < S; pos += nwarps) {
            const bool pad_p = pos >= len;
            const size_t op = (size_t)n * S + pos;
            // packed rows: the partials of (n, pos < len) live in row seq_off[n] + pos; PAD positions have none
            size_t prow = p.seq_off ? (size_t)(p.seq_off[n] + pos) : op;
            bool have = !(pad_p && p.seq_off);
            // second-level packing: only the previously re-masked positions have partials (the merge below
            // ignores the others)
            if (have && p.part_slot) {
                have = p.upd_mask[op] != 0;
                if (have) prow = (size_t)p.part_slot[prow];
            }
            SoftPart acc;
            acc.m = -INFINITY; acc.s = 0.f; acc.i = 0x7fffffff;
            if (have) {
                const size_t pr = prow * p.n_tiles;
                for (int t = lane; t < p.n_tiles; t += 32) {
                    SoftPart q;
                    q.m = p.part_max[pr + t]; q.s = p.part_sum[pr + t]; q.i = p.part_idx[pr + t];
                    acc = soft_combine(acc, q);
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) {
                    SoftPart q;
                    q.m = __shfl_xor_sync(0xffffffffu, acc.m, sh);
                    q.s = __shfl_xor_sync(0xffffffffu, acc.s, sh);
                    q.i = __shfl_xor_sync(0xffffffffu, acc.i, sh);
                    acc = soft_combine(acc, q);
                }
            }
            if (lane == 0) {
                s_tok[pos] = (have && !pad_p) ? acc.i : NAVC_PAD;
                s_prob[pos] = (have && !pad_p) ? 1.0f / acc.s : 1.0f;
            }
        }
        __syncthreads();
        int64_t ntok = NAVC_PAD;
        float nprob = 1.f;
        if (active) {
            ntok = s_tok[i];
            nprob = s_prob[i];
            if (p.is_ct && ntok == NAVC_MASK) nprob = 0.0f;
        }
        if (p.merge == NAVC_MERGE_ALL) {
            tok = ntok; prob = nprob;
        } else if (p.merge == NAVC_MERGE_MASKED) {
            if (active && p.upd_mask[o]) { tok = ntok; prob = nprob; }
        } else {  // NAVC_MERGE_EF
            const bool is_masked = active && tok == NAVC_MASK;
            const float cand = is_masked ? nprob : 0.f;
            if (active) key[i] = cand;
            const int remaining = __syncthreads_count(is_masked);
            const int k = remaining < p.q ? remaining : p.q;
            if (active) {
                int rank = 0;
                for (int j = 0; j < S; ++j) {
                    float kj = key[j];
                    rank += (kj > cand) || (kj == cand && j < i);
                }
                if (rank < k) { tok = ntok; prob = cand; }
            }
            __syncthreads();
        }
    }

    const int left = __syncthreads_count(active && tok == NAVC_MASK);
    if (active) {
        if (p.visual) p.visual[o] = (tok != NAVC_MASK && tok != NAVC_PAD) ? 1 : 0;
        if (p.masked0) p.masked0[o] = (tok == NAVC_MASK && i < len) ? 1 : 0;
    }

    const float tprob = (active && p.teacher) ? p.teacher[o] : 1.f;
    bool sel = false;
    if (p.select == NAVC_SELECT_WORST) {
        int k = (int)((float)len * p.ratio);
        k = k < 1 ? 1 : k;
        const float mykey = prob * tprob;
        if (active) key[i] = mykey;
        __syncthreads();
        if (active) {
            int rank = 0;
            for (int j = 0; j < S; ++j) {
                float kj = key[j];
                rank += (kj < mykey) || (kj == mykey && j < i);
            }
            sel = rank < k;
        }
    } else if (p.select == NAVC_SELECT_MASKTOK) {
        sel = active && tok == NAVC_MASK;
    } else if (p.select == NAVC_SELECT_GIVEN) {
        sel = active && p.given[o] != 0;
    } else if (p.select == NAVC_SELECT_WINDOW) {
        const int g = (active && p.given[o] != 0) ? 1 : 0;
        if (active) flag[i] = g;
        __syncthreads();
        if (g) {
            int ord = 0;
            for (int j = 0; j < i; ++j) ord += flag[j];
            sel = ord >= p.win_lo && ord < p.win_hi;
        }
    }
    const int nsel = __syncthreads_count(sel);
    if (active) {
        p.tokens[o] = tok;
        p.probs[o] = prob;
        if (p.upd_mask) p.upd_mask[o] = sel ? 1 : 0;
        if (p.canvas) p.canvas[o] = sel ? (int64_t)NAVC_MASK : tok;
        if (sel && p.sel_rows && !pad) {
            const int prow = p.seq_off[n] + i;
            const int k = atomicAdd(p.sel_count, 1);
            p.sel_rows[k] = prow;
            p.sel_slot[prow] = k;
        } else if (!p.sel_rows && p.sel_slot && !pad) {
            p.sel_slot[p.seq_off[n] + i] = sel ? 1 : 0;   // flags only: navc_compact_rows orders them
        }
        if (p.select == NAVC_SELECT_NONE && p.lprobs) p.lprobs[o] = logf(prob * tprob);
    }
    if (i == 0 && p.counters) {
        if (left) atomicAdd(p.counters + 0, left);
        if (nsel) atomicAdd(p.counters + 1, nsel);
    }
}

__global__ void teacher_probs_kernel(const float* __restrict__ pm, const float* __restrict__ ps, int n_tiles,
                                     const float* __restrict__ tl, const int32_t* __restrict__ lens, int N, int S,
                                     float* __restrict__ out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    int n = idx / S, s = idx - n * S;
    if (s >= lens[n]) { out[idx] = 1.0f; return; }
    SoftPart acc;
    acc.m = -INFINITY; acc.s = 0.f; acc.i = 0;
    for (int t = 0; t < n_tiles; ++t) {
        SoftPart q;
        q.m = pm[(size_t)idx * n_tiles + t]; q.s = ps[(size_t)idx * n_tiles + t]; q.i = 0;
        acc = soft_combine(acc, q);
    }
    out[idx] = expf(tl[idx] - acc.m) / acc.s;
}

__global__ void teacher_inputs_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ map, int N,
                                      int S, int64_t* __restrict__ shifted, int64_t* __restrict__ mapped) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S) return;
    int s = idx % S;
    int64_t t = tokens[idx];
    if (map) t = map[t];
    mapped[idx] = t;
    if (s == 0) shifted[idx] = NAVC_BOS;
    if (s + 1 < S) shifted[idx + 1] = t;
}

// one warp per video
__global__ void select_best_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ lprobs,
                                   const int32_t* __restrict__ lens, int B, int lbs, int S, float alpha,
                                   int64_t* __restrict__ hyp, float* __restrict__ score) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= B) return;
    float best = -INFINITY;
    int bj = 0;
    for (int j = 0; j < lbs; ++j) {
        const size_t row = (size_t)b * lbs + j;
        // sequential sum in position order (matches a left-to-right fp32 reduction closely; the
        // comparison between candidates only needs to be stable, ties resolve to the lowest j)
        float s = 0.f;
        for (int t = lane; t < S; t += 32) s += lprobs[row * S + t];
        s = warp_sum(s);
        float sc = s / powf((float)lens[row], alpha);
        if (score && lane == 0) score[row] = sc;
        if (sc > best) { best = sc; bj = j; }
    }
    const size_t row = (size_t)b * lbs + bj;
    for (int t = lane; t < S; t += 32) hyp[(size_t)b * S + t] = tokens[row * S + t];
}

}  // namespace navc

using namespace navc;

extern "C" int navc_length_beam(const float* pred_length, int B, int max_len, int lbs, int length_bias,
                                int32_t* beam, int32_t* smax, void* stream) {
    NAVC_REQUIRE(pred_length && beam && smax, "navc_length_beam: null pointer");
    NAVC_REQUIRE(B > 0 && max_len > 4 && lbs > 0 && lbs <= max_len, "navc_length_beam: bad shape");
    const int wpb = 4;
    size_t smem = (size_t)wpb * max_len * sizeof(float);
    length_beam_kernel<<<(B + wpb - 1) / wpb, wpb * 32, smem, as_stream(stream)>>>(pred_length, B, max_len, lbs,
                                                                                  length_bias, beam, smax);
    return check_launch("navc_length_beam");
}

extern "C" int navc_init_canvas(const int32_t* beam, int N, int S, int64_t fill, int64_t* canvas, int64_t* tokens,
                                float* probs, void* stream) {
    NAVC_REQUIRE(beam && N > 0 && S > 0, "navc_init_canvas: bad arguments");
    init_canvas_kernel<<<(N * S + 255) / 256, 256, 0, as_stream(stream)>>>(beam, N, S, fill, canvas, tokens, probs);
    return check_launch("navc_init_canvas");
}

extern "C" int navc_refine_step(const navc_step_t* p, int N, int S, void* stream) {
    NAVC_REQUIRE(p && p->lens && p->tokens && p->probs, "navc_refine_step: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 1024, "navc_refine_step: bad shape");
    NAVC_REQUIRE(p->merge == NAVC_MERGE_NONE || (p->part_max && p->part_sum && p->part_idx && p->n_tiles > 0),
                 "navc_refine_step: merge requested without partials");
    NAVC_REQUIRE(p->merge != NAVC_MERGE_MASKED || p->upd_mask, "navc_refine_step: MERGE_MASKED needs upd_mask");
    NAVC_REQUIRE(!p->part_slot || (p->seq_off && p->merge == NAVC_MERGE_MASKED), "navc_refine_step: part_slot needs seq_off and MERGE_MASKED");
    NAVC_REQUIRE(!p->sel_rows || (p->seq_off && p->sel_count && p->sel_slot), "navc_refine_step: sel_rows needs seq_off, sel_count, sel_slot");
    NAVC_REQUIRE(!p->sel_slot || p->seq_off, "navc_refine_step: sel_slot needs seq_off");
    NAVC_REQUIRE((p->select != NAVC_SELECT_GIVEN && p->select != NAVC_SELECT_WINDOW) || p->given,
                 "navc_refine_step: selection needs `given`");
    StepArgs a;
    a.p = *p;
    int threads = ((S + 31) / 32) * 32;
    if (threads < 256) threads = 256;  // extra warps only help phase A (combining the vocabulary partials)
    size_t smem = (size_t)S * 2 * (sizeof(float) + sizeof(int));
    refine_step_kernel<<<N, threads, smem, as_stream(stream)>>>(a, S);
    return check_launch("navc_refine_step");
}

// Ordered compaction of the selected packed rows (flags in `slot`, written by navc_refine_step without sel_rows): one block,
// thread t owns a contiguous chunk of rows.  slot[r] becomes the number of selected rows before r (for a selected row: its
// index in `rows`), rows[] the selected rows in ascending order, seq_off_c[n] = slot[seq_off[n]] the packed offsets of the
// compacted row space (a sequence's / video's selected rows stay contiguous), count = seq_off_c[N].
__global__ void __launch_bounds__(1024) compact_rows_kernel(int32_t* __restrict__ slot, const int32_t* __restrict__ seq_off, int N, int max_rows,
                                    int32_t* __restrict__ rows, int32_t* __restrict__ count, int32_t* __restrict__ seq_off_c) {
    // warp w owns a contiguous segment of 32-row groups; lane = row inside a group (coalesced, independent loads; the first
    // version walked 21 dependent loads per thread: 12 us).  Flags of a warp's groups stay in registers between the passes.
    constexpr int kMaxGroups = 32;                 // groups per warp: 32 warps x 32 groups x 32 rows = 32768 rows
    __shared__ int wsum[32];
    __shared__ int total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int R = seq_off[N];
    if (R > max_rows) R = max_rows;
    const int groups = (max_rows + 31) >> 5;
    const int gpw = (groups + 31) >> 5;            // groups per warp (<= kMaxGroups, checked by the host)
    const int g0 = warp * gpw;
    uint32_t bits[kMaxGroups];
    int c = 0;
#pragma unroll
    for (int k = 0; k < kMaxGroups; ++k) {
        bits[k] = 0u;
        if (k < gpw) {
            const int r = (g0 + k) * 32 + lane;
            const int f = (r < R) ? (slot[r] != 0) : 0;
            bits[k] = __ballot_sync(0xffffffffu, f);
            c += __popc(bits[k]);                  // (warp-uniform)
        }
    }
    if (lane == 0) wsum[warp] = c;
    __syncthreads();
    if (warp == 0) {
        const int w = wsum[lane];
        int wi = w;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, sh);
            if (lane >= sh) wi += v;
        }
        wsum[lane] = wi - w;                       // exclusive prefix of the warp totals
        if (lane == 31) total_s = wi;
    }
    __syncthreads();
    int base = wsum[warp];
#pragma unroll
    for (int k = 0; k < kMaxGroups; ++k) {
        if (k < gpw) {
            const int r = (g0 + k) * 32 + lane;
            const uint32_t b = bits[k];
            const int pre = base + __popc(b & ((1u << lane) - 1u));
            if (r < R) {
                slot[r] = pre;
                if ((b >> lane) & 1u) rows[pre] = r;
            }
            base += __popc(b);
        }
    }
    const int total = total_s;
    if (tid == 0) { count[0] = total; slot[R] = total; }
    __syncthreads();
    for (int n = tid; n <= N; n += blockDim.x) {
        const int r = seq_off[n];
        seq_off_c[n] = r >= R ? total : slot[r];
    }
}

// any row count: thread t walks a contiguous chunk of rows (used beyond 32768 rows)
__global__ void compact_rows_serial_kernel(int32_t* __restrict__ slot, const int32_t* __restrict__ seq_off, int N, int max_rows,
                                    int32_t* __restrict__ rows, int32_t* __restrict__ count, int32_t* __restrict__ seq_off_c) {
    __shared__ int wsum[32];
    __shared__ int total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int R = seq_off[N];
    if (R > max_rows) R = max_rows;
    const int per = (max_rows + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = tid * per < R ? tid * per : R;
    const int hi = lo + per < R ? lo + per : R;
    int c = 0;
    for (int r = lo; r < hi; ++r) c += slot[r] != 0;
    int incl = c;
#pragma unroll
    for (int sh = 1; sh < 32; sh <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, sh);
        if (lane >= sh) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
        int wi = w;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, sh);
            if (lane >= sh) wi += v;
        }
        wsum[lane] = wi - w;                      // exclusive prefix of the warp totals
        if (lane == 31) total_s = wi;
    }
    __syncthreads();
    int base = wsum[warp] + incl - c;
    for (int r = lo; r < hi; ++r) {
        const int f = slot[r] != 0;
        slot[r] = base;
        if (f) rows[base] = r;
        base += f;
    }
    const int total = total_s;
    if (tid == 0) { count[0] = total; slot[R] = total; }
    __syncthreads();
    for (int n = tid; n <= N; n += blockDim.x) {
        const int r = seq_off[n];
        seq_off_c[n] = r >= R ? total : slot[r];
    }
}

extern "C" int navc_compact_rows(int32_t* slot, const int32_t* seq_off, int N, int max_rows, int32_t* rows, int32_t* count,
                                 int32_t* seq_off_c, void* stream) {
    NAVC_REQUIRE(slot && seq_off && rows && count && seq_off_c && N > 0 && max_rows > 0, "navc_compact_rows: bad arguments");
    if (max_rows > 32768)
        compact_rows_serial_kernel<<<1, 1024, 0, as_stream(stream)>>>(slot, seq_off, N, max_rows, rows, count, seq_off_c);
    else
        compact_rows_kernel<<<1, 1024, 0, as_stream(stream)>>>(slot, seq_off, N, max_rows, rows, count, seq_off_c);
    return check_launch("navc_compact_rows");
}

extern "C" int navc_teacher_probs(const float* part_max, const float* part_sum, int n_tiles,
                                  const float* target_logit, const int32_t* lens, int N, int S, float* teacher,
                                  void* stream) {
    NAVC_REQUIRE(part_max && part_sum && target_logit && lens && teacher && n_tiles > 0,
                 "navc_teacher_probs: bad arguments");
    teacher_probs_kernel<<<(N * S + 255) / 256, 256, 0, as_stream(stream)>>>(part_max, part_sum, n_tiles,
                                                                            target_logit, lens, N, S, teacher);
    return check_launch("navc_teacher_probs");
}

extern "C" int navc_teacher_inputs(const int64_t* tokens, const int64_t* map, int N, int S, int64_t* shifted,
                                   int64_t* mapped, void* stream) {
    NAVC_REQUIRE(tokens && shifted && mapped, "navc_teacher_inputs: null pointer");
    teacher_inputs_kernel<<<(N * S + 255) / 256, 256, 0, as_stream(stream)>>>(tokens, map, N, S, shifted, mapped);
    return check_launch("navc_teacher_inputs");
}

extern "C" int navc_select_best(const int64_t* tokens, const float* lprobs, const int32_t* lens, int B, int lbs,
                                int S, float alpha, int64_t* hyp, float* score, void* stream) {
    NAVC_REQUIRE(tokens && lprobs && lens && hyp, "navc_select_best: null pointer");
    const int wpb = 4;
    select_best_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(tokens, lprobs, lens, B, lbs, S,
                                                                               alpha, hyp, score);
    return check_launch("navc_select_best");
}
