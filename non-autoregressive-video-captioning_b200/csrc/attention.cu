// Attention cores (fp32 CUDA cores, exact in every precision mode): token self-attention and
// text->video cross-attention.  The projections around them are GEMMs (gemm_simt.cu / gemm_tc.cu);
// these kernels do softmax(mask_fill(Q K^T / sqrt(dk), -1e7)) V per (key/value owner, head).
// Masks are derived from the token ids in-kernel.
//
// Fast path `attn_tile_kernel`: one block per (K/V owner, head, query chunk).  K is staged
// TRANSPOSED in shared memory ([dk][keys]) and V row-major, so the inner loops are one broadcast
// LDS.128 per four FMAs; every query is owned by TPQ threads that keep its KH scores in registers
// (softmax needs no shuffles beyond the TPQ pair) and then accumulate dk/TPQ output columns each.
// The warp-per-query kernels below remain as the generic fallback (any dk, S <= 128, E <= 256).
#include "common.cuh"

namespace navc {

constexpr float kMaskFill = -10e6f;  // models/bert.py:161

__device__ __forceinline__ void store_ctx(float v, size_t o, float* f32, uint16_t* hi, uint16_t* lo) {
    if (f32) f32[o] = v;
    if (hi) {
        uint16_t h, l;
        split_bf16(v, h, l);
        hi[o] = h;
        if (lo) lo[o] = l;
    }
}

// grid (N, ceil(H / hpb)); block = hpb warps; each warp owns one head of one sequence.
template <int KPL>
__global__ void self_attention_kernel(const float* __restrict__ qkv, int ld, const int64_t* __restrict__ tokens,
                                      int N, int S, int D, int H, int dk, int mask_kind, int watch, int hpb,
                                      float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, float* probs) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x, h = blockIdx.y * hpb + warp;
    if (h >= H) return;
    const int spad = KPL * 32;
    const int per_warp = S * (dk + 1) + S * dk + dk + spad;
    float* Ks = sm + (size_t)warp * per_warp;
    float* Vs = Ks + S * (dk + 1);
    float* Qs = Vs + S * dk;
    float* Ps = Qs + dk;

    const float* base = qkv + (size_t)n * S * ld + h * dk;
    for (int idx = lane; idx < S * dk; idx += 32) {
        int j = idx / dk, d = idx - j * dk;
        Ks[j * (dk + 1) + d] = base[(size_t)j * ld + D + d];
        Vs[j * dk + d] = base[(size_t)j * ld + 2 * D + d];
    }
    bool kpad[KPL];
#pragma unroll
    for (int t = 0; t < KPL; ++t) {
        int j = lane + 32 * t;
        kpad[t] = (j < S) ? (tokens[(size_t)n * S + j] == NAVC_PAD) : true;
    }
    const float sqrt_dk = sqrtf((float)dk);
    const bool use_watch = (mask_kind == NAVC_MASK_CAUSAL) && watch != 0 && S >= watch;
    __syncwarp();

    for (int i = 0; i < S; ++i) {
        for (int d = lane; d < dk; d += 32) Qs[d] = base[(size_t)i * ld + d];
        __syncwarp();
        float sc[KPL];
        float m = -INFINITY;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            int j = lane + 32 * t;
            sc[t] = -INFINITY;
            if (j < S) {
                const float* kr = Ks + j * (dk + 1);
                float s = 0.f;
                for (int d = 0; d < dk; ++d) s = fmaf(Qs[d], kr[d], s);
                s = s / sqrt_dk;
                bool masked = kpad[t];
                if (mask_kind == NAVC_MASK_CAUSAL) masked = masked || (j > i) || (use_watch && j <= i - watch);
                if (mask_kind == NAVC_MASK_SELF) masked = masked || (j == i);
                sc[t] = masked ? kMaskFill : s;
                m = fmaxf(m, sc[t]);
            }
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            int j = lane + 32 * t;
            sc[t] = (j < S) ? expf(sc[t] - m) : 0.f;
            sum += sc[t];
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            int j = lane + 32 * t;
            if (j < S) {
                float p = sc[t] / sum;
                Ps[j] = p;
                if (probs) probs[(((size_t)h * N + n) * S + i) * S + j] = p;
            }
        }
        __syncwarp();
        for (int d = lane; d < dk; d += 32) {
            float o = 0.f;
            for (int j = 0; j < S; ++j) o = fmaf(Ps[j], Vs[j * dk + d], o);
            store_ctx(o, ((size_t)n * S + i) * D + h * dk + d, ctx_f32, ctx_hi, ctx_lo);
        }
        __syncwarp();
    }
}

// grid (N/group, H); block = nw warps.  K/V of (video, head) staged once, shared by the `group`
// candidate rows of that video; warps take queries round-robin.
template <int KPL>
__global__ void cross_attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ kv,
                                       int ldkv, int N, int S, int E, int D, int H, int dk, int group,
                                       float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, float* probs) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int g = blockIdx.x, h = blockIdx.y;
    const int epad = KPL * 32;
    float* Ks = sm;
    float* Vs = Ks + E * (dk + 1);
    float* Qs = Vs + E * dk + (size_t)warp * (dk + epad);
    float* Ps = Qs + dk;

    const float* kbase = kv + (size_t)g * E * ldkv + h * dk;
    for (int idx = threadIdx.x; idx < E * dk; idx += blockDim.x) {
        int j = idx / dk, d = idx - j * dk;
        Ks[j * (dk + 1) + d] = kbase[(size_t)j * ldkv + d];
        Vs[j * dk + d] = kbase[(size_t)j * ldkv + D + d];
    }
    __syncthreads();
    const float sqrt_dk = sqrtf((float)dk);
    const int nq = group * S;
    for (int qi = warp; qi < nq; qi += nw) {
        const int n = g * group + qi / S, i = qi % S;
        if (n >= N) break;
        const float* qr = q + ((size_t)n * S + i) * ldq + h * dk;
        for (int d = lane; d < dk; d += 32) Qs[d] = qr[d];
        __syncwarp();
        float sc[KPL];
        float m = -INFINITY;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            int j = lane + 32 * t;
            sc[t] = -INFINITY;
            if (j < E) {
                const float* kr = Ks + j * (dk + 1);
                float s = 0.f;
                for (int d = 0; d < dk; ++d) s = fmaf(Qs[d], kr[d], s);
                sc[t] = s / sqrt_dk;
                m = fmaxf(m, sc[t]);
            }
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            int j = lane + 32 * t;
            sc[t] = (j < E) ? expf(sc[t] - m) : 0.f;
            sum += sc[t];
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            int j = lane + 32 * t;
            if (j < E) {
                float p = sc[t] / sum;
                Ps[j] = p;
                if (probs) probs[(((size_t)h * N + n) * S + i) * E + j] = p;
            }
        }
        __syncwarp();
        for (int d = lane; d < dk; d += 32) {
            float o = 0.f;
            for (int j = 0; j < E; ++j) o = fmaf(Ps[j], Vs[j * dk + d], o);
            store_ctx(o, ((size_t)n * S + i) * D + h * dk + d, ctx_f32, ctx_hi, ctx_lo);
        }
        __syncwarp();
    }
}


// ---- register-tiled fast path ---------------------------------------------------------------------
// q      [.., ldq]   query rows; block (g, h, z) owns query rows g*NQ + z*QB + [0, QB) (clipped to NQ)
// k, v   [.., ldkv]  key/value rows g*Sk + [0, Sk)
// tokens [.., S]     self-attention only (NQ == Sk == S, g = sequence): key-pad / causal / diagonal masks
template <int DK, int KH, int TPQ>
__global__ void __launch_bounds__(384) attn_tile_kernel(
    const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v, int ldkv,
    const int64_t* __restrict__ tokens, int NQ, int QB, int S, int Sk, int D, int H, int mask_kind, int watch,
    float* __restrict__ ctx_f32, uint16_t* __restrict__ ctx_hi, uint16_t* __restrict__ ctx_lo,
    float* __restrict__ probs, int n_seq_total) {
    constexpr int KP = KH * TPQ;   // padded key count
    constexpr int DT = DK / TPQ;   // output columns per thread
    extern __shared__ __align__(16) float sm[];
    float* Kt = sm;                      // [DK][KP]
    float* Vs = Kt + DK * KP;            // [Sk][DK]
    float* QP = Vs + (size_t)Sk * DK;    // Q rows [QB][DK+1] during QK^T, then P rows [QB][KP+1]
    const int g = blockIdx.x, h = blockIdx.y;
    const int q_lo = blockIdx.z * QB;
    const int nq = min(QB, NQ - q_lo);
    const int tid = threadIdx.x, nthr = blockDim.x;

    // ---- stage K (transposed), V, Q ----
    const float* kb = k + (size_t)g * Sk * ldkv + h * DK;
    const float* vb = v + (size_t)g * Sk * ldkv + h * DK;
    for (int idx = tid; idx < KP * (DK / 4); idx += nthr) {
        const int j = idx % KP, d4 = idx / KP;
        float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < Sk) kv4 = *reinterpret_cast<const float4*>(kb + (size_t)j * ldkv + d4 * 4);
        Kt[(d4 * 4 + 0) * KP + j] = kv4.x; Kt[(d4 * 4 + 1) * KP + j] = kv4.y;
        Kt[(d4 * 4 + 2) * KP + j] = kv4.z; Kt[(d4 * 4 + 3) * KP + j] = kv4.w;
    }
    for (int idx = tid; idx < Sk * (DK / 4); idx += nthr) {
        const int j = idx / (DK / 4), d4 = idx % (DK / 4);
        *reinterpret_cast<float4*>(Vs + j * DK + d4 * 4) = *reinterpret_cast<const float4*>(vb + (size_t)j * ldkv + d4 * 4);
    }
    const float* qb = q + ((size_t)g * NQ + q_lo) * ldq + h * DK;
    for (int idx = tid; idx < nq * (DK / 4); idx += nthr) {
        const int i = idx / (DK / 4), d4 = idx % (DK / 4);
        const float4 qv = *reinterpret_cast<const float4*>(qb + (size_t)i * ldq + d4 * 4);
        float* dst = QP + i * (DK + 1) + d4 * 4;
        dst[0] = qv.x; dst[1] = qv.y; dst[2] = qv.z; dst[3] = qv.w;
    }
    __syncthreads();

    const int qi = tid / TPQ, t = tid % TPQ;
    const bool live = qi < nq;
    const int qs = live ? qi : 0;

    // ---- scores: KH keys per thread in registers ----
    float sc[KH];
#pragma unroll
    for (int i = 0; i < KH; ++i) sc[i] = 0.f;
    {
        const float* qrow = QP + qs * (DK + 1);
        const float* kt = Kt + t * KH;
#pragma unroll 2
        for (int d = 0; d < DK; ++d) {
            const float qd = qrow[d];
#pragma unroll
            for (int i4 = 0; i4 < KH / 4; ++i4) {
                const float4 kk = *reinterpret_cast<const float4*>(kt + d * KP + i4 * 4);
                sc[i4 * 4 + 0] = fmaf(qd, kk.x, sc[i4 * 4 + 0]);
                sc[i4 * 4 + 1] = fmaf(qd, kk.y, sc[i4 * 4 + 1]);
                sc[i4 * 4 + 2] = fmaf(qd, kk.z, sc[i4 * 4 + 2]);
                sc[i4 * 4 + 3] = fmaf(qd, kk.w, sc[i4 * 4 + 3]);
            }
        }
    }
    __syncthreads();  // all Q rows consumed: the region is reused for P

    // ---- mask + softmax (models/bert.py:157-164) ----
    const float sqrt_dk = sqrtf((float)DK);
    const int ipos = (q_lo + qs) % S;  // position of this query inside its sequence
    const bool use_watch = (mask_kind == NAVC_MASK_CAUSAL) && watch != 0 && S >= watch;
    const int64_t* trow = tokens ? tokens + (size_t)g * S : nullptr;
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        const int j = t * KH + i;
        float sv = -INFINITY;
        if (j < Sk) {
            sv = sc[i] / sqrt_dk;
            if (trow) {
                bool masked = trow[j] == NAVC_PAD;
                if (mask_kind == NAVC_MASK_CAUSAL) masked = masked || (j > ipos) || (use_watch && j <= ipos - watch);
                if (mask_kind == NAVC_MASK_SELF) masked = masked || (j == ipos);
                if (masked) sv = kMaskFill;
            }
        }
        sc[i] = sv;
        m = fmaxf(m, sv);
    }
    if (TPQ == 2) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
        sc[i] = expf(sc[i] - m);  // exp(-inf) = 0 for the padded keys
        sum += sc[i];
    }
    if (TPQ == 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    {
        float* prow = QP + qs * (KP + 1) + t * KH;
        const int n_glob = g * (NQ / S) + (q_lo + qs) / S;  // sequence index for the probs output
        float* pg = (probs && live) ? probs + (((size_t)h * n_seq_total + n_glob) * S + ipos) * Sk : nullptr;
#pragma unroll
        for (int i = 0; i < KH; ++i) {
            const float pv = sc[i] / sum;
            if (live) prow[i] = pv;
            if (pg && t * KH + i < Sk) pg[t * KH + i] = pv;
        }
    }
    __syncwarp();

    // ---- context: DT output columns per thread ----
    float out[DT];
#pragma unroll
    for (int i = 0; i < DT; ++i) out[i] = 0.f;
    {
        const float* prow = QP + qs * (KP + 1);
        const float* vcol = Vs + t * DT;
#pragma unroll 4
        for (int j = 0; j < Sk; ++j) {
            const float pj = prow[j];
#pragma unroll
            for (int mm = 0; mm < DT / 4; ++mm) {
                const float4 vv = *reinterpret_cast<const float4*>(vcol + j * DK + mm * 4);
                out[mm * 4 + 0] = fmaf(pj, vv.x, out[mm * 4 + 0]);
                out[mm * 4 + 1] = fmaf(pj, vv.y, out[mm * 4 + 1]);
                out[mm * 4 + 2] = fmaf(pj, vv.z, out[mm * 4 + 2]);
                out[mm * 4 + 3] = fmaf(pj, vv.w, out[mm * 4 + 3]);
            }
        }
    }
    if (live) {
        const size_t o = ((size_t)g * NQ + q_lo + qi) * D + h * DK + t * DT;
#pragma unroll
        for (int mm = 0; mm < DT / 4; ++mm) {
            const float4 vv = make_float4(out[mm * 4], out[mm * 4 + 1], out[mm * 4 + 2], out[mm * 4 + 3]);
            if (ctx_f32) *reinterpret_cast<float4*>(ctx_f32 + o + mm * 4) = vv;
            if (ctx_hi) {
                uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_bf16(vv.x, h0, l0); split_bf16(vv.y, h1, l1); split_bf16(vv.z, h2, l2); split_bf16(vv.w, h3, l3);
                *reinterpret_cast<uint2*>(ctx_hi + o + mm * 4) =
                    make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
                if (ctx_lo)
                    *reinterpret_cast<uint2*>(ctx_lo + o + mm * 4) =
                        make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
            }
        }
    }
}

// Returns 0 = launched, -1 = shape not covered by the fast path (caller falls back), >0 = error.
template <int DK, int KH, int TPQ>
static int launch_tile(const float* q, int ldq, const float* k, const float* v, int ldkv, const int64_t* tokens,
                       int G, int NQ, int S, int Sk, int D, int H, int mask_kind, int watch, float* f32,
                       uint16_t* hi, uint16_t* lo, float* probs, int n_seq_total, cudaStream_t st, const char* what) {
    constexpr int KP = KH * TPQ;
    const int qb_max = 384 / TPQ;
    const int QB = NQ < qb_max ? NQ : qb_max;
    const int threads = ((QB * TPQ + 31) / 32) * 32;
    const size_t qp = (size_t)QB * ((KP + 1) > (DK + 1) ? (KP + 1) : (DK + 1));
    const size_t smem = ((size_t)DK * KP + (size_t)Sk * DK + qp) * sizeof(float);
    if (smem > 220 * 1024) return -1;
    auto kern = attn_tile_kernel<DK, KH, TPQ>;
    if (smem > 48 * 1024) NAVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(G, H, (NQ + QB - 1) / QB);
    kern<<<grid, threads, smem, st>>>(q, ldq, k, v, ldkv, tokens, NQ, QB, S, Sk, D, H, mask_kind, watch, f32, hi, lo,
                                      probs, n_seq_total);
    return check_launch(what);
}

template <int KH, int TPQ>
static int dispatch_tile(int dk, const float* q, int ldq, const float* k, const float* v, int ldkv,
                         const int64_t* tokens, int G, int NQ, int S, int Sk, int D, int H, int mask_kind, int watch,
                         float* f32, uint16_t* hi, uint16_t* lo, float* probs, int n_seq_total, cudaStream_t st,
                         const char* what) {
    if (dk == 64) return launch_tile<64, KH, TPQ>(q, ldq, k, v, ldkv, tokens, G, NQ, S, Sk, D, H, mask_kind, watch, f32, hi, lo, probs, n_seq_total, st, what);
    if (dk == 32) return launch_tile<32, KH, TPQ>(q, ldq, k, v, ldkv, tokens, G, NQ, S, Sk, D, H, mask_kind, watch, f32, hi, lo, probs, n_seq_total, st, what);
    if (dk == 16) return launch_tile<16, KH, TPQ>(q, ldq, k, v, ldkv, tokens, G, NQ, S, Sk, D, H, mask_kind, watch, f32, hi, lo, probs, n_seq_total, st, what);
    return -1;
}

template <int KPL>
static int launch_self(const float* qkv, int ld, const int64_t* tokens, int N, int S, int D, int H, int mask_kind,
                       int watch, float* f32, uint16_t* hi, uint16_t* lo, float* probs, cudaStream_t st) {
    const int dk = D / H;
    int hpb = H < 4 ? H : 4;
    size_t per_warp = (size_t)(S * (dk + 1) + S * dk + dk + KPL * 32) * sizeof(float);
    while (hpb > 1 && per_warp * hpb > 200 * 1024) --hpb;
    size_t smem = per_warp * hpb;
    NAVC_REQUIRE(smem <= 227 * 1024, "navc_self_attention: S*dk too large for shared memory");
    auto kern = self_attention_kernel<KPL>;
    if (smem > 48 * 1024) NAVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(N, (H + hpb - 1) / hpb);
    kern<<<grid, hpb * 32, smem, st>>>(qkv, ld, tokens, N, S, D, H, dk, mask_kind, watch, hpb, f32, hi, lo, probs);
    return check_launch("navc_self_attention");
}

template <int KPL>
static int launch_cross(const float* q, int ldq, const float* kv, int ldkv, int N, int S, int E, int D, int H,
                        int group, float* f32, uint16_t* hi, uint16_t* lo, float* probs, cudaStream_t st) {
    const int dk = D / H;
    const int nw = 8;
    size_t smem = (size_t)(E * (dk + 1) + E * dk + nw * (dk + KPL * 32)) * sizeof(float);
    NAVC_REQUIRE(smem <= 227 * 1024, "navc_cross_attention: E*dk too large for shared memory");
    auto kern = cross_attention_kernel<KPL>;
    if (smem > 48 * 1024) NAVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + group - 1) / group, H);
    kern<<<grid, nw * 32, smem, st>>>(q, ldq, kv, ldkv, N, S, E, D, H, dk, group, f32, hi, lo, probs);
    return check_launch("navc_cross_attention");
}

}  // namespace navc

using namespace navc;

extern "C" int navc_self_attention(const float* qkv, int ld, const int64_t* tokens, int N, int S, int D, int H,
                                   int mask_kind, int watch, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                                   float* probs, void* stream) {
    NAVC_REQUIRE(qkv && tokens && (ctx_f32 || ctx_hi), "navc_self_attention: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && H > 0 && D % H == 0 && ld >= 3 * D, "navc_self_attention: bad shape");
    NAVC_REQUIRE(S <= 128, "navc_self_attention: S > 128 unsupported (max_len is 30 in the reference)");
    NAVC_REQUIRE(mask_kind >= 0 && mask_kind <= 2, "navc_self_attention: bad mask kind");
    cudaStream_t st = as_stream(stream);
    if (ld % 4 == 0 && D % 4 == 0 && (((uintptr_t)qkv) & 15) == 0 && S <= 64) {
        const int dk = D / H;
        int rc = (S <= 32)
            ? dispatch_tile<32, 1>(dk, qkv, ld, qkv + D, qkv + 2 * D, ld, tokens, N, S, S, S, D, H, mask_kind, watch, ctx_f32, ctx_hi, ctx_lo, probs, N, st, "navc_self_attention")
            : dispatch_tile<64, 1>(dk, qkv, ld, qkv + D, qkv + 2 * D, ld, tokens, N, S, S, S, D, H, mask_kind, watch, ctx_f32, ctx_hi, ctx_lo, probs, N, st, "navc_self_attention");
        if (rc >= 0) return rc;
    }
    if (S <= 32) return launch_self<1>(qkv, ld, tokens, N, S, D, H, mask_kind, watch, ctx_f32, ctx_hi, ctx_lo, probs, st);
    if (S <= 64) return launch_self<2>(qkv, ld, tokens, N, S, D, H, mask_kind, watch, ctx_f32, ctx_hi, ctx_lo, probs, st);
    return launch_self<4>(qkv, ld, tokens, N, S, D, H, mask_kind, watch, ctx_f32, ctx_hi, ctx_lo, probs, st);
}

extern "C" int navc_cross_attention(const float* q, int ldq, const float* kv, int ldkv, int N, int S, int E, int D,
                                    int H, int group, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                                    float* probs, void* stream) {
    NAVC_REQUIRE(q && kv && (ctx_f32 || ctx_hi), "navc_cross_attention: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && E > 0 && H > 0 && D % H == 0 && group >= 1 && N % group == 0,
                 "navc_cross_attention: bad shape");
    NAVC_REQUIRE(E <= 256, "navc_cross_attention: E > 256 unsupported");
    cudaStream_t st = as_stream(stream);
    if (ldq % 4 == 0 && ldkv % 4 == 0 && D % 4 == 0 && ((((uintptr_t)q) | ((uintptr_t)kv)) & 15) == 0 && E <= 128) {
        const int dk = D / H;
        const int G = N / group;
        int rc = (E <= 64)
            ? dispatch_tile<32, 2>(dk, q, ldq, kv, kv + D, ldkv, nullptr, G, group * S, S, E, D, H, 0, 0, ctx_f32, ctx_hi, ctx_lo, probs, N, st, "navc_cross_attention")
            : dispatch_tile<64, 2>(dk, q, ldq, kv, kv + D, ldkv, nullptr, G, group * S, S, E, D, H, 0, 0, ctx_f32, ctx_hi, ctx_lo, probs, N, st, "navc_cross_attention");
        if (rc >= 0) return rc;
    }
    if (E <= 32) return launch_cross<1>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
    if (E <= 64) return launch_cross<2>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
    if (E <= 128) return launch_cross<4>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
    return launch_cross<8>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
}
