// Attention cores (fp32 CUDA-core version): token self-attention and text->video cross-attention.
// The projections around them are GEMMs (gemm_simt.cu / gemm_tc.cu); these kernels do
// softmax(mask_fill(Q K^T / sqrt(dk), -1e7)) V per (sequence, head) with K/V staged in shared
// memory and one warp per query row.  Masks are derived from the token ids in-kernel.
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
    if (E <= 32) return launch_cross<1>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
    if (E <= 64) return launch_cross<2>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
    if (E <= 128) return launch_cross<4>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
    return launch_cross<8>(q, ldq, kv, ldkv, N, S, E, D, H, group, ctx_f32, ctx_hi, ctx_lo, probs, st);
}
