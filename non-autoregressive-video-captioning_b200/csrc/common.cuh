// Shared device/host helpers for libnavc (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/navc.h"

namespace navc {

// ---- error reporting (thread local message, C ABI returns int) ---------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define NAVC_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            navc::set_error(__VA_ARGS__);  \
            return 1;                      \
        }                                  \
    } while (0)

#define NAVC_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            navc::set_error("%s failed: %s", #call, cudaGetErrorString(e__));            \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- bf16 split ------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t bf16_bits(float x) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float bf16_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
    hi = bf16_bits(x);
    lo = bf16_bits(x - bf16_to_f32(hi));
}

// Packed split of two floats: one F2FP (cvt.rn.bf16x2.f32) per pair instead of two F2F per element
// (F2F runs on a slow pipe; the packed form is an ALU instruction).  Low half = first element.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hb << 16), hbf = __uint_as_float(hb & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hbf);
    hi = hb;
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 4 consecutive values -> packed hi / lo words
__device__ __forceinline__ void split_bf16x4(float4 v, uint2& hi, uint2& lo) {
    split_bf16x2(v.x, v.y, hi.x, lo.x);
    split_bf16x2(v.z, v.w, hi.y, lo.y);
}

// ---- activations (models/bert.py:9-19) ----------------------------------------------------------
__device__ __forceinline__ float act_apply(float x, int act) {
    switch (act) {
        case NAVC_ACT_GELU_NEW: {
            const float c = 0.7978845608028654f;  // sqrt(2/pi)
            float x3 = x * x * x;
            return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * x3)));
        }
        case NAVC_ACT_GELU: return x * 0.5f * (1.0f + erff(x * 0.7071067811865475f));
        case NAVC_ACT_RELU: return fmaxf(x, 0.0f);
        case NAVC_ACT_SWISH: return x / (1.0f + expf(-x));
        default: return x;
    }
}

// ex2.approx (MUFU): 2 ulp, exp2(-inf) = 0
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Tensor-core GEMM epilogues: gelu_new / swish through one MUFU exp + one MUFU rcp.
// 0.5*x*(1+tanh(u)) == x*sigmoid(2u) == x / (1 + exp(-2u)) exactly; this form has no cancellation
// for u << 0 (where the reference's own 1+tanh(u) loses digits), abs error <~ 2e-7*|x|.
__device__ __forceinline__ float act_apply_fast(float x, int act) {
    switch (act) {
        case NAVC_ACT_GELU_NEW: {
            const float k = -2.0f * 0.7978845608028654f * 1.4426950408889634f;  // -2*sqrt(2/pi)*log2(e)
            const float u = x * fmaf(0.044715f * x, x, 1.0f);
            return __fdividef(x, 1.0f + fast_exp2(k * u));
        }
        case NAVC_ACT_SWISH: return __fdividef(x, 1.0f + fast_exp2(-1.4426950408889634f * x));
        case NAVC_ACT_RELU: return fmaxf(x, 0.0f);
        default: return act_apply(x, act);
    }
}

// Epilogue for `n` consecutive columns of one row (n <= 8 typical); col0 multiple of 4 when n>=4.
struct EpiParams {
    const float* bias;
    const float* residual;
    const int64_t* row_tokens;
    int act;
    int ld_res;
    float* out_f32;
    uint16_t* out_hi;
    uint16_t* out_lo;
    int ld_out;
    int dbg;  // profiling aid (navc_epilogue_t.reserved): 1 = skip epilogue, 2 = no phase 2, 3 = no global stores
    int split_k;     // tcgen05 path only
    int accumulate;  // out_f32 += (atomic)
    const uint16_t* res_hi;  // bf16 hi/lo residual (tcgen05 pair epilogue)
    const uint16_t* res_lo;
    const int* m_dev;        // device-side row count (tcgen05 path)
    int m_hint;              // host estimate of *m_dev (tile-shape choice only; 0 = none)
    // split-K of the tiles of the last, partly filled wave of the persistent grid (set by the launcher, pair epilogue):
    float* sk_ws;            // partial accumulator tiles [slot][128][TBN] fp32, NULL = off
    int* sk_cnt;             // [2 * grid] arrival / departure counters, all zero between launches
    int* sk_err;             // set to 1 if a finisher gave up waiting (never observed; keeps a bug from hanging the GPU)
};
static inline EpiParams to_params(const navc_epilogue_t* e) {
    EpiParams p;
    p.bias = e->bias; p.residual = e->residual; p.row_tokens = e->row_tokens; p.act = e->act;
    p.ld_res = e->ld_res; p.out_f32 = e->out_f32; p.out_hi = e->out_hi; p.out_lo = e->out_lo;
    p.ld_out = e->ld_out;
    p.dbg = e->reserved;
    p.split_k = e->split_k > 1 ? e->split_k : 1;
    p.accumulate = (e->accumulate != 0 || p.split_k > 1) ? 1 : 0;
    p.res_hi = e->res_hi;
    p.res_lo = e->res_lo;
    p.m_dev = e->m_dev;
    p.m_hint = e->m_hint;
    p.sk_ws = nullptr; p.sk_cnt = nullptr; p.sk_err = nullptr;
    return p;
}

// store 4 consecutive columns [col, col+4) of `row`; caller guarantees col+4 <= N or uses the tail path
__device__ __forceinline__ void epi_store4(const EpiParams& p, int row, int col, float4 v, bool row_zero) {
    if (p.bias) {
        float4 b = *reinterpret_cast<const float4*>(p.bias + col);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (p.act) {
        v.x = act_apply(v.x, p.act); v.y = act_apply(v.y, p.act);
        v.z = act_apply(v.z, p.act); v.w = act_apply(v.w, p.act);
    }
    if (p.residual) {
        float4 r = *reinterpret_cast<const float4*>(p.residual + (size_t)row * p.ld_res + col);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (row_zero) v = make_float4(0.f, 0.f, 0.f, 0.f);
    size_t o = (size_t)row * p.ld_out + col;
    if (p.accumulate) {
        atomicAdd(reinterpret_cast<float4*>(p.out_f32 + o), v);
        return;
    }
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = v;
    if (p.out_hi) {
        uint2 hv, lv;
        split_bf16x4(v, hv, lv);
        *reinterpret_cast<uint2*>(p.out_hi + o) = hv;
        if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + o) = lv;
    }
}
__device__ __forceinline__ void epi_store1(const EpiParams& p, int row, int col, float v, bool row_zero) {
    if (p.bias) v += p.bias[col];
    if (p.act) v = act_apply(v, p.act);
    if (p.residual) v += p.residual[(size_t)row * p.ld_res + col];
    if (row_zero) v = 0.f;
    size_t o = (size_t)row * p.ld_out + col;
    if (p.accumulate) {
        atomicAdd(p.out_f32 + o, v);
        return;
    }
    if (p.out_f32) p.out_f32[o] = v;
    if (p.out_hi) {
        uint16_t h, l;
        split_bf16(v, h, l);
        p.out_hi[o] = h;
        if (p.out_lo) p.out_lo[o] = l;
    }
}

// ---- warp helpers --------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// online-softmax partial (max, sum of exp(x-max), argmax with lowest index on ties)
struct SoftPart {
    float m, s;
    int i;
};
__device__ __forceinline__ SoftPart soft_combine(SoftPart a, SoftPart b) {
    SoftPart r;
    if (b.m > a.m || (b.m == a.m && b.i < a.i)) { r.m = b.m; r.i = b.i; } else { r.m = a.m; r.i = a.i; }
    float sa = (a.s == 0.f) ? 0.f : a.s * expf(a.m - r.m);
    float sb = (b.s == 0.f) ? 0.f : b.s * expf(b.m - r.m);
    r.s = sa + sb;
    return r;
}

}  // namespace navc
