// Training path: train-mode forward helpers (dropout, BatchNorm batch statistics) and the
// hand-written backward kernels of every elementwise / normalisation / attention op of the hot
// path.  Gradient GEMMs reuse gemm_tc.cu / gemm_simt.cu on operands produced by
// navc_transpose_pack.  Everything here is fp32 and HBM- or latency-bound.
#include "common.cuh"

namespace navc {

// ---- counter-based dropout ------------------------------------------------------------------------
// keep(seed, i): one splitmix64 hash per PAIR of consecutive elements (low / high 32 bits), compared
// against thr = p * 2^32.  Forward and backward regenerate the same mask from (seed, index).
struct DropCfg {
    uint64_t seed;
    uint32_t thr;   // 0 = dropout off
    float scale;    // 1 / (1 - p)
};
static inline DropCfg make_drop(uint64_t seed, float p) {
    DropCfg d;
    d.seed = seed;
    d.thr = (p <= 0.f) ? 0u : (uint32_t)fmin(4294967295.0, (double)p * 4294967296.0);
    d.scale = (p <= 0.f) ? 1.f : 1.0f / (1.0f - p);
    return d;
}
__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t pair) {
    uint64_t z = seed + (pair + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float drop_factor1(const DropCfg& d, uint64_t idx) {
    const uint64_t z = mix64(d.seed, idx >> 1);
    const uint32_t u = (idx & 1) ? (uint32_t)(z >> 32) : (uint32_t)z;
    return u >= d.thr ? d.scale : 0.f;
}
// factors for the 4 consecutive elements e..e+3 (e % 4 == 0)
__device__ __forceinline__ float4 drop_factor4(const DropCfg& d, uint64_t e) {
    const uint64_t z0 = mix64(d.seed, e >> 1), z1 = mix64(d.seed, (e >> 1) + 1);
    float4 f;
    f.x = (uint32_t)z0 >= d.thr ? d.scale : 0.f;
    f.y = (uint32_t)(z0 >> 32) >= d.thr ? d.scale : 0.f;
    f.z = (uint32_t)z1 >= d.thr ? d.scale : 0.f;
    f.w = (uint32_t)(z1 >> 32) >= d.thr ? d.scale : 0.f;
    return f;
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ void store_out4(float4 v, size_t o, float* f32, uint16_t* hi, uint16_t* lo) {
    if (f32) *reinterpret_cast<float4*>(f32 + o) = v;
    if (hi) {
        uint2 hv, lv;
        split_bf16x4(v, hv, lv);
        *reinterpret_cast<uint2*>(hi + o) = hv;
        if (lo) *reinterpret_cast<uint2*>(lo + o) = lv;
    }
}

static inline int ew_blocks(int64_t n_items, int threads) {
    int64_t b = (n_items + threads - 1) / threads;
    if (b > 148 * 16) b = 148 * 16;
    return b < 1 ? 1 : (int)b;
}

// ---- dropout + residual + row mask ------------------------------------------------------------------
__global__ void drop_add_kernel(const float* __restrict__ y, const float* __restrict__ res, DropCfg d1, DropCfg d2,
                                const int64_t* __restrict__ row_tokens, int64_t n4, int D, float* o32,
                                uint16_t* ohi, uint16_t* olo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        float4 v = *reinterpret_cast<const float4*>(y + e);
        if (d1.thr) v = mul4(v, drop_factor4(d1, (uint64_t)e));
        if (res) v = add4(v, *reinterpret_cast<const float4*>(res + e));
        if (d2.thr) v = mul4(v, drop_factor4(d2, (uint64_t)e));
        if (row_tokens && row_tokens[e / D] == NAVC_PAD) v = make_float4(0.f, 0.f, 0.f, 0.f);
        store_out4(v, (size_t)e, o32, ohi, olo);
    }
}

__global__ void drop_add_bwd_kernel(const float* __restrict__ dout, DropCfg d1, DropCfg d2,
                                    const int64_t* __restrict__ row_tokens, int64_t n4, int D, float* d_y,
                                    float* d_res) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        float4 g = *reinterpret_cast<const float4*>(dout + e);
        if (row_tokens && row_tokens[e / D] == NAVC_PAD) g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d2.thr) g = mul4(g, drop_factor4(d2, (uint64_t)e));
        if (d_res) *reinterpret_cast<float4*>(d_res + e) = g;
        if (d1.thr) g = mul4(g, drop_factor4(d1, (uint64_t)e));
        if (d_y) *reinterpret_cast<float4*>(d_y + e) = g;
    }
}

// ---- activation (+dropout) --------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float x, int act) {
    switch (act) {
        case NAVC_ACT_GELU_NEW: {
            const float c = 0.7978845608028654f;
            const float t = tanhf(c * (x + 0.044715f * x * x * x));
            return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x * x);
        }
        case NAVC_ACT_GELU:
            return 0.5f * (1.0f + erff(x * 0.7071067811865475f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
        case NAVC_ACT_RELU: return x > 0.f ? 1.f : 0.f;
        case NAVC_ACT_SWISH: {
            const float s = 1.0f / (1.0f + expf(-x));
            return s + x * s * (1.0f - s);
        }
        default: return 1.f;
    }
}

__global__ void act_drop_kernel(const float* __restrict__ u, int act, DropCfg d, int64_t n, float* o32,
                                uint16_t* ohi, uint16_t* olo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        if (e + 4 <= n) {
            float4 v = *reinterpret_cast<const float4*>(u + e);
            v.x = act_apply(v.x, act); v.y = act_apply(v.y, act); v.z = act_apply(v.z, act); v.w = act_apply(v.w, act);
            if (d.thr) v = mul4(v, drop_factor4(d, (uint64_t)e));
            store_out4(v, (size_t)e, o32, ohi, olo);
        } else {
            for (int64_t j = e; j < n; ++j) {
                float v = act_apply(u[j], act);
                if (d.thr) v *= drop_factor1(d, (uint64_t)j);
                if (o32) o32[j] = v;
                if (ohi) {
                    uint16_t h, l;
                    split_bf16(v, h, l);
                    ohi[j] = h;
                    if (olo) olo[j] = l;
                }
            }
        }
    }
}

__global__ void act_drop_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ u, int act, DropCfg d,
                                    int64_t n, float* du) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float g = dout[j];
        if (d.thr) g *= drop_factor1(d, (uint64_t)j);
        du[j] = g * act_grad(u[j], act);
    }
}

// ---- transpose / split / column sums ------------------------------------------------------------------
// block (32, 8) handles a 32 x 32 tile of x [M, N].
__global__ void transpose_pack_kernel(const float* __restrict__ x, int M, int N, int ld, uint16_t* hi, uint16_t* lo,
                                      int ld_s, float* t32, uint16_t* thi, uint16_t* tlo, int ld_t, float* colsum) {
    __shared__ float tile[32][33];
    __shared__ float csum[8][32];
    const int n0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    float cs = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int m = m0 + ty + 8 * k, n = n0 + tx;
        float v = 0.f;
        if (m < M && n < N) v = x[(size_t)m * ld + n];
        if (hi && m < M && n < ld_s) {  // pad columns N..ld_s-1 are written as zeros
            uint16_t h, l;
            split_bf16(v, h, l);
            hi[(size_t)m * ld_s + n] = h;
            if (lo) lo[(size_t)m * ld_s + n] = l;
        }
        tile[ty + 8 * k][tx] = v;
        cs += v;
    }
    if (colsum) csum[ty][tx] = cs;
    __syncthreads();
    if (colsum && ty == 0 && n0 + tx < N) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += csum[k][tx];
        atomicAdd(colsum + n0 + tx, s);
    }
    if (t32 || thi) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n = n0 + ty + 8 * k, m = m0 + tx;
            if (n < N && m < ld_t) {
                const float v = tile[tx][ty + 8 * k];  // zero beyond M
                const size_t o = (size_t)n * ld_t + m;
                if (t32) t32[o] = v;
                if (thi) {
                    uint16_t h, l;
                    split_bf16(v, h, l);
                    thi[o] = h;
                    if (tlo) tlo[o] = l;
                }
            }
        }
    }
}

// The straight-only case (every gradient GEMM of the tensor-core path: bf16 hi / lo copy of dY + column sums = bias gradient,
// no transposed copy), vectorised: block (32, 8) covers 64 rows x 128 columns, thread = 4 columns of every 8th row, 128-bit
// loads / 64-bit stores, column sums in registers -> shared memory -> one atomicAdd per column and block.  (The tile kernel
// above moves 2-byte elements: 61 us for a [16.5k, 512] gradient, 6x off the HBM time, 8 % of a training step.)
__global__ void __launch_bounds__(256) split_colsum_kernel(const float* __restrict__ x, int M, int N, int ld, uint16_t* __restrict__ hi,
                                                         uint16_t* __restrict__ lo, int ld_s, float* __restrict__ colsum) {
    __shared__ float4 csum[8][32];
    const int c = (blockIdx.x * 32 + threadIdx.x) * 4;   // first of this thread's 4 columns
    const int m0 = blockIdx.y * 64;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < ld_s) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int m = m0 + threadIdx.y + 8 * k;
            if (m >= M) break;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c + 4 <= N) {
                v = *reinterpret_cast<const float4*>(x + (size_t)m * ld + c);
            } else if (c < N) {   // the row's last, partial group: columns >= N are pad (written as zeros)
                const float* xr = x + (size_t)m * ld + c;
                v.x = xr[0];
                if (c + 1 < N) v.y = xr[1];
                if (c + 2 < N) v.z = xr[2];
            }
            uint2 h, l;
            split_bf16x4(v, h, l);
            *reinterpret_cast<uint2*>(hi + (size_t)m * ld_s + c) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + (size_t)m * ld_s + c) = l;
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
        }
    }
    if (!colsum) return;
    csum[threadIdx.y][threadIdx.x] = cs;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float4 t = csum[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            const float4 u = csum[k][threadIdx.x];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        atomicAdd(colsum + c, t.x);
        if (c + 1 < N) atomicAdd(colsum + c + 1, t.y);
        if (c + 2 < N) atomicAdd(colsum + c + 2, t.z);
        if (c + 3 < N) atomicAdd(colsum + c + 3, t.w);
    }
}

// Producers that write dY of the next gradient GEMM pair directly in operand form (bf16 hi / lo rows + column sums = bias
// gradient) instead of an fp32 tensor that navc_transpose_pack would re-read: same tiling as split_colsum_kernel.
// dropout + residual backward (models/bert.py:193-200): d_res = dout * mask2 [* rowmask], dY = d_res * mask1
__global__ void __launch_bounds__(256) drop_add_bwd_split_kernel(const float* __restrict__ dout, DropCfg d1, DropCfg d2,
                                                               const int64_t* __restrict__ row_tokens, int M, int D,
                                                               float* __restrict__ d_res, uint16_t* __restrict__ hi,
                                                               uint16_t* __restrict__ lo, int ld_s, float* __restrict__ colsum) {
    __shared__ float4 csum[8][32];
    const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int m0 = blockIdx.y * 64;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int m = m0 + threadIdx.y + 8 * k;
            if (m >= M) break;
            const int64_t e = (int64_t)m * D + c;
            float4 g = *reinterpret_cast<const float4*>(dout + e);
            if (row_tokens && row_tokens[m] == NAVC_PAD) g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d2.thr) g = mul4(g, drop_factor4(d2, (uint64_t)e));
            if (d_res) *reinterpret_cast<float4*>(d_res + e) = g;
            if (d1.thr) g = mul4(g, drop_factor4(d1, (uint64_t)e));
            uint2 h, l;
            split_bf16x4(g, h, l);
            *reinterpret_cast<uint2*>(hi + (size_t)m * ld_s + c) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + (size_t)m * ld_s + c) = l;
            cs.x += g.x; cs.y += g.y; cs.z += g.z; cs.w += g.w;
        }
    }
    if (!colsum) return;
    csum[threadIdx.y][threadIdx.x] = cs;
    __syncthreads();
    if (threadIdx.y == 0 && c < D) {
        float4 t = csum[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            const float4 u = csum[k][threadIdx.x];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        atomicAdd(colsum + c, t.x); atomicAdd(colsum + c + 1, t.y); atomicAdd(colsum + c + 2, t.z); atomicAdd(colsum + c + 3, t.w);
    }
}

// activation (+ dropout) backward (models/bert.py:227-230): dY = dout * mask * act'(u)
__global__ void __launch_bounds__(256) act_drop_bwd_split_kernel(const float* __restrict__ dout, const float* __restrict__ u, int act,
                                                               DropCfg d, int M, int N, uint16_t* __restrict__ hi,
                                                               uint16_t* __restrict__ lo, int ld_s, float* __restrict__ colsum) {
    __shared__ float4 csum[8][32];
    const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int m0 = blockIdx.y * 64;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < N) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int m = m0 + threadIdx.y + 8 * k;
            if (m >= M) break;
            const int64_t e = (int64_t)m * N + c;
            float4 g = *reinterpret_cast<const float4*>(dout + e);
            if (d.thr) {
                g.x *= drop_factor1(d, (uint64_t)e); g.y *= drop_factor1(d, (uint64_t)e + 1);
                g.z *= drop_factor1(d, (uint64_t)e + 2); g.w *= drop_factor1(d, (uint64_t)e + 3);
            }
            const float4 uv = *reinterpret_cast<const float4*>(u + e);
            g.x *= act_grad(uv.x, act); g.y *= act_grad(uv.y, act); g.z *= act_grad(uv.z, act); g.w *= act_grad(uv.w, act);
            uint2 h, l;
            split_bf16x4(g, h, l);
            *reinterpret_cast<uint2*>(hi + (size_t)m * ld_s + c) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + (size_t)m * ld_s + c) = l;
            cs.x += g.x; cs.y += g.y; cs.z += g.z; cs.w += g.w;
        }
    }
    if (!colsum) return;
    csum[threadIdx.y][threadIdx.x] = cs;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float4 t = csum[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            const float4 v = csum[k][threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        atomicAdd(colsum + c, t.x); atomicAdd(colsum + c + 1, t.y); atomicAdd(colsum + c + 2, t.z); atomicAdd(colsum + c + 3, t.w);
    }
}

// ---- highway (train) ----------------------------------------------------------------------------------
__global__ void highway_fwd_train_kernel(const float* __restrict__ x, const float* __restrict__ yg, int gate,
                                         int64_t n, int D, DropCfg d, float* __restrict__ o) {
    const int ldy = gate ? 2 * D : D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / D;
        const int c = (int)(i - r * D);
        const float xv = x[i];
        const float y = tanhf(yg[r * ldy + c]);
        float ov;
        if (gate) {
            const float g = 1.0f / (1.0f + expf(-yg[r * ldy + D + c]));
            ov = g * xv + (1.0f - g) * y;
        } else {
            ov = xv + y;
        }
        if (d.thr) ov *= drop_factor1(d, (uint64_t)i);
        o[i] = ov;
    }
}

__global__ void highway_bwd_kernel(const float* __restrict__ d_o, const float* __restrict__ x,
                                   const float* __restrict__ yg, int gate, int64_t n, int D, DropCfg d,
                                   float* __restrict__ d_x, float* __restrict__ d_yg) {
    const int ldy = gate ? 2 * D : D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / D;
        const int c = (int)(i - r * D);
        float go = d_o[i];
        if (d.thr) go *= drop_factor1(d, (uint64_t)i);
        const float xv = x[i];
        const float y = tanhf(yg[r * ldy + c]);
        if (gate) {
            const float g = 1.0f / (1.0f + expf(-yg[r * ldy + D + c]));
            d_x[i] = go * g;
            d_yg[r * ldy + c] = go * (1.0f - g) * (1.0f - y * y);
            d_yg[r * ldy + D + c] = go * (xv - y) * g * (1.0f - g);
        } else {
            d_x[i] = go;
            d_yg[r * ldy + c] = go * (1.0f - y * y);
        }
    }
}

// ---- BatchNorm batch statistics / apply / backward ------------------------------------------------------
// block (32, 8): 32 columns x a chunk of rows; pass 0 accumulates sum(x) into `a`, pass 1 accumulates
// sum((x - a/M)^2) into `b`.
__global__ void bn_colreduce_kernel(const float* __restrict__ o, int M, int D, int rows_per_block, int pass, float* a,
                                    float* b) {
    __shared__ float red[8][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float acc = 0.f;
    if (c < D) {
        const float mean = pass ? a[c] / (float)M : 0.f;
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            const float v = o[(size_t)r * D + c];
            acc += pass ? (v - mean) * (v - mean) : v;
        }
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < D) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        atomicAdd((pass ? b : a) + c, s);
    }
}
__global__ void bn_finalize_kernel(float* mean, float* var, int M, int D, float momentum, float* rm, float* rv) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    const float m = mean[c] / (float)M, v = var[c] / (float)M;
    mean[c] = m;
    var[c] = v;
    if (rm) {
        rm[c] = (1.0f - momentum) * rm[c] + momentum * m;
        const float unbiased = M > 1 ? v * ((float)M / (float)(M - 1)) : v;
        rv[c] = (1.0f - momentum) * rv[c] + momentum * unbiased;
    }
}

// one block per video, thread per column (strided), loop over frames
__global__ void bn_apply_concat_kernel(const float* __restrict__ o, const float* __restrict__ mean,
                                       const float* __restrict__ var, const float* __restrict__ w,
                                       const float* __restrict__ bias, float eps, int F, int D, int E, int slot,
                                       float inv_fm, int accumulate, float* enc_hidden, float* enc_out,
                                       uint16_t* enc_hi, uint16_t* enc_lo) {
    const int b = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float mu = 0.f, istd = 1.f, ww = 1.f, bb = 0.f;
        if (mean) {
            mu = mean[d];
            istd = 1.0f / sqrtf(var[d] + eps);
            ww = w ? w[d] : 1.f;
            bb = bias ? bias[d] : 0.f;
        }
        float hsum = 0.f;
        for (int f = 0; f < F; ++f) {
            const float ov = o[((size_t)b * F + f) * D + d];
            hsum += ov;
            const float v = mean ? ((ov - mu) * istd * ww + bb) : ov;
            const size_t oi = ((size_t)b * E + (size_t)slot * F + f) * D + d;
            enc_out[oi] = v;
            if (enc_hi) {
                uint16_t h, l;
                split_bf16(v, h, l);
                enc_hi[oi] = h;
                if (enc_lo) enc_lo[oi] = l;
            }
        }
        if (enc_hidden) {
            const float hv = hsum * inv_fm;
            const size_t hi_ = (size_t)b * D + d;
            enc_hidden[hi_] = accumulate ? enc_hidden[hi_] + hv : hv;
        }
    }
}

// s1[c] += sum dy ; s2[c] += sum dy * xhat      (rows = (b, f) pairs of this modality slot)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ d_enc_out, const float* __restrict__ o,
                                     const float* __restrict__ mean, const float* __restrict__ var, float eps, int M,
                                     int F, int D, int E, int slot, int rows_per_block, float* s1, float* s2) {
    __shared__ float red1[8][32], red2[8][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float a1 = 0.f, a2 = 0.f;
    if (c < D) {
        const float mu = mean[c], istd = 1.0f / sqrtf(var[c] + eps);
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            const int b = r / F, f = r - b * F;
            const float dy = d_enc_out[((size_t)b * E + (size_t)slot * F + f) * D + c];
            const float xh = (o[(size_t)r * D + c] - mu) * istd;
            a1 += dy;
            a2 += dy * xh;
        }
    }
    red1[threadIdx.y][threadIdx.x] = a1;
    red2[threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.y == 0 && c < D) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t1 += red1[k][threadIdx.x]; t2 += red2[k][threadIdx.x]; }
        atomicAdd(s1 + c, t1);
        atomicAdd(s2 + c, t2);
    }
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ d_enc_out, const float* __restrict__ d_enc_hidden,
                                    const float* __restrict__ o, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ w, float eps, int M,
                                    int F, int D, int E, int slot, float inv_fm, const float* __restrict__ s1,
                                    const float* __restrict__ s2, float* __restrict__ d_o) {
    const int64_t n = (int64_t)M * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / D), c = (int)(i - (int64_t)r * D);
        const int b = r / F, f = r - b * F;
        const float dy = d_enc_out[((size_t)b * E + (size_t)slot * F + f) * D + c];
        float g;
        if (mean) {
            const float istd = 1.0f / sqrtf(var[c] + eps);
            const float xh = (o[i] - mean[c]) * istd;
            g = (w ? w[c] : 1.f) * istd * (dy - s1[c] / (float)M - xh * s2[c] / (float)M);
        } else {
            g = dy;
        }
        if (d_enc_hidden) g += d_enc_hidden[(size_t)b * D + c] * inv_fm;
        d_o[i] = g;
    }
}

__global__ void mean_bwd_kernel(const float* __restrict__ d_mean, int E, int D, float inv_e, int64_t n,
                                float* __restrict__ d_enc_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / ((int64_t)E * D);
        const int c = (int)(i % D);
        d_enc_out[i] += d_mean[b * D + c] * inv_e;
    }
}

// ---- log-softmax backward -------------------------------------------------------------------------------
__global__ void log_softmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ logp, int V, int ld_in,
                                       float* __restrict__ out, int ld_out, const int32_t* __restrict__ rowmap) {
    __shared__ float red[32];
    const size_t src = rowmap ? (size_t)rowmap[blockIdx.x] : (size_t)blockIdx.x;   // packed rows: g / logp live in the padded layout
    const float* gr = g + src * ld_in;
    const float* lr = logp + src * ld_in;
    float* orow = out + (size_t)blockIdx.x * ld_out;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) s += gr[i];
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = (lane < nw) ? red[lane] : 0.f;
    s = warp_sum(s);
    for (int i = threadIdx.x; i < ld_out; i += blockDim.x)
        orow[i] = (i < V) ? (gr[i] - (s != 0.f ? expf(lr[i]) * s : 0.f)) : 0.f;
}

// ---- LayerNorm backward ------------------------------------------------------------------------------------
constexpr int kLnChunks = 8;  // D <= 1024, one warp per row, lane owns float4 chunks lane + 32 t

// Given the row x (pre-LN) and dy in registers, computes dx in place of dy and accumulates the
// per-column dw/db partials of this warp into shared memory (block-level, flushed by the caller).
__device__ __forceinline__ void warp_ln_bwd(float4 (&xv)[kLnChunks], float4 (&dy)[kLnChunks], int nchunk, int D, int lane,
                                            const float* __restrict__ w, float eps, float* sm_dw, float* sm_db) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kLnChunks; ++t)
        if (t < nchunk && lane + 32 * t < D / 4) s += xv[t].x + xv[t].y + xv[t].z + xv[t].w;
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < kLnChunks; ++t)
        if (t < nchunk && lane + 32 * t < D / 4) {
            const float a = xv[t].x - mean, b = xv[t].y - mean, c = xv[t].z - mean, d = xv[t].w - mean;
            q += a * a + b * b + c * c + d * d;
        }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int t = 0; t < kLnChunks; ++t) {
        const int c = lane + 32 * t;
        if (t < nchunk && c < D / 4) {
            const float4 ww = *reinterpret_cast<const float4*>(w + c * 4);
            float4 xh = make_float4((xv[t].x - mean) * rstd, (xv[t].y - mean) * rstd, (xv[t].z - mean) * rstd,
                                    (xv[t].w - mean) * rstd);
            atomicAdd(sm_dw + c * 4 + 0, dy[t].x * xh.x); atomicAdd(sm_dw + c * 4 + 1, dy[t].y * xh.y);
            atomicAdd(sm_dw + c * 4 + 2, dy[t].z * xh.z); atomicAdd(sm_dw + c * 4 + 3, dy[t].w * xh.w);
            atomicAdd(sm_db + c * 4 + 0, dy[t].x); atomicAdd(sm_db + c * 4 + 1, dy[t].y);
            atomicAdd(sm_db + c * 4 + 2, dy[t].z); atomicAdd(sm_db + c * 4 + 3, dy[t].w);
            const float4 gx = make_float4(dy[t].x * ww.x, dy[t].y * ww.y, dy[t].z * ww.z, dy[t].w * ww.w);
            s1 += gx.x + gx.y + gx.z + gx.w;
            s2 += gx.x * xh.x + gx.y * xh.y + gx.z * xh.z + gx.w * xh.w;
            dy[t] = gx;
            xv[t] = xh;
        }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int t = 0; t < kLnChunks; ++t) {
        if (t < nchunk && lane + 32 * t < D / 4) {
            dy[t].x = rstd * (dy[t].x - s1 - xv[t].x * s2);
            dy[t].y = rstd * (dy[t].y - s1 - xv[t].y * s2);
            dy[t].z = rstd * (dy[t].z - s1 - xv[t].z * s2);
            dy[t].w = rstd * (dy[t].w - s1 - xv[t].w * s2);
        }
    }
}

// blockDim = 256 (8 warps, one row each per iteration); dynamic smem 2*D floats
__global__ void layernorm_bwd_kernel(const float* __restrict__ dyp, const float* __restrict__ x,
                                     const float* __restrict__ w, float eps, const int64_t* __restrict__ row_tokens,
                                     int R, int D, float* dx, float* dw, float* db) {
    extern __shared__ float sm[];
    float* sm_dw = sm;
    float* sm_db = sm + D;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int nchunk = (D / 4 + 31) / 32;
    for (int row = blockIdx.x * wpb + wib; row < R; row += gridDim.x * wpb) {
        const bool zero = row_tokens ? (row_tokens[row] == NAVC_PAD) : false;
        float4 xv[kLnChunks], dy[kLnChunks];
#pragma unroll
        for (int t = 0; t < kLnChunks; ++t) {
            const int c = lane + 32 * t;
            if (t < nchunk && c < D / 4) {
                xv[t] = *reinterpret_cast<const float4*>(x + (size_t)row * D + c * 4);
                dy[t] = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dyp + (size_t)row * D + c * 4);
            }
        }
        warp_ln_bwd(xv, dy, nchunk, D, lane, w, eps, sm_dw, sm_db);
#pragma unroll
        for (int t = 0; t < kLnChunks; ++t) {
            const int c = lane + 32 * t;
            if (t < nchunk && c < D / 4) *reinterpret_cast<float4*>(dx + (size_t)row * D + c * 4) = dy[t];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        atomicAdd(dw + i, sm_dw[i]);
        atomicAdd(db + i, sm_db[i]);
    }
}

__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
    atomicAdd(p + 0, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}

__global__ void embed_ln_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ tokens,
                                    const int64_t* __restrict__ category, const float* __restrict__ word,
                                    const float* __restrict__ pos, const float* __restrict__ cat,
                                    const float* __restrict__ extra, int group, const float* __restrict__ lw, float eps,
                                    int R, int S, int D, float* d_word, float* d_pos, float* d_cat, float* d_extra,
                                    float* d_lw, float* d_lb, const int32_t* __restrict__ rowmap) {
    extern __shared__ float sm[];
    float* sm_dw = sm;
    float* sm_db = sm + D;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int nchunk = (D / 4 + 31) / 32;
    for (int row = blockIdx.x * wpb + wib; row < R; row += gridDim.x * wpb) {
        const int src = rowmap ? rowmap[row] : row;   // packed rows: row -> padded position
        const int n = src / S, s = src % S;
        const int64_t tok = tokens[src];
        const float* wr = word + (size_t)tok * D;
        const float* pr = pos + (size_t)s * D;
        const int64_t ci = cat ? category[n / group] : 0;
        const float* cr = cat ? cat + (size_t)ci * D : nullptr;
        const float* er = extra ? extra + (size_t)(n / group) * D : nullptr;
        float4 xv[kLnChunks], dy[kLnChunks];
#pragma unroll
        for (int t = 0; t < kLnChunks; ++t) {
            const int c = lane + 32 * t;
            if (t < nchunk && c < D / 4) {
                float4 a = add4(*reinterpret_cast<const float4*>(wr + c * 4), *reinterpret_cast<const float4*>(pr + c * 4));
                if (cr) a = add4(a, *reinterpret_cast<const float4*>(cr + c * 4));
                if (er) a = add4(a, *reinterpret_cast<const float4*>(er + c * 4));
                xv[t] = a;
                dy[t] = *reinterpret_cast<const float4*>(dout + (size_t)row * D + c * 4);
            }
        }
        warp_ln_bwd(xv, dy, nchunk, D, lane, lw, eps, sm_dw, sm_db);
#pragma unroll
        for (int t = 0; t < kLnChunks; ++t) {
            const int c = lane + 32 * t;
            if (t < nchunk && c < D / 4) {
                if (tok != NAVC_PAD) atomic_add4(d_word + (size_t)tok * D + c * 4, dy[t]);  // padding_idx row: no gradient
                atomic_add4(d_pos + (size_t)s * D + c * 4, dy[t]);
                if (d_cat) atomic_add4(d_cat + (size_t)ci * D + c * 4, dy[t]);
                if (d_extra) atomic_add4(d_extra + (size_t)(n / group) * D + c * 4, dy[t]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        atomicAdd(d_lw + i, sm_dw[i]);
        atomicAdd(d_lb + i, sm_db[i]);
    }
}

// ---- fused cross-entropy (seq2seq.py:102-103 + misc/crit.py:62-84 without the [rows, V] log-prob tensor) ----
// one thread per row: combine the vocabulary partials of navc_vocab_partials_* into lse / nll / argmax
__global__ void ce_stats_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const int32_t* __restrict__ pi,
                                int nt, const float* __restrict__ tl, const int64_t* __restrict__ labels, int R,
                                float* __restrict__ lse, float* __restrict__ nll, int32_t* __restrict__ arg) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float m = -INFINITY;
    int idx = 0x7fffffff;
    for (int t = 0; t < nt; ++t) {
        const float v = pm[(size_t)r * nt + t];
        const int i = pi[(size_t)r * nt + t];
        if (v > m || (v == m && i < idx)) { m = v; idx = i; }
    }
    float s = 0.f;
    for (int t = 0; t < nt; ++t) {
        const float q = ps[(size_t)r * nt + t];
        if (q != 0.f) s += q * expf(pm[(size_t)r * nt + t] - m);
    }
    const float l = m + logf(s);
    lse[r] = l;
    nll[r] = (labels[r] == NAVC_PAD) ? 0.f : l - tl[r];
    arg[r] = idx;
}
// block per row, in place: logits -> (softmax - onehot(label)) * scale[0] * (label != PAD); columns V..ld-1 zero
__global__ void ce_grad_kernel(float* __restrict__ logits, const float* __restrict__ lse,
                               const int64_t* __restrict__ labels, const float* __restrict__ scale, int V, int ld) {
    float* row = logits + (size_t)blockIdx.x * ld;
    const int64_t lab = labels[blockIdx.x];
    const float g = (lab == NAVC_PAD) ? 0.f : scale[0];
    const float l = lse[blockIdx.x];
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        float v = 0.f;
        if (i < V && g != 0.f) v = (expf(row[i] - l) - ((int64_t)i == lab ? 1.f : 0.f)) * g;
        row[i] = v;
    }
}

// ---- fused clip + Adam (misc/run.py:260-261, misc/optim.py:61-62: torch.optim.Adam with L2 weight decay) ----
__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float clip, float lr_over_bc1, float b1, float b2,
                                 float eps, float wd, float inv_sqrt_bc2) {
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
        if (i + 4 <= n) {
            float4 pp = *reinterpret_cast<const float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
            float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gr = ga[k];
                if (clip > 0.f) gr = fminf(fmaxf(gr, -clip), clip);       // clip_grad_value_
                gr = fmaf(wd, pa[k], gr);                                  // L2 weight decay
                ma[k] = fmaf(1.0f - b1, gr - ma[k], ma[k]);                // exp_avg.lerp_(grad, 1 - beta1)
                va[k] = fmaf(1.0f - b2, gr * gr, b2 * va[k]);
                const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
                pa[k] -= lr_over_bc1 * (ma[k] / denom);
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (int64_t j = i; j < n; ++j) {
                float gr = g[j];
                if (clip > 0.f) gr = fminf(fmaxf(gr, -clip), clip);
                gr = fmaf(wd, p[j], gr);
                m[j] = fmaf(1.0f - b1, gr - m[j], m[j]);
                v[j] = fmaf(1.0f - b2, gr * gr, b2 * v[j]);
                p[j] -= lr_over_bc1 * (m[j] / (sqrtf(v[j]) * inv_sqrt_bc2 + eps));
            }
        }
    }
}

// ---- attention backward ----------------------------------------------------------------------------------------
// grid (G, H), 256 threads.  Block (g, h) owns the Sk key/value rows of owner g and its NQ query
// rows (self: NQ = Sk = S, owner = sequence; cross: NQ = group * S, owner = video), processed in
// chunks of QB = 32 queries.  All four contractions are register tiled 4 x 4 per thread:
//   A  scores S = Q K^T and dP = dO V^T   (K, V staged transposed: one LDS.128 feeds 4 keys)
//   B  softmax, dS = P * (dP - rowsum(P dP)) / sqrt(dk), zero at masked positions (warp per row)
//   C  dV = P^T dO, dK = dS^T Q           D  dQ = dS K
constexpr float kMaskFillB = -10e6f;  // models/bert.py:161
constexpr int kAttnBwdQB = 32;

template <int DK>
struct AttnBwdSmem {
    static constexpr int LD = DK + 4;  // row stride of the row-major tiles (float4 aligned, 4-bank skew)
    static __host__ __device__ int kp(int Sk) { return (Sk + 3) / 4 * 4; }          // padded key count
    static __host__ __device__ int sp(int Sk) { return kp(Sk) + 4; }                // P / dS row stride
    static __host__ __device__ int kpt(int Sk) { return kp(Sk) + 4; }               // K^T / V^T row stride (bank skew for phase D)
    static __host__ __device__ size_t floats(int Sk) {
        // E = 120 keys: 112.6 KB -> two blocks per SM (a third, row-major copy of K used to cost the second block)
        return (size_t)2 * DK * kpt(Sk) + (size_t)2 * kAttnBwdQB * LD + (size_t)2 * kAttnBwdQB * sp(Sk);
    }
};

template <int DK>
__global__ void __launch_bounds__(256) attn_bwd_kernel(
    const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v, int ldkv,
    const int64_t* __restrict__ tokens, int NQ, int S, int Sk, int mask_kind, int watch,
    const float* __restrict__ d_ctx, int ld_dctx, float* __restrict__ dq, int ld_dq, float* __restrict__ dk,
    float* __restrict__ dv, int ld_dkv, const int32_t* __restrict__ seq_off, int packed_self) {
    using SM = AttnBwdSmem<DK>;
    constexpr int LD = SM::LD;
    constexpr int QB = kAttnBwdQB;
    extern __shared__ __align__(16) float sm[];
    const int SP = SM::sp(Sk), KPT = SM::kpt(Sk);
    float* Kt = sm;                         // [DK][KPT]  K transposed (columns >= sk zero)
    float* Vt = Kt + (size_t)DK * KPT;      // [DK][KPT]  V transposed
    float* Qs = Vt + (size_t)DK * KPT;      // [QB][LD]
    float* dOs = Qs + QB * LD;              // [QB][LD]
    float* Ps = dOs + QB * LD;              // [QB][SP]
    float* dSs = Ps + (size_t)QB * SP;      // [QB][SP]
    const int g = blockIdx.x, h = blockIdx.y;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;

    // packed rows (seq_off != NULL): the query rows of owner g are rows [seq_off[g], seq_off[g+1]) of q / d_ctx / dq;
    // for self-attention the keys are those same rows (sk = len), for cross-attention the owner's Sk rows of kv.
    // Shared-memory layout always uses the maximum Sk; `sk` is the number of keys that exist.
    size_t q_row = (size_t)g * NQ, kv_row = (size_t)g * Sk;
    int sk = Sk;
    if (seq_off) {
        const int r0 = seq_off[g];
        NQ = seq_off[g + 1] - r0;
        q_row = (size_t)r0;
        if (packed_self) { kv_row = (size_t)r0; sk = NQ; }
    }
    const float* kb = k + kv_row * ldkv + h * DK;
    const float* vb = v + kv_row * ldkv + h * DK;
    const int KE = SM::kp(sk);  // keys that are looped over (sk rounded up to the 4-key register tile)
    for (int idx = tid; idx < KE * (DK / 4); idx += nthr) {
        const int j = idx / (DK / 4), d4 = idx - j * (DK / 4);
        float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
        if (j < sk) {
            kk = *reinterpret_cast<const float4*>(kb + (size_t)j * ldkv + d4 * 4);
            vv = *reinterpret_cast<const float4*>(vb + (size_t)j * ldkv + d4 * 4);
        }
        Kt[(d4 * 4 + 0) * KPT + j] = kk.x; Kt[(d4 * 4 + 1) * KPT + j] = kk.y;
        Kt[(d4 * 4 + 2) * KPT + j] = kk.z; Kt[(d4 * 4 + 3) * KPT + j] = kk.w;
        Vt[(d4 * 4 + 0) * KPT + j] = vv.x; Vt[(d4 * 4 + 1) * KPT + j] = vv.y;
        Vt[(d4 * 4 + 2) * KPT + j] = vv.z; Vt[(d4 * 4 + 3) * KPT + j] = vv.w;
    }
    const float inv_sqrt = 1.0f / sqrtf((float)DK);
    const float sqrt_dk = sqrtf((float)DK);
    const bool use_watch = (mask_kind == NAVC_MASK_CAUSAL) && watch != 0 && S >= watch;
    const int64_t* trow = tokens ? tokens + (size_t)g * S : nullptr;
    const int jt_n = KE / 4;

    for (int q0 = 0; q0 < NQ; q0 += QB) {
        const int nq = min(QB, NQ - q0);
        const int nq4 = (nq + 3) & ~3;  // query rows that are looped over (packed rows: len is ~half of QB on average)
        __syncthreads();
        const float* qb = q + (q_row + q0) * ldq + h * DK;
        const float* ob = d_ctx + (q_row + q0) * ld_dctx + h * DK;
        for (int idx = tid; idx < QB * (DK / 4); idx += nthr) {
            const int i = idx / (DK / 4), d4 = idx - i * (DK / 4);
            float4 qq = make_float4(0.f, 0.f, 0.f, 0.f), oo = qq;
            if (i < nq) {
                qq = *reinterpret_cast<const float4*>(qb + (size_t)i * ldq + d4 * 4);
                oo = *reinterpret_cast<const float4*>(ob + (size_t)i * ld_dctx + d4 * 4);
            }
            *reinterpret_cast<float4*>(Qs + i * LD + d4 * 4) = qq;
            *reinterpret_cast<float4*>(dOs + i * LD + d4 * 4) = oo;
        }
        __syncthreads();
        // ---- A: 4 queries x 4 keys per thread ----
        for (int t = tid; t < (nq4 / 4) * jt_n; t += nthr) {
            const int ti = t / jt_n, tj = t - ti * jt_n;
            float sacc[4][4], pacc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) { sacc[a][b] = 0.f; pacc[a][b] = 0.f; }
            const float* qr = Qs + (ti * 4) * LD;
            const float* orw = dOs + (ti * 4) * LD;
#pragma unroll 4
            for (int d = 0; d < DK; ++d) {
                const float4 kk = *reinterpret_cast<const float4*>(Kt + d * KPT + tj * 4);
                const float4 vv = *reinterpret_cast<const float4*>(Vt + d * KPT + tj * 4);
                const float kf[4] = {kk.x, kk.y, kk.z, kk.w}, vf[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float qd = qr[a * LD + d], od = orw[a * LD + d];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        sacc[a][b] = fmaf(qd, kf[b], sacc[a][b]);
                        pacc[a][b] = fmaf(od, vf[b], pacc[a][b]);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                *reinterpret_cast<float4*>(Ps + (ti * 4 + a) * SP + tj * 4) =
                    make_float4(sacc[a][0] / sqrt_dk, sacc[a][1] / sqrt_dk, sacc[a][2] / sqrt_dk, sacc[a][3] / sqrt_dk);
                *reinterpret_cast<float4*>(dSs + (ti * 4 + a) * SP + tj * 4) =
                    make_float4(pacc[a][0], pacc[a][1], pacc[a][2], pacc[a][3]);
            }
        }
        __syncthreads();
        // ---- B: softmax + dS, one warp per query row (rows >= nq and keys >= Sk become zero) ----
        for (int i = warp; i < nq4; i += nw) {
            if (i >= nq) {
                for (int j = lane; j < KE; j += 32) { Ps[i * SP + j] = 0.f; dSs[i * SP + j] = 0.f; }
                continue;
            }
            const int ipos = (q0 + i) % S;
            float m = -INFINITY;
            for (int j = lane; j < sk; j += 32) {
                float sv = Ps[i * SP + j];
                if (trow) {
                    bool masked = trow[j] == NAVC_PAD;
                    if (mask_kind == NAVC_MASK_CAUSAL) masked = masked || (j > ipos) || (use_watch && j <= ipos - watch);
                    if (mask_kind == NAVC_MASK_SELF) masked = masked || (j == ipos);
                    if (masked) sv = kMaskFillB;
                }
                Ps[i * SP + j] = sv;
                m = fmaxf(m, sv);
            }
            m = warp_max(m);
            float sum = 0.f;
            for (int j = lane; j < sk; j += 32) {
                const float e = expf(Ps[i * SP + j] - m);
                Ps[i * SP + j] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            float dot = 0.f;
            for (int j = lane; j < sk; j += 32) {
                const float p = Ps[i * SP + j] / sum;
                Ps[i * SP + j] = p;
                dot += p * dSs[i * SP + j];
            }
            dot = warp_sum(dot);
            for (int j = lane; j < KE; j += 32) {
                if (j >= sk) {
                    Ps[i * SP + j] = 0.f;
                    dSs[i * SP + j] = 0.f;
                    continue;
                }
                bool masked = false;
                if (trow) {
                    masked = trow[j] == NAVC_PAD;
                    if (mask_kind == NAVC_MASK_CAUSAL) masked = masked || (j > ipos) || (use_watch && j <= ipos - watch);
                    if (mask_kind == NAVC_MASK_SELF) masked = masked || (j == ipos);
                }
                const float ds = Ps[i * SP + j] * (dSs[i * SP + j] - dot) * inv_sqrt;
                dSs[i * SP + j] = masked ? 0.f : ds;  // masked_fill: no gradient to the filled scores
            }
        }
        __syncthreads();
        // ---- C: dV / dK, 4 keys x 4 columns per thread; accumulated over query chunks in global memory ----
        float* dkb = dk + kv_row * ld_dkv + h * DK;
        float* dvb = dv + kv_row * ld_dkv + h * DK;
        for (int t = tid; t < jt_n * (DK / 4); t += nthr) {
            const int tj = t / (DK / 4), td = t - tj * (DK / 4);
            float va[4][4], ka[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) { va[a][b] = 0.f; ka[a][b] = 0.f; }
#pragma unroll 4
            for (int i = 0; i < nq4; ++i) {
                const float4 pp = *reinterpret_cast<const float4*>(Ps + i * SP + tj * 4);
                const float4 ss = *reinterpret_cast<const float4*>(dSs + i * SP + tj * 4);
                const float4 oo = *reinterpret_cast<const float4*>(dOs + i * LD + td * 4);
                const float4 qq = *reinterpret_cast<const float4*>(Qs + i * LD + td * 4);
                const float pf[4] = {pp.x, pp.y, pp.z, pp.w}, sf[4] = {ss.x, ss.y, ss.z, ss.w};
                const float of[4] = {oo.x, oo.y, oo.z, oo.w}, qf[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        va[a][b] = fmaf(pf[a], of[b], va[a][b]);
                        ka[a][b] = fmaf(sf[a], qf[b], ka[a][b]);
                    }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int j = tj * 4 + a;
                if (j < sk) {
                    float4* pk = reinterpret_cast<float4*>(dkb + (size_t)j * ld_dkv + td * 4);
                    float4* pv = reinterpret_cast<float4*>(dvb + (size_t)j * ld_dkv + td * 4);
                    float4 nk = make_float4(ka[a][0], ka[a][1], ka[a][2], ka[a][3]);
                    float4 nv = make_float4(va[a][0], va[a][1], va[a][2], va[a][3]);
                    if (q0 > 0) { nk = add4(nk, *pk); nv = add4(nv, *pv); }
                    *pk = nk;
                    *pv = nv;
                }
            }
        }
        // ---- D: dQ = dS K from K^T, 4 queries x 4 columns {td, td+16, ..} per thread (consecutive K^T rows across the
        // threads of a quarter-warp: conflict-free float4 loads with the KPT skew) ----
        float* dqb = dq + (q_row + q0) * ld_dq + h * DK;
        constexpr int CG = DK / 4;  // column groups = threads per query tile
        for (int t = tid; t < (nq4 / 4) * CG; t += nthr) {
            const int ti = t / CG, td = t - ti * CG;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
            const float* sr = dSs + (ti * 4) * SP;
#pragma unroll 2
            for (int j = 0; j < KE; j += 4) {
                float4 sv[4], kv4[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) sv[a] = *reinterpret_cast<const float4*>(sr + a * SP + j);
#pragma unroll
                for (int b = 0; b < 4; ++b) kv4[b] = *reinterpret_cast<const float4*>(Kt + (td + CG * b) * KPT + j);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        acc[a][b] = fmaf(sv[a].x, kv4[b].x, acc[a][b]);
                        acc[a][b] = fmaf(sv[a].y, kv4[b].y, acc[a][b]);
                        acc[a][b] = fmaf(sv[a].z, kv4[b].z, acc[a][b]);
                        acc[a][b] = fmaf(sv[a].w, kv4[b].w, acc[a][b]);
                    }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = ti * 4 + a;
                if (i < nq) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) dqb[(size_t)i * ld_dq + td + CG * b] = acc[a][b];
                }
            }
        }
    }
}

template <int DK>
static int launch_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const int64_t* tokens,
                           int G, int NQ, int S, int Sk, int H, int mask_kind, int watch, const float* d_ctx, int ld_dctx,
                           float* dq, int ld_dq, float* dk, float* dv, int ld_dkv, cudaStream_t st, const char* what,
                           const int32_t* seq_off = nullptr, int packed_self = 0) {
    const size_t smem = AttnBwdSmem<DK>::floats(Sk) * sizeof(float);
    NAVC_REQUIRE(smem <= 227 * 1024, "%s: Sk=%d too large for shared memory", what, Sk);
    NAVC_REQUIRE(ldq % 4 == 0 && ldkv % 4 == 0 && ld_dctx % 4 == 0 && ld_dq % 4 == 0 && ld_dkv % 4 == 0 &&
                     ((((uintptr_t)q) | ((uintptr_t)k) | ((uintptr_t)v) | ((uintptr_t)d_ctx) | ((uintptr_t)dq) |
                       ((uintptr_t)dk) | ((uintptr_t)dv)) & 15) == 0,
                 "%s: operands must be 16-byte aligned with leading dimensions multiple of 4", what);
    auto kern = attn_bwd_kernel<DK>;
    if (smem > 48 * 1024) NAVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(G, H), 256, smem, st>>>(q, ldq, k, v, ldkv, tokens, NQ, S, Sk, mask_kind, watch, d_ctx, ld_dctx, dq, ld_dq,
                                        dk, dv, ld_dkv, seq_off, packed_self);
    return check_launch(what);
}

static int dispatch_attn_bwd(int dkk, const float* q, int ldq, const float* k, const float* v, int ldkv,
                             const int64_t* tokens, int G, int NQ, int S, int Sk, int H, int mask_kind, int watch,
                             const float* d_ctx, int ld_dctx, float* dq, int ld_dq, float* dk, float* dv, int ld_dkv,
                             cudaStream_t st, const char* what, const int32_t* seq_off = nullptr, int packed_self = 0) {
#define NAVC_ATTN_BWD(DKV) \
    if (dkk == DKV) return launch_attn_bwd<DKV>(q, ldq, k, v, ldkv, tokens, G, NQ, S, Sk, H, mask_kind, watch, d_ctx, ld_dctx, dq, ld_dq, dk, dv, ld_dkv, st, what, seq_off, packed_self)
    NAVC_ATTN_BWD(64);
    NAVC_ATTN_BWD(32);
    NAVC_ATTN_BWD(16);
    NAVC_ATTN_BWD(8);
#undef NAVC_ATTN_BWD
    set_error("%s: head size %d unsupported (8/16/32/64)", what, dkk);
    return 1;
}

}  // namespace navc

using namespace navc;

extern "C" int navc_drop_add(const float* y, const float* res, uint64_t seed1, float p1, uint64_t seed2, float p2,
                             const int64_t* row_tokens, int M, int D, float* out_f32, uint16_t* out_hi,
                             uint16_t* out_lo, void* stream) {
    NAVC_REQUIRE(y && (out_f32 || out_hi) && M > 0 && D > 0 && D % 4 == 0, "navc_drop_add: bad arguments");
    NAVC_REQUIRE(p1 >= 0.f && p1 < 1.f && p2 >= 0.f && p2 < 1.f, "navc_drop_add: p must be in [0,1)");
    const int64_t n4 = (int64_t)M * D / 4;
    drop_add_kernel<<<ew_blocks(n4, 256), 256, 0, as_stream(stream)>>>(y, res, make_drop(seed1, p1), make_drop(seed2, p2),
                                                                     row_tokens, n4, D, out_f32, out_hi, out_lo);
    return check_launch("navc_drop_add");
}

extern "C" int navc_drop_add_bwd(const float* dout, uint64_t seed1, float p1, uint64_t seed2, float p2,
                                 const int64_t* row_tokens, int M, int D, float* d_y, float* d_res, void* stream) {
    NAVC_REQUIRE(dout && (d_y || d_res) && M > 0 && D > 0 && D % 4 == 0, "navc_drop_add_bwd: bad arguments");
    const int64_t n4 = (int64_t)M * D / 4;
    drop_add_bwd_kernel<<<ew_blocks(n4, 256), 256, 0, as_stream(stream)>>>(dout, make_drop(seed1, p1), make_drop(seed2, p2),
                                                                         row_tokens, n4, D, d_y, d_res);
    return check_launch("navc_drop_add_bwd");
}

extern "C" int navc_drop_add_bwd_split(const float* dout, uint64_t seed1, float p1, uint64_t seed2, float p2,
                                       const int64_t* row_tokens, int M, int D, float* d_res, uint16_t* dy_hi, uint16_t* dy_lo,
                                       int ld_s, float* colsum, void* stream) {
    NAVC_REQUIRE(dout && dy_hi && M > 0 && D > 0 && D % 4 == 0 && ld_s >= D && ld_s % 4 == 0, "navc_drop_add_bwd_split: bad arguments");
    NAVC_REQUIRE(((((uintptr_t)dout) | ((uintptr_t)d_res)) & 15) == 0 && ((((uintptr_t)dy_hi) | ((uintptr_t)dy_lo)) & 7) == 0,
                 "navc_drop_add_bwd_split: operands must be 16-byte (fp32) / 8-byte (bf16) aligned");
    dim3 g((D + 127) / 128, (M + 63) / 64);
    NAVC_REQUIRE(g.y <= 65535, "navc_drop_add_bwd_split: M too large");
    drop_add_bwd_split_kernel<<<g, dim3(32, 8), 0, as_stream(stream)>>>(dout, make_drop(seed1, p1), make_drop(seed2, p2), row_tokens, M, D,
                                                                       d_res, dy_hi, dy_lo, ld_s, colsum);
    return check_launch("navc_drop_add_bwd_split");
}

extern "C" int navc_act_drop_bwd_split(const float* dout, const float* u, int act, uint64_t seed, float p, int M, int N,
                                       uint16_t* dy_hi, uint16_t* dy_lo, int ld_s, float* colsum, void* stream) {
    NAVC_REQUIRE(dout && u && dy_hi && M > 0 && N > 0 && N % 4 == 0 && ld_s >= N && ld_s % 4 == 0, "navc_act_drop_bwd_split: bad arguments");
    NAVC_REQUIRE(((((uintptr_t)dout) | ((uintptr_t)u)) & 15) == 0 && ((((uintptr_t)dy_hi) | ((uintptr_t)dy_lo)) & 7) == 0,
                 "navc_act_drop_bwd_split: operands must be 16-byte (fp32) / 8-byte (bf16) aligned");
    dim3 g((N + 127) / 128, (M + 63) / 64);
    NAVC_REQUIRE(g.y <= 65535, "navc_act_drop_bwd_split: M too large");
    act_drop_bwd_split_kernel<<<g, dim3(32, 8), 0, as_stream(stream)>>>(dout, u, act, make_drop(seed, p), M, N, dy_hi, dy_lo, ld_s, colsum);
    return check_launch("navc_act_drop_bwd_split");
}

extern "C" int navc_act_drop(const float* u, int act, uint64_t seed, float p, int64_t n, float* out_f32,
                             uint16_t* out_hi, uint16_t* out_lo, void* stream) {
    NAVC_REQUIRE(u && (out_f32 || out_hi) && n > 0 && p >= 0.f && p < 1.f, "navc_act_drop: bad arguments");
    act_drop_kernel<<<ew_blocks((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(u, act, make_drop(seed, p), n, out_f32,
                                                                              out_hi, out_lo);
    return check_launch("navc_act_drop");
}

extern "C" int navc_act_drop_bwd(const float* dout, const float* u, int act, uint64_t seed, float p, int64_t n,
                                 float* du, void* stream) {
    NAVC_REQUIRE(dout && u && du && n > 0, "navc_act_drop_bwd: bad arguments");
    act_drop_bwd_kernel<<<ew_blocks(n, 256), 256, 0, as_stream(stream)>>>(dout, u, act, make_drop(seed, p), n, du);
    return check_launch("navc_act_drop_bwd");
}

extern "C" int navc_transpose_pack(const float* x, int M, int N, int ld, uint16_t* hi, uint16_t* lo, int ld_s,
                                   float* t_f32, uint16_t* t_hi, uint16_t* t_lo, int ld_t, float* colsum,
                                   void* stream) {
    NAVC_REQUIRE(x && M > 0 && N > 0 && ld >= N, "navc_transpose_pack: bad arguments");
    NAVC_REQUIRE(!(t_f32 || t_hi) || ld_t >= M, "navc_transpose_pack: ld_t < M");
    NAVC_REQUIRE(!hi || ld_s >= N, "navc_transpose_pack: ld_s < N");
    if (hi && !t_f32 && !t_hi && ld % 4 == 0 && ld_s % 4 == 0 && (((uintptr_t)x) & 15) == 0 &&
        ((((uintptr_t)hi) | ((uintptr_t)lo)) & 7) == 0) {
        dim3 g((ld_s + 127) / 128, (M + 63) / 64);
        NAVC_REQUIRE(g.y <= 65535, "navc_transpose_pack: M too large");
        split_colsum_kernel<<<g, dim3(32, 8), 0, as_stream(stream)>>>(x, M, N, ld, hi, lo, ld_s, colsum);
        return check_launch("navc_transpose_pack");
    }
    const int m_ext = (t_f32 || t_hi) && ld_t > M ? ld_t : M;  // tiles also cover the zero-filled pad columns
    const int n_ext = hi && ld_s > N ? ld_s : N;
    dim3 grid((n_ext + 31) / 32, (m_ext + 31) / 32);
    NAVC_REQUIRE(grid.y <= 65535, "navc_transpose_pack: M too large");
    transpose_pack_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, M, N, ld, hi, lo, ld_s, t_f32, t_hi, t_lo, ld_t,
                                                                      colsum);
    return check_launch("navc_transpose_pack");
}

extern "C" int navc_highway_fwd_train(const float* x, const float* yg, int gate, int BF, int D, uint64_t seed, float p,
                                      float* o, void* stream) {
    NAVC_REQUIRE(x && yg && o && BF > 0 && D > 0 && p >= 0.f && p < 1.f, "navc_highway_fwd_train: bad arguments");
    const int64_t n = (int64_t)BF * D;
    highway_fwd_train_kernel<<<ew_blocks(n, 256), 256, 0, as_stream(stream)>>>(x, yg, gate, n, D, make_drop(seed, p), o);
    return check_launch("navc_highway_fwd_train");
}

extern "C" int navc_highway_bwd(const float* d_o, const float* x, const float* yg, int gate, int BF, int D,
                                uint64_t seed, float p, float* d_x, float* d_yg, void* stream) {
    NAVC_REQUIRE(d_o && x && yg && d_x && d_yg && BF > 0 && D > 0, "navc_highway_bwd: bad arguments");
    const int64_t n = (int64_t)BF * D;
    highway_bwd_kernel<<<ew_blocks(n, 256), 256, 0, as_stream(stream)>>>(d_o, x, yg, gate, n, D, make_drop(seed, p), d_x, d_yg);
    return check_launch("navc_highway_bwd");
}

extern "C" int navc_bn_stats(const float* o, int M, int D, float* mean, float* var, float momentum,
                             float* running_mean, float* running_var, void* stream) {
    NAVC_REQUIRE(o && mean && var && M > 0 && D > 0, "navc_bn_stats: bad arguments");
    NAVC_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "navc_bn_stats: need both running stats");
    cudaStream_t st = as_stream(stream);
    NAVC_CUDA(cudaMemsetAsync(mean, 0, sizeof(float) * D, st));
    NAVC_CUDA(cudaMemsetAsync(var, 0, sizeof(float) * D, st));
    const int rpb = 256;
    dim3 grid((D + 31) / 32, (M + rpb - 1) / rpb);
    bn_colreduce_kernel<<<grid, dim3(32, 8), 0, st>>>(o, M, D, rpb, 0, mean, var);
    bn_colreduce_kernel<<<grid, dim3(32, 8), 0, st>>>(o, M, D, rpb, 1, mean, var);
    bn_finalize_kernel<<<(D + 255) / 256, 256, 0, st>>>(mean, var, M, D, momentum, running_mean, running_var);
    return check_launch("navc_bn_stats");
}

extern "C" int navc_bn_apply_concat(const float* o, const float* mean, const float* var, const float* w,
                                    const float* b, float eps, int B, int F, int D, int E, int slot, int n_modalities,
                                    int accumulate, float* enc_hidden, float* enc_out, uint16_t* enc_hi,
                                    uint16_t* enc_lo, void* stream) {
    NAVC_REQUIRE(o && enc_out && B > 0 && F > 0 && D > 0 && (slot + 1) * F <= E, "navc_bn_apply_concat: bad arguments");
    NAVC_REQUIRE((mean == nullptr) == (var == nullptr), "navc_bn_apply_concat: need both statistics");
    const int threads = D >= 512 ? 512 : ((D + 31) / 32) * 32;
    bn_apply_concat_kernel<<<B, threads, 0, as_stream(stream)>>>(o, mean, var, w, b, eps, F, D, E, slot,
                                                                 1.0f / ((float)F * (float)n_modalities), accumulate,
                                                                 enc_hidden, enc_out, enc_hi, enc_lo);
    return check_launch("navc_bn_apply_concat");
}

extern "C" int navc_bn_bwd(const float* d_enc_out, const float* d_enc_hidden, const float* o, const float* mean,
                           const float* var, const float* w, float eps, int B, int F, int D, int E, int slot,
                           int n_modalities, float* d_w, float* d_b, float* d_o, void* stream) {
    NAVC_REQUIRE(d_enc_out && d_o && B > 0 && F > 0 && D > 0 && (slot + 1) * F <= E, "navc_bn_bwd: bad arguments");
    NAVC_REQUIRE(!mean || (o && var && d_w && d_b), "navc_bn_bwd: missing statistics / gradient buffers");
    cudaStream_t st = as_stream(stream);
    const int M = B * F;
    if (mean) {
        NAVC_CUDA(cudaMemsetAsync(d_w, 0, sizeof(float) * D, st));
        NAVC_CUDA(cudaMemsetAsync(d_b, 0, sizeof(float) * D, st));
        const int rpb = 256;
        dim3 grid((D + 31) / 32, (M + rpb - 1) / rpb);
        bn_bwd_reduce_kernel<<<grid, dim3(32, 8), 0, st>>>(d_enc_out, o, mean, var, eps, M, F, D, E, slot, rpb, d_b, d_w);
    }
    bn_bwd_apply_kernel<<<ew_blocks((int64_t)M * D, 256), 256, 0, st>>>(d_enc_out, d_enc_hidden, o, mean, var, w, eps, M, F,
                                                                      D, E, slot, 1.0f / ((float)F * (float)n_modalities),
                                                                      d_b, d_w, d_o);
    return check_launch("navc_bn_bwd");
}

extern "C" int navc_mean_bwd(const float* d_mean, int B, int E, int D, float* d_enc_out, void* stream) {
    NAVC_REQUIRE(d_mean && d_enc_out && B > 0 && E > 0 && D > 0, "navc_mean_bwd: bad arguments");
    const int64_t n = (int64_t)B * E * D;
    mean_bwd_kernel<<<ew_blocks(n, 256), 256, 0, as_stream(stream)>>>(d_mean, E, D, 1.0f / (float)E, n, d_enc_out);
    return check_launch("navc_mean_bwd");
}

extern "C" int navc_log_softmax_bwd(const float* g, const float* logp, int M, int V, int ld_in, float* dlogits,
                                    int ld_out, void* stream) {
    NAVC_REQUIRE(g && logp && dlogits && M > 0 && V > 0 && ld_in >= V && ld_out >= V, "navc_log_softmax_bwd: bad arguments");
    log_softmax_bwd_kernel<<<M, 256, 0, as_stream(stream)>>>(g, logp, V, ld_in, dlogits, ld_out, nullptr);
    return check_launch("navc_log_softmax_bwd");
}

extern "C" int navc_log_softmax_bwd_rows(const float* g, const float* logp, const int32_t* rowmap, int rows, int V, int ld_in,
                                         float* dlogits, int ld_out, void* stream) {
    NAVC_REQUIRE(g && logp && rowmap && dlogits && rows > 0 && V > 0 && ld_in >= V && ld_out >= V,
                 "navc_log_softmax_bwd_rows: bad arguments");
    log_softmax_bwd_kernel<<<rows, 256, 0, as_stream(stream)>>>(g, logp, V, ld_in, dlogits, ld_out, rowmap);
    return check_launch("navc_log_softmax_bwd_rows");
}

// out[rowmap[i], :] = in[i, :] (scatter) or out[i, :] = in[rowmap[i], :] (gather), fp32 rows of D % 4 == 0 floats
__global__ void rows_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int D, const int32_t* __restrict__ rowmap,
                                int rows, int scatter) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= rows) return;
    const size_t m = (size_t)rowmap[i];
    const float* src = in + (scatter ? (size_t)i : m) * D;
    float* dst = out + (scatter ? m : (size_t)i) * D;
    for (int c = (threadIdx.x & 31) * 4; c < D; c += 128) *reinterpret_cast<float4*>(dst + c) = *reinterpret_cast<const float4*>(src + c);
}

extern "C" int navc_rows_f32(const float* in, float* out, int D, const int32_t* rowmap, int rows, int scatter, void* stream) {
    NAVC_REQUIRE(in && out && rowmap && rows > 0 && D % 4 == 0, "navc_rows_f32: bad arguments");
    rows_f32_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(in, out, D, rowmap, rows, scatter);
    return check_launch("navc_rows_f32");
}

extern "C" int navc_layernorm_bwd(const float* dy, const float* x, const float* w, float eps, const int64_t* row_tokens,
                                  int M, int D, float* dx, float* dw, float* db, void* stream) {
    NAVC_REQUIRE(dy && x && w && dx && dw && db && M > 0, "navc_layernorm_bwd: bad arguments");
    NAVC_REQUIRE(D % 4 == 0 && D <= 4 * 32 * kLnChunks, "navc_layernorm_bwd: D must be a multiple of 4 and <= 1024");
    int blocks = (M + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    layernorm_bwd_kernel<<<blocks, 256, 2 * D * sizeof(float), as_stream(stream)>>>(dy, x, w, eps, row_tokens, M, D, dx, dw, db);
    return check_launch("navc_layernorm_bwd");
}

extern "C" int navc_embed_ln_bwd(const float* dout, const int64_t* tokens, const int64_t* category,
                                 const float* word_emb, const float* pos_emb, const float* cat_emb, const float* extra,
                                 int group, const float* ln_w, const float* ln_b, float eps, int N, int S, int D,
                                 float* d_word, float* d_pos, float* d_cat, float* d_extra, float* d_ln_w,
                                 float* d_ln_b, void* stream) {
    (void)ln_b;
    NAVC_REQUIRE(dout && tokens && word_emb && pos_emb && ln_w && d_word && d_pos && d_ln_w && d_ln_b,
                 "navc_embed_ln_bwd: null pointer");
    NAVC_REQUIRE(!cat_emb || (category && d_cat), "navc_embed_ln_bwd: category embeddings without ids / gradient");
    NAVC_REQUIRE(!extra || d_extra, "navc_embed_ln_bwd: extra without gradient buffer");
    NAVC_REQUIRE(D % 4 == 0 && D <= 4 * 32 * kLnChunks && group >= 1, "navc_embed_ln_bwd: bad shape");
    const int R = N * S;
    int blocks = (R + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    embed_ln_bwd_kernel<<<blocks, 256, 2 * D * sizeof(float), as_stream(stream)>>>(
        dout, tokens, category, word_emb, pos_emb, cat_emb, extra, group, ln_w, eps, R, S, D, d_word, d_pos,
        cat_emb ? d_cat : nullptr, extra ? d_extra : nullptr, d_ln_w, d_ln_b, nullptr);
    return check_launch("navc_embed_ln_bwd");
}

extern "C" int navc_embed_ln_bwd_packed(const float* dout, const int64_t* tokens, const int64_t* category,
                                        const float* word_emb, const float* pos_emb, const float* cat_emb,
                                        const float* extra, int group, const float* ln_w, float eps, int S, int D,
                                        const int32_t* rowmap, int rows, float* d_word, float* d_pos, float* d_cat,
                                        float* d_extra, float* d_ln_w, float* d_ln_b, void* stream) {
    NAVC_REQUIRE(dout && tokens && word_emb && pos_emb && ln_w && d_word && d_pos && d_ln_w && d_ln_b && rowmap && rows > 0,
                 "navc_embed_ln_bwd_packed: bad arguments");
    NAVC_REQUIRE(!cat_emb || (category && d_cat), "navc_embed_ln_bwd_packed: category embeddings without ids / gradient");
    NAVC_REQUIRE(!extra || d_extra, "navc_embed_ln_bwd_packed: extra without gradient buffer");
    NAVC_REQUIRE(D % 4 == 0 && D <= 4 * 32 * kLnChunks && group >= 1, "navc_embed_ln_bwd_packed: bad shape");
    int blocks = (rows + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    embed_ln_bwd_kernel<<<blocks, 256, 2 * D * sizeof(float), as_stream(stream)>>>(
        dout, tokens, category, word_emb, pos_emb, cat_emb, extra, group, ln_w, eps, rows, S, D, d_word, d_pos,
        cat_emb ? d_cat : nullptr, extra ? d_extra : nullptr, d_ln_w, d_ln_b, rowmap);
    return check_launch("navc_embed_ln_bwd_packed");
}

extern "C" int navc_ce_stats(const float* part_max, const float* part_sum, const int32_t* part_idx, int n_tiles,
                             const float* target_logit, const int64_t* labels, int R, float* lse, float* nll,
                             int32_t* argmax, void* stream) {
    NAVC_REQUIRE(part_max && part_sum && part_idx && target_logit && labels && lse && nll && argmax && R > 0 && n_tiles > 0,
                 "navc_ce_stats: bad arguments");
    ce_stats_kernel<<<(R + 127) / 128, 128, 0, as_stream(stream)>>>(part_max, part_sum, part_idx, n_tiles, target_logit, labels,
                                                                    R, lse, nll, argmax);
    return check_launch("navc_ce_stats");
}

extern "C" int navc_ce_grad(float* logits, const float* lse, const int64_t* labels, const float* scale, int rows, int V,
                            int ld, void* stream) {
    NAVC_REQUIRE(logits && lse && labels && scale && rows > 0 && V > 0 && ld >= V, "navc_ce_grad: bad arguments");
    ce_grad_kernel<<<rows, 256, 0, as_stream(stream)>>>(logits, lse, labels, scale, V, ld);
    return check_launch("navc_ce_grad");
}

extern "C" int navc_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float clip, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, void* stream) {
    NAVC_REQUIRE(p && g && m && v && n > 0 && step >= 1, "navc_clip_adam: bad arguments");
    NAVC_REQUIRE(((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0, "navc_clip_adam: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    clip_adam_kernel<<<ew_blocks((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(
        p, g, m, v, n, clip, (float)((double)lr / bc1), beta1, beta2, eps, weight_decay, (float)(1.0 / sqrt(bc2)));
    return check_launch("navc_clip_adam");
}

extern "C" int navc_self_attention_bwd(const float* qkv, int ld, const int64_t* tokens, int N, int S, int D, int H,
                                       int mask_kind, int watch, const float* d_ctx, float* d_qkv, void* stream) {
    NAVC_REQUIRE(qkv && tokens && d_ctx && d_qkv, "navc_self_attention_bwd: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && H > 0 && D % H == 0 && ld >= 3 * D, "navc_self_attention_bwd: bad shape");
    return dispatch_attn_bwd(D / H, qkv, ld, qkv + D, qkv + 2 * D, ld, tokens, N, S, S, S, H, mask_kind, watch, d_ctx, D,
                             d_qkv, ld, d_qkv + D, d_qkv + 2 * D, ld, as_stream(stream), "navc_self_attention_bwd");
}

extern "C" int navc_cross_attention_bwd(const float* q, int ldq, const float* kv, int ldkv, int N, int S, int E, int D,
                                        int H, int group, const float* d_ctx, float* d_q, int ld_dq, float* d_kv,
                                        int ld_dkv, void* stream) {
    NAVC_REQUIRE(q && kv && d_ctx && d_q && d_kv, "navc_cross_attention_bwd: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && E > 0 && H > 0 && D % H == 0 && group >= 1 && N % group == 0,
                 "navc_cross_attention_bwd: bad shape");
    return dispatch_attn_bwd(D / H, q, ldq, kv, kv + D, ldkv, nullptr, N / group, group * S, S, E, H, 0, 0, d_ctx, D, d_q,
                             ld_dq, d_kv, d_kv + D, ld_dkv, as_stream(stream), "navc_cross_attention_bwd");
}

extern "C" int navc_self_attention_bwd_packed(const float* qkv, int ld, const int64_t* tokens, const int32_t* seq_off, int N,
                                              int S, int D, int H, int mask_kind, int watch, const float* d_ctx,
                                              float* d_qkv, void* stream) {
    NAVC_REQUIRE(qkv && tokens && seq_off && d_ctx && d_qkv, "navc_self_attention_bwd_packed: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && H > 0 && D % H == 0 && ld >= 3 * D, "navc_self_attention_bwd_packed: bad shape");
    return dispatch_attn_bwd(D / H, qkv, ld, qkv + D, qkv + 2 * D, ld, tokens, N, S, S, S, H, mask_kind, watch, d_ctx, D,
                             d_qkv, ld, d_qkv + D, d_qkv + 2 * D, ld, as_stream(stream), "navc_self_attention_bwd_packed",
                             seq_off, 1);
}

extern "C" int navc_cross_attention_bwd_packed(const float* q, int ldq, const float* kv, int ldkv, const int32_t* seq_off,
                                               int N, int S, int E, int D, int H, const float* d_ctx, float* d_q, int ld_dq,
                                               float* d_kv, int ld_dkv, void* stream) {
    NAVC_REQUIRE(q && kv && seq_off && d_ctx && d_q && d_kv, "navc_cross_attention_bwd_packed: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && E > 0 && H > 0 && D % H == 0, "navc_cross_attention_bwd_packed: bad shape");
    // one owner (video) per sequence: the training path has one token row per video
    return dispatch_attn_bwd(D / H, q, ldq, kv, kv + D, ldkv, nullptr, N, S, S, E, H, 0, 0, d_ctx, D, d_q, ld_dq, d_kv,
                             d_kv + D, ld_dkv, as_stream(stream), "navc_cross_attention_bwd_packed", seq_off, 0);
}

// Packed vocabulary projection: rows that were not computed (PAD positions, hidden == 0) all have the same
// log-probabilities const_logp = log_softmax(bias).  Their dlogits only reach the bias gradient.  A loss that
// ignores PAD positions (misc/crit.py:62-84) leaves these rows of g all zero -> one read pass, no atomics.
__global__ void log_softmax_bwd_padrows_kernel(const float* __restrict__ g, int ld, const int32_t* __restrict__ pad_rows,
                                               const float* __restrict__ const_logp, int V, float* __restrict__ db) {
    __shared__ float red[32];
    __shared__ int any;
    const float* gr = g + (size_t)pad_rows[blockIdx.x] * ld;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    float sum = 0.f;
    bool nz = false;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float x = gr[v];
        sum += x;
        nz = nz || (x != 0.f);
    }
    if (nz) any = 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (!any) return;
    sum = warp_sum(lane < nw ? red[lane] : 0.f);
    for (int v = threadIdx.x; v < V; v += blockDim.x) atomicAdd(db + v, gr[v] - expf(const_logp[v]) * sum);
}

extern "C" int navc_log_softmax_bwd_padrows(const float* g, int ld, const int32_t* pad_rows, int n_pad,
                                            const float* const_logp, int V, float* db, void* stream) {
    NAVC_REQUIRE(g && pad_rows && const_logp && db && n_pad > 0 && V > 0 && ld >= V, "navc_log_softmax_bwd_padrows: bad arguments");
    log_softmax_bwd_padrows_kernel<<<n_pad, 256, 0, as_stream(stream)>>>(g, ld, pad_rows, const_logp, V, db);
    return check_launch("navc_log_softmax_bwd_padrows");
}
