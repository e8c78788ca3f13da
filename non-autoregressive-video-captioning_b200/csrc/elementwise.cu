// HBM-bound vectorised kernels: bf16 split, highway gate + BatchNorm + concat, length head,
// embedding gather + LayerNorm, LayerNorm, log-softmax.
#include "common.cuh"

namespace navc {

// ------------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi,
                                  uint16_t* __restrict__ lo, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        if (i + 4 <= n) {
            float4 v = *reinterpret_cast<const float4*>(x + i);
            uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
            split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
            *reinterpret_cast<uint2*>(hi + i) =
                make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
            if (lo)
                *reinterpret_cast<uint2*>(lo + i) =
                    make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
        } else {
            for (int64_t j = i; j < n; ++j) {
                uint16_t h, l;
                split_bf16(x[j], h, l);
                hi[j] = h;
                if (lo) lo[j] = l;
            }
        }
    }
}

__global__ void join_bf16_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                 float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = bf16_to_f32(hi[i]) + (lo ? bf16_to_f32(lo[i]) : 0.f);
}

// ------------------------------------------------------------------------------------------------
// one block per video; thread d owns column d (strided), loops over the F frames.
__global__ void highway_bn_kernel(const float* __restrict__ x, const float* __restrict__ yg, int gate, int F,
                                  int D, int E, int slot, float inv_fm, int accumulate,
                                  const float* __restrict__ rm, const float* __restrict__ rv,
                                  const float* __restrict__ bw, const float* __restrict__ bb, float eps,
                                  float* __restrict__ enc_hidden, float* __restrict__ enc_out,
                                  uint16_t* __restrict__ enc_hi, uint16_t* __restrict__ enc_lo) {
    const int b = blockIdx.x;
    const int ldy = gate ? 2 * D : D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float mean = 0.f, istd = 1.f, w = 1.f, bias = 0.f;
        if (rm) {
            mean = rm[d];
            istd = 1.0f / sqrtf(rv[d] + eps);
            w = bw ? bw[d] : 1.f;
            bias = bb ? bb[d] : 0.f;
        }
        float hsum = 0.f;
        for (int f = 0; f < F; ++f) {
            size_t r = (size_t)b * F + f;
            float xv = x[r * D + d];
            float y = tanhf(yg[r * ldy + d]);
            float o;
            if (gate) {
                float g = 1.0f / (1.0f + expf(-yg[r * ldy + D + d]));
                o = g * xv + (1.0f - g) * y;
            } else {
                o = xv + y;
            }
            hsum += o;
            float v = rm ? ((o - mean) * istd * w + bias) : o;
            size_t oi = ((size_t)b * E + (size_t)slot * F + f) * D + d;
            enc_out[oi] = v;
            if (enc_hi) {
                uint16_t h, l;
                split_bf16(v, h, l);
                enc_hi[oi] = h;
                if (enc_lo) enc_lo[oi] = l;
            }
        }
        if (enc_hidden) {
            float hv = hsum * inv_fm;
            size_t hi_ = (size_t)b * D + d;
            enc_hidden[hi_] = accumulate ? enc_hidden[hi_] + hv : hv;
        }
    }
}

// The same op, vectorised: thread = (4 columns, frame group); a block covers one video with up to 8 frame groups that
// stride over the F frames (the scalar kernel above walks the 60 frames serially per column: 43 us at 128 videos against
// ~11 us of HBM time).  The per-group frame sums meet in shared memory and are added in group order (deterministic).
__global__ void __launch_bounds__(1024) highway_bn_vec_kernel(const float* __restrict__ x, const float* __restrict__ yg, int gate, int F,
                                      int D, int E, int slot, float inv_fm, int accumulate,
                                      const float* __restrict__ rm, const float* __restrict__ rv,
                                      const float* __restrict__ bw, const float* __restrict__ bb, float eps,
                                      float* __restrict__ enc_hidden, float* __restrict__ enc_out,
                                      uint16_t* __restrict__ enc_hi, uint16_t* __restrict__ enc_lo) {
    extern __shared__ float4 hsum_s[];   // [groups][D / 4]
    const int b = blockIdx.x;
    const int nc = D >> 2, groups = blockDim.x / nc;
    const int c = threadIdx.x % nc, fg = threadIdx.x / nc;
    const int ldy = gate ? 2 * D : D;
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), istd = make_float4(1.f, 1.f, 1.f, 1.f), w = istd, bias = mean;
    if (rm) {
        mean = *reinterpret_cast<const float4*>(rm + c * 4);
        const float4 var = *reinterpret_cast<const float4*>(rv + c * 4);
        istd = make_float4(1.0f / sqrtf(var.x + eps), 1.0f / sqrtf(var.y + eps), 1.0f / sqrtf(var.z + eps), 1.0f / sqrtf(var.w + eps));
        if (bw) w = *reinterpret_cast<const float4*>(bw + c * 4);
        if (bb) bias = *reinterpret_cast<const float4*>(bb + c * 4);
    }
    float4 hs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fg < groups)
        for (int f = fg; f < F; f += groups) {
            const size_t r = (size_t)b * F + f;
            const float4 xv = *reinterpret_cast<const float4*>(x + r * D + c * 4);
            const float4 yv = *reinterpret_cast<const float4*>(yg + r * ldy + c * 4);
            float4 o;
            if (gate) {
                const float4 gv = *reinterpret_cast<const float4*>(yg + r * ldy + D + c * 4);
                const float g0 = 1.0f / (1.0f + expf(-gv.x)), g1 = 1.0f / (1.0f + expf(-gv.y));
                const float g2 = 1.0f / (1.0f + expf(-gv.z)), g3 = 1.0f / (1.0f + expf(-gv.w));
                o = make_float4(g0 * xv.x + (1.0f - g0) * tanhf(yv.x), g1 * xv.y + (1.0f - g1) * tanhf(yv.y),
                                g2 * xv.z + (1.0f - g2) * tanhf(yv.z), g3 * xv.w + (1.0f - g3) * tanhf(yv.w));
            } else {
                o = make_float4(xv.x + tanhf(yv.x), xv.y + tanhf(yv.y), xv.z + tanhf(yv.z), xv.w + tanhf(yv.w));
            }
            hs.x += o.x; hs.y += o.y; hs.z += o.z; hs.w += o.w;
            float4 v = o;
            if (rm) v = make_float4((o.x - mean.x) * istd.x * w.x + bias.x, (o.y - mean.y) * istd.y * w.y + bias.y,
                                    (o.z - mean.z) * istd.z * w.z + bias.z, (o.w - mean.w) * istd.w * w.w + bias.w);
            const size_t oi = ((size_t)b * E + (size_t)slot * F + f) * D + c * 4;
            *reinterpret_cast<float4*>(enc_out + oi) = v;
            if (enc_hi) {
                uint2 h, l;
                split_bf16x4(v, h, l);
                *reinterpret_cast<uint2*>(enc_hi + oi) = h;
                if (enc_lo) *reinterpret_cast<uint2*>(enc_lo + oi) = l;
            }
        }
    if (!enc_hidden) return;
    if (fg < groups) hsum_s[fg * nc + c] = hs;
    __syncthreads();
    if (fg == 0) {
        float4 t = hsum_s[c];
        for (int gq = 1; gq < groups; ++gq) {
            const float4 u = hsum_s[gq * nc + c];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        float4* dst = reinterpret_cast<float4*>(enc_hidden + (size_t)b * D + c * 4);
        float4 hv = make_float4(t.x * inv_fm, t.y * inv_fm, t.z * inv_fm, t.w * inv_fm);
        if (accumulate) { const float4 old = *dst; hv.x += old.x; hv.y += old.y; hv.z += old.z; hv.w += old.w; }
        *dst = hv;
    }
}

// norm_type='ln' variant (models/joint_representation.py:20, 46-47): LayerNorm over the D features of every frame
// row instead of the BatchNorm affine.  One block per video, one warp per frame row (strided); enc_hidden (the
// frame mean of the UN-normalised highway output) is accumulated through shared memory.  D <= 1024.
__global__ void highway_ln_kernel(const float* __restrict__ x, const float* __restrict__ yg, int gate, int F,
                                  int D, int E, int slot, float inv_fm, int accumulate,
                                  const float* __restrict__ lw, const float* __restrict__ lb, float eps,
                                  float* __restrict__ enc_hidden, float* __restrict__ enc_out,
                                  uint16_t* __restrict__ enc_hi, uint16_t* __restrict__ enc_lo) {
    extern __shared__ float hsum[];  // [D]
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int ldy = gate ? 2 * D : D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) hsum[d] = 0.f;
    __syncthreads();
    for (int f = warp; f < F; f += nw) {
        const size_t r = (size_t)b * F + f;
        float o[32];  // D / 32 values per lane
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int d = lane + i * 32;
            o[i] = 0.f;
            if (d < D) {
                const float xv = x[r * D + d];
                const float y = tanhf(yg[r * ldy + d]);
                if (gate) {
                    const float g = 1.0f / (1.0f + expf(-yg[r * ldy + D + d]));
                    o[i] = g * xv + (1.0f - g) * y;
                } else {
                    o[i] = xv + y;
                }
                s += o[i];
                atomicAdd(&hsum[d], o[i]);
            }
        }
        const float mean = warp_sum(s) / (float)D;
        float v2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int d = lane + i * 32;
            if (d < D) v2 += (o[i] - mean) * (o[i] - mean);
        }
        const float istd = 1.0f / sqrtf(warp_sum(v2) / (float)D + eps);   // biased variance, as nn.LayerNorm
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int d = lane + i * 32;
            if (d < D) {
                const float v = (o[i] - mean) * istd * lw[d] + lb[d];
                const size_t oi = ((size_t)b * E + (size_t)slot * F + f) * D + d;
                enc_out[oi] = v;
                if (enc_hi) {
                    uint16_t h, l;
                    split_bf16(v, h, l);
                    enc_hi[oi] = h;
                    if (enc_lo) enc_lo[oi] = l;
                }
            }
        }
    }
    __syncthreads();
    if (enc_hidden)
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const float hv = hsum[d] * inv_fm;
            const size_t hi_ = (size_t)b * D + d;
            enc_hidden[hi_] = accumulate ? enc_hidden[hi_] + hv : hv;
        }
}

// ------------------------------------------------------------------------------------------------
// one block (256 threads) per video.  dynamic smem: mean[D] + h1[D] + logits[max_len]
__global__ void length_head_kernel(const float* __restrict__ enc_out, int E, int D,
                                   const float* __restrict__ w1, const float* __restrict__ b1,
                                   const float* __restrict__ w2, const float* __restrict__ b2, int max_len,
                                   float* __restrict__ enc_mean, float* __restrict__ pred_length) {
    extern __shared__ float sm[];
    float* mean = sm;
    float* h1 = sm + D;
    float* logit = sm + 2 * D;
    const int b = blockIdx.x;
    const float* src = enc_out + (size_t)b * E * D;
    // frame mean: thread = (column, frame group); the groups' partial sums meet in h1[] and are added in group order
    // (E serial loads per column left 3/4 of a 1024-thread block idle at D = 512)
    {
        const int groups = (int)blockDim.x / D;
        if (groups >= 2) {
            // partials live in dynamic shared memory behind the head's buffers: [groups][D]
            float* part = sm + 2 * D + (max_len > 0 ? max_len : 0);
            const int d = threadIdx.x % D, gq = threadIdx.x / D;
            if (gq < groups) {
                float s = 0.f;
                for (int e = gq; e < E; e += groups) s += src[(size_t)e * D + d];
                part[gq * D + d] = s;
            }
            __syncthreads();
            if (gq == 0) {
                float s = part[d];
                for (int q = 1; q < groups; ++q) s += part[q * D + d];
                const float m = s / (float)E;
                mean[d] = m;
                if (enc_mean) enc_mean[(size_t)b * D + d] = m;
            }
        } else {
            for (int d = threadIdx.x; d < D; d += blockDim.x) {
                float s = 0.f;
                for (int e = 0; e < E; ++e) s += src[(size_t)e * D + d];
                float m = s / (float)E;
                mean[d] = m;
                if (enc_mean) enc_mean[(size_t)b * D + d] = m;
            }
        }
    }
    if (!w1) return;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < D; j += nw) {
        const float* wr = w1 + (size_t)j * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(wr[d], mean[d], s);
        s = warp_sum(s);
        if (lane == 0) h1[j] = fmaxf(s + b1[j], 0.f);
    }
    __syncthreads();
    for (int j = warp; j < max_len; j += nw) {
        const float* wr = w2 + (size_t)j * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(wr[d], h1[d], s);
        s = warp_sum(s);
        if (lane == 0) logit[j] = s + b2[j];
    }
    __syncthreads();
    if (warp == 0) {
        float m = -INFINITY;
        for (int j = lane; j < max_len; j += 32) m = fmaxf(m, logit[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < max_len; j += 32) s += expf(logit[j] - m);
        s = warp_sum(s);
        float lse = m + logf(s);
        for (int j = lane; j < max_len; j += 32) pred_length[(size_t)b * max_len + j] = logit[j] - lse;
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm of one row held in registers by one warp: lane owns float4 chunks c = lane + 32*t.
constexpr int kMaxChunks = 8;  // D <= 1024

__device__ __forceinline__ void warp_ln_store(float4 (&v)[kMaxChunks], int nchunk, int D, int lane,
                                              const float* __restrict__ w, const float* __restrict__ b,
                                              float eps, bool zero, size_t row, float* out_f32,
                                              uint16_t* out_hi, uint16_t* out_lo) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxChunks; ++t)
        if (t < nchunk && lane + 32 * t < D / 4) s += v[t].x + v[t].y + v[t].z + v[t].w;
    float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxChunks; ++t)
        if (t < nchunk && lane + 32 * t < D / 4) {
            float a = v[t].x - mean, bq = v[t].y - mean, c = v[t].z - mean, d = v[t].w - mean;
            q += a * a + bq * bq + c * c + d * d;
        }
    float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int t = 0; t < kMaxChunks; ++t) {
        int c = lane + 32 * t;
        if (t < nchunk && c < D / 4) {
            float4 ww = *reinterpret_cast<const float4*>(w + c * 4);
            float4 bv = *reinterpret_cast<const float4*>(b + c * 4);
            float4 o;
            o.x = (v[t].x - mean) * rstd * ww.x + bv.x;
            o.y = (v[t].y - mean) * rstd * ww.y + bv.y;
            o.z = (v[t].z - mean) * rstd * ww.z + bv.z;
            o.w = (v[t].w - mean) * rstd * ww.w + bv.w;
            if (zero) o = make_float4(0.f, 0.f, 0.f, 0.f);
            size_t oi = row * D + (size_t)c * 4;
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi) = o;
            if (out_hi) {
                uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_bf16(o.x, h0, l0); split_bf16(o.y, h1, l1); split_bf16(o.z, h2, l2); split_bf16(o.w, h3, l3);
                *reinterpret_cast<uint2*>(out_hi + oi) =
                    make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
                if (out_lo)
                    *reinterpret_cast<uint2*>(out_lo + oi) =
                        make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
            }
        }
    }
}

// single block: exclusive prefix sum of lens -> seq_off[0..N]; rowmap[seq_off[n] + s] = n*S + s
__global__ void pack_rows_kernel(const int32_t* __restrict__ lens, int N, int S, int32_t* __restrict__ seq_off,
                                 int32_t* __restrict__ rowmap) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += blockDim.x) {
        const int n = base + threadIdx.x;
        const int len = n < N ? min(max(lens[n], 0), S) : 0;
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        const int start = carry + woff + inc - len;
        if (n < N) {
            seq_off[n] = start;
            for (int j = 0; j < len; ++j) rowmap[start + j] = n * S + j;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = start + len;
        __syncthreads();
        (void)nw;
    }
    if (threadIdx.x == 0) seq_off[N] = carry;
}

__global__ void embed_ln_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ category,
                                const float* __restrict__ word, const float* __restrict__ pos,
                                const float* __restrict__ cat, const float* __restrict__ extra, int group,
                                const float* __restrict__ lw, const float* __restrict__ lb, float eps, int R,
                                int S, int D, float* out_f32, uint16_t* out_hi, uint16_t* out_lo,
                                const int32_t* __restrict__ rowmap, const int32_t* __restrict__ count,
                                int64_t* __restrict__ tok_out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    int src = row;
    if (rowmap) {  // packed rows: row -> padded position
        if (row >= __ldg(count)) return;
        src = rowmap[row];
    }
    const int n = src / S, s = src % S;
    const int64_t tok = tokens[src];
    if (tok_out && lane == 0) tok_out[row] = tok;
    const float* wr = word + (size_t)tok * D;
    const float* pr = pos + (size_t)s * D;
    const float* cr = cat ? cat + (size_t)category[n / group] * D : nullptr;
    const float* er = extra ? extra + (size_t)(n / group) * D : nullptr;
    const int nchunk = (D / 4 + 31) / 32;
    float4 v[kMaxChunks];
#pragma unroll
    for (int t = 0; t < kMaxChunks; ++t) {
        int c = lane + 32 * t;
        if (t < nchunk && c < D / 4) {
            float4 a = *reinterpret_cast<const float4*>(wr + c * 4);
            float4 p = *reinterpret_cast<const float4*>(pr + c * 4);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
            if (cr) {
                float4 q = *reinterpret_cast<const float4*>(cr + c * 4);
                a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
            }
            if (er) {
                float4 q = *reinterpret_cast<const float4*>(er + c * 4);
                a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
            }
            v[t] = a;
        }
    }
    warp_ln_store(v, nchunk, D, lane, lw, lb, eps, false, (size_t)row, out_f32, out_hi, out_lo);
}

__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                                 const float* __restrict__ lb, float eps, const int64_t* __restrict__ row_tokens,
                                 int R, int D, float* out_f32, uint16_t* out_hi, uint16_t* out_lo) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const int nchunk = (D / 4 + 31) / 32;
    float4 v[kMaxChunks];
#pragma unroll
    for (int t = 0; t < kMaxChunks; ++t) {
        int c = lane + 32 * t;
        if (t < nchunk && c < D / 4) v[t] = *reinterpret_cast<const float4*>(x + (size_t)row * D + c * 4);
    }
    bool zero = row_tokens ? (row_tokens[row] == NAVC_PAD) : false;
    warp_ln_store(v, nchunk, D, lane, lw, lb, eps, zero, (size_t)row, out_f32, out_hi, out_lo);
}

// ------------------------------------------------------------------------------------------------
// one block per row: log_softmax over V columns
__global__ void log_softmax_kernel(const float* __restrict__ x, float* __restrict__ out, int V, int ld, int ld_out,
                                   const int32_t* __restrict__ rowmap) {
    __shared__ float red[32];
    const float* r = x + (size_t)blockIdx.x * ld;
    float* o = out + (rowmap ? (size_t)rowmap[blockIdx.x] : (size_t)blockIdx.x) * ld_out;   // packed rows -> padded output
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, r[i]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = (lane < nw) ? red[lane] : -INFINITY;
    m = warp_max(m);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(r[i] - m);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = (lane < nw) ? red[lane] : 0.f;
    s = warp_sum(s);
    const float lse = m + logf(s);
    for (int i = threadIdx.x; i < V; i += blockDim.x) o[i] = r[i] - lse;
}

}  // namespace navc

using namespace navc;

extern "C" int navc_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, int64_t n, void* stream) {
    NAVC_REQUIRE(x && hi && n >= 0, "navc_split_bf16: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    split_bf16_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, hi, lo, n);
    return check_launch("navc_split_bf16");
}

// Refresh of the packed weights after an optimizer step (same parameter storage, new values): ONE launch copies every
// source tensor into its rows of the concatenated fp32 operand (where that is a copy) and rewrites the bf16 hi / lo copies.
// items: [n][5] int64 = {src fp32, dst fp32 or 0, hi or 0, lo or 0, element count}; blockIdx.y = item.
__global__ void refresh_pack_kernel(const long long* __restrict__ items) {
    const long long* it = items + (size_t)blockIdx.y * 5;
    const float* __restrict__ src = reinterpret_cast<const float*>(it[0]);
    float* __restrict__ dst = reinterpret_cast<float*>(it[1]);
    uint16_t* __restrict__ hi = reinterpret_cast<uint16_t*>(it[2]);
    uint16_t* __restrict__ lo = reinterpret_cast<uint16_t*>(it[3]);
    const int64_t n = it[4];
    const bool vec = ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0 && ((((uintptr_t)hi) | ((uintptr_t)lo)) & 7) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (vec && i + 4 <= n) {
            const float4 v = *reinterpret_cast<const float4*>(src + i);
            if (dst) *reinterpret_cast<float4*>(dst + i) = v;
            if (hi) {
                uint2 h, l;
                split_bf16x4(v, h, l);
                *reinterpret_cast<uint2*>(hi + i) = h;
                if (lo) *reinterpret_cast<uint2*>(lo + i) = l;
            }
        } else {
            for (int64_t j = i; j < n && j < i + 4; ++j) {
                const float x = src[j];
                if (dst) dst[j] = x;
                if (hi) {
                    uint16_t h, l;
                    split_bf16(x, h, l);
                    hi[j] = h;
                    if (lo) lo[j] = l;
                }
            }
        }
    }
}

extern "C" int navc_refresh_pack(const int64_t* items, int n_items, void* stream) {
    NAVC_REQUIRE(items && n_items > 0 && n_items <= 65535, "navc_refresh_pack: bad arguments");
    refresh_pack_kernel<<<dim3(64, n_items), 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(items));
    return check_launch("navc_refresh_pack");
}

extern "C" int navc_join_bf16(const uint16_t* hi, const uint16_t* lo, float* out, int64_t n, void* stream) {
    NAVC_REQUIRE(hi && out && n >= 0, "navc_join_bf16: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    join_bf16_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(hi, lo, out, n);
    return check_launch("navc_join_bf16");
}

extern "C" int navc_highway_bn(const float* x, const float* yg, int gate, int B, int F, int D, int E, int slot,
                               int n_modalities, int accumulate, const float* bn_rm, const float* bn_rv,
                               const float* bn_w, const float* bn_b, float bn_eps, float* enc_hidden,
                               float* enc_out, uint16_t* enc_hi, uint16_t* enc_lo, void* stream) {
    NAVC_REQUIRE(x && yg && enc_out, "navc_highway_bn: null pointer");
    NAVC_REQUIRE(B > 0 && F > 0 && D > 0 && (slot + 1) * F <= E, "navc_highway_bn: bad shape");
    NAVC_REQUIRE((bn_rm == nullptr) == (bn_rv == nullptr), "navc_highway_bn: need both running stats");
    const bool vec = D % 4 == 0 && D / 4 <= 1024 && ((((uintptr_t)x) | ((uintptr_t)yg) | ((uintptr_t)enc_out) | ((uintptr_t)enc_hidden) |
                                                      ((uintptr_t)bn_rm) | ((uintptr_t)bn_rv) | ((uintptr_t)bn_w) | ((uintptr_t)bn_b)) & 15) == 0 &&
                     ((((uintptr_t)enc_hi) | ((uintptr_t)enc_lo)) & 7) == 0;
    if (vec) {
        const int nc = D / 4;
        int groups = 1024 / nc;
        if (groups > 8) groups = 8;
        if (groups > F) groups = F;
        highway_bn_vec_kernel<<<B, nc * groups, (size_t)groups * nc * sizeof(float4), as_stream(stream)>>>(
            x, yg, gate, F, D, E, slot, 1.0f / ((float)F * (float)n_modalities), accumulate, bn_rm, bn_rv, bn_w, bn_b, bn_eps,
            enc_hidden, enc_out, enc_hi, enc_lo);
        return check_launch("navc_highway_bn");
    }
    int threads = D >= 512 ? 512 : ((D + 31) / 32) * 32;
    highway_bn_kernel<<<B, threads, 0, as_stream(stream)>>>(x, yg, gate, F, D, E, slot,
                                                            1.0f / ((float)F * (float)n_modalities), accumulate,
                                                            bn_rm, bn_rv, bn_w, bn_b, bn_eps, enc_hidden, enc_out,
                                                            enc_hi, enc_lo);
    return check_launch("navc_highway_bn");
}

extern "C" int navc_highway_ln(const float* x, const float* yg, int gate, int B, int F, int D, int E, int slot,
                               int n_modalities, int accumulate, const float* ln_w, const float* ln_b, float ln_eps,
                               float* enc_hidden, float* enc_out, uint16_t* enc_hi, uint16_t* enc_lo, void* stream) {
    NAVC_REQUIRE(x && yg && enc_out && ln_w && ln_b, "navc_highway_ln: null pointer");
    NAVC_REQUIRE(B > 0 && F > 0 && D > 0 && D <= 1024 && (slot + 1) * F <= E, "navc_highway_ln: bad shape (D <= 1024)");
    highway_ln_kernel<<<B, 256, (size_t)D * sizeof(float), as_stream(stream)>>>(
        x, yg, gate, F, D, E, slot, 1.0f / ((float)F * (float)n_modalities), accumulate, ln_w, ln_b, ln_eps, enc_hidden,
        enc_out, enc_hi, enc_lo);
    return check_launch("navc_highway_ln");
}

extern "C" int navc_length_head(const float* enc_out, int B, int E, int D, const float* w1, const float* b1,
                                const float* w2, const float* b2, int max_len, float* enc_mean,
                                float* pred_length, void* stream) {
    NAVC_REQUIRE(enc_out && B > 0 && E > 0 && D > 0, "navc_length_head: bad arguments");
    NAVC_REQUIRE(!w1 || (b1 && w2 && b2 && pred_length && max_len > 0), "navc_length_head: missing head weights");
    const int threads = 1024;
    const int groups = threads / D;   // frame groups of the mean (>= 2: partial sums in shared memory)
    size_t smem = (size_t)(2 * D + (max_len > 0 ? max_len : 0) + (groups >= 2 ? groups * D : 0)) * sizeof(float);
    length_head_kernel<<<B, threads, smem, as_stream(stream)>>>(enc_out, E, D, w1, b1, w2, b2, max_len, enc_mean,
                                                           pred_length);
    return check_launch("navc_length_head");
}

extern "C" int navc_embed_ln(const int64_t* tokens, const int64_t* category, const float* word_emb,
                             const float* pos_emb, const float* cat_emb, const float* extra, int group,
                             const float* ln_w, const float* ln_b, float eps, int N, int S, int D, float* out_f32,
                             uint16_t* out_hi, uint16_t* out_lo, void* stream) {
    NAVC_REQUIRE(tokens && word_emb && pos_emb && ln_w && ln_b, "navc_embed_ln: null pointer");
    NAVC_REQUIRE(!cat_emb || category, "navc_embed_ln: category embeddings without category ids");
    NAVC_REQUIRE(D % 4 == 0 && D <= 4 * 32 * kMaxChunks, "navc_embed_ln: D must be a multiple of 4 and <= 1024");
    NAVC_REQUIRE(group >= 1, "navc_embed_ln: group must be >= 1");
    int R = N * S;
    int wpb = 8;
    embed_ln_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(
        tokens, category, word_emb, pos_emb, cat_emb, extra, group, ln_w, ln_b, eps, R, S, D, out_f32, out_hi,
        out_lo, nullptr, nullptr, nullptr);
    return check_launch("navc_embed_ln");
}

// one warp per output row: 16-byte copies of the bf16 hi / lo rows
__global__ void gather_rows_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo, int D,
                                   const int32_t* __restrict__ rows, const int32_t* __restrict__ count, int max_rows,
                                   uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= max_rows || k >= __ldg(count)) return;
    const size_t src = (size_t)rows[k] * D, dst = (size_t)k * D;
    for (int c = lane * 8; c < D; c += 256) {
        *reinterpret_cast<uint4*>(out_hi + dst + c) = *reinterpret_cast<const uint4*>(in_hi + src + c);
        if (in_lo) *reinterpret_cast<uint4*>(out_lo + dst + c) = *reinterpret_cast<const uint4*>(in_lo + src + c);
    }
}

extern "C" int navc_gather_rows(const uint16_t* in_hi, const uint16_t* in_lo, int D, const int32_t* rows,
                                const int32_t* count, int max_rows, uint16_t* out_hi, uint16_t* out_lo, void* stream) {
    NAVC_REQUIRE(in_hi && rows && count && out_hi && (!in_lo || out_lo) && D % 8 == 0 && max_rows > 0,
                 "navc_gather_rows: bad arguments");
    gather_rows_kernel<<<(max_rows + 7) / 8, 256, 0, as_stream(stream)>>>(in_hi, in_lo, D, rows, count, max_rows, out_hi, out_lo);
    return check_launch("navc_gather_rows");
}

// two row gathers through the same list in one launch (the last layer's context rows and residual rows)
__global__ void gather_rows2_kernel(const uint16_t* __restrict__ a_hi, const uint16_t* __restrict__ a_lo, uint16_t* __restrict__ oa_hi,
                                    uint16_t* __restrict__ oa_lo, const uint16_t* __restrict__ b_hi, const uint16_t* __restrict__ b_lo,
                                    uint16_t* __restrict__ ob_hi, uint16_t* __restrict__ ob_lo, int D, const int32_t* __restrict__ rows,
                                    const int32_t* __restrict__ count, int max_rows) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= max_rows || k >= __ldg(count)) return;
    const size_t src = (size_t)rows[k] * D, dst = (size_t)k * D;
    for (int c = lane * 8; c < D; c += 256) {
        const uint4 v0 = *reinterpret_cast<const uint4*>(a_hi + src + c);
        const uint4 v1 = *reinterpret_cast<const uint4*>(b_hi + src + c);
        uint4 v2 = make_uint4(0u, 0u, 0u, 0u), v3 = v2;
        if (a_lo) v2 = *reinterpret_cast<const uint4*>(a_lo + src + c);
        if (b_lo) v3 = *reinterpret_cast<const uint4*>(b_lo + src + c);
        *reinterpret_cast<uint4*>(oa_hi + dst + c) = v0;
        *reinterpret_cast<uint4*>(ob_hi + dst + c) = v1;
        if (a_lo) *reinterpret_cast<uint4*>(oa_lo + dst + c) = v2;
        if (b_lo) *reinterpret_cast<uint4*>(ob_lo + dst + c) = v3;
    }
}

extern "C" int navc_gather_rows2(const uint16_t* a_hi, const uint16_t* a_lo, uint16_t* oa_hi, uint16_t* oa_lo, const uint16_t* b_hi,
                                 const uint16_t* b_lo, uint16_t* ob_hi, uint16_t* ob_lo, int D, const int32_t* rows,
                                 const int32_t* count, int max_rows, void* stream) {
    NAVC_REQUIRE(a_hi && b_hi && oa_hi && ob_hi && rows && count && (!a_lo || oa_lo) && (!b_lo || ob_lo) && D % 8 == 0 && max_rows > 0,
                 "navc_gather_rows2: bad arguments");
    gather_rows2_kernel<<<(max_rows + 7) / 8, 256, 0, as_stream(stream)>>>(a_hi, a_lo, oa_hi, oa_lo, b_hi, b_lo, ob_hi, ob_lo, D, rows, count,
                                                                        max_rows);
    return check_launch("navc_gather_rows2");
}

extern "C" int navc_pack_rows(const int32_t* lens, int N, int S, int32_t* seq_off, int32_t* rowmap, void* stream) {
    NAVC_REQUIRE(lens && seq_off && rowmap && N > 0 && S > 0, "navc_pack_rows: bad arguments");
    pack_rows_kernel<<<1, 1024, 0, as_stream(stream)>>>(lens, N, S, seq_off, rowmap);
    return check_launch("navc_pack_rows");
}

extern "C" int navc_embed_ln_packed(const int64_t* tokens, const int64_t* category, const float* word_emb,
                                    const float* pos_emb, const float* cat_emb, const float* extra, int group,
                                    const float* ln_w, const float* ln_b, float eps, int N, int S, int D,
                                    const int32_t* seq_off, const int32_t* rowmap, int64_t* tok_out, float* out_f32,
                                    uint16_t* out_hi, uint16_t* out_lo, void* stream) {
    NAVC_REQUIRE(tokens && word_emb && pos_emb && ln_w && ln_b && seq_off && rowmap, "navc_embed_ln_packed: null pointer");
    NAVC_REQUIRE(!cat_emb || category, "navc_embed_ln_packed: category embeddings without category ids");
    NAVC_REQUIRE(D % 4 == 0 && D <= 4 * 32 * kMaxChunks && group >= 1, "navc_embed_ln_packed: bad shape");
    const int R = N * S, wpb = 8;
    embed_ln_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(
        tokens, category, word_emb, pos_emb, cat_emb, extra, group, ln_w, ln_b, eps, R, S, D, out_f32, out_hi,
        out_lo, rowmap, seq_off + N, tok_out);
    return check_launch("navc_embed_ln_packed");
}

extern "C" int navc_layernorm(const float* x, const float* w, const float* b, float eps, const int64_t* row_tokens,
                              int M, int D, float* out_f32, uint16_t* out_hi, uint16_t* out_lo, void* stream) {
    NAVC_REQUIRE(x && w && b, "navc_layernorm: null pointer");
    NAVC_REQUIRE(D % 4 == 0 && D <= 4 * 32 * kMaxChunks, "navc_layernorm: D must be a multiple of 4 and <= 1024");
    int wpb = 8;
    layernorm_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(x, w, b, eps, row_tokens, M, D,
                                                                             out_f32, out_hi, out_lo);
    return check_launch("navc_layernorm");
}

extern "C" int navc_log_softmax(const float* logits, float* out, int M, int V, int ld, void* stream) {
    NAVC_REQUIRE(logits && out && M > 0 && V > 0 && ld >= V, "navc_log_softmax: bad arguments");
    log_softmax_kernel<<<M, 256, 0, as_stream(stream)>>>(logits, out, V, ld, ld, nullptr);
    return check_launch("navc_log_softmax");
}

extern "C" int navc_log_softmax_ld(const float* logits, int ld_in, float* out, int ld_out, int M, int V, void* stream) {
    NAVC_REQUIRE(logits && out && M > 0 && V > 0 && ld_in >= V && ld_out >= V, "navc_log_softmax_ld: bad arguments");
    log_softmax_kernel<<<M, 256, 0, as_stream(stream)>>>(logits, out, V, ld_in, ld_out, nullptr);
    return check_launch("navc_log_softmax_ld");
}

extern "C" int navc_log_softmax_rows(const float* logits, int ld_in, float* out, int ld_out, const int32_t* rowmap, int rows,
                                     int V, void* stream) {
    NAVC_REQUIRE(logits && out && rowmap && rows > 0 && V > 0 && ld_in >= V && ld_out >= V, "navc_log_softmax_rows: bad arguments");
    log_softmax_kernel<<<rows, 256, 0, as_stream(stream)>>>(logits, out, V, ld_in, ld_out, rowmap);
    return check_launch("navc_log_softmax_rows");
}
