// tcgen05 attention backward over packed rows (training, dk = 64): the autograd of models/bert.py:154-176
// (scores = Q K^T / sqrt(dk), masked_fill(-1e7), softmax, context = P V) for one (owner, head) per CTA.
//
// The training shapes are the opposite of the decode loop's: an owner has FEW query rows (one caption per video:
// len <= 30 rows) and, for text -> video attention, E = 120 keys; 8 heads x 1024 videos = 8192 independent problems per
// launch.  The round-1 kernel (backward.cu attn_bwd_kernel) contracts them on the CUDA cores with 4 x 4 register tiles
// (182 us per cross-attention launch at 256 videos, 17 % of the NACF step).  Here the five products run as UMMAs:
//
//   S  = Q K^T          A = Q  [q rows, 64]  K-major      B = K  [keys, 64]  K-major      -> TMEM columns   0..127
//   dP = dO V^T         A = dO [q rows, 64]  K-major      B = V  [keys, 64]  K-major      -> TMEM columns 128..255
//   (warp 0, one query row per lane: softmax statistics, P, dS = P o (dP - sum_j P dP) / sqrt(dk) -> shared memory)
//   dV = P^T dO         A = P  [q rows, keys] MN-major    B = dO [q rows, 64] MN-major    -> TMEM columns   0..63
//   dK = dS^T Q         A = dS [q rows, keys] MN-major    B = Q  [q rows, 64] MN-major    -> TMEM columns  64..127
//   dQ = dS K           A = dS [q rows, keys] K-major     B = K  [keys, 64]   MN-major    -> TMEM columns 128..191
//
// The M = 128 rows of S / dP / dQ are the (<= 32) query rows plus don't-care rows whose results are never read (their
// A rows are whatever follows the 32-row operand in shared memory); the M = 128 rows of dV / dK are the keys.  All
// operands are written into the 128-byte-swizzled shared-memory image a TMA box load would produce, straight from the
// fp32 activations the training path keeps (split into bf16 hi / lo on the way), so the kernel needs no operand copies
// from the forward pass.  P and dS reuse V's buffer (V is dead once dP has been issued), dK / dV are staged through the
// K / V buffers for coalesced fp32 stores.  80 KB of operands + 256 TMEM columns per CTA: two CTAs per SM.
// Split mode issues lo*hi + hi*lo + hi*hi for every product.  Masks as backward.cu (causal (+ watch) / diagonal, -1e7 fill
// after the scale; filled scores get no gradient).
#include <stdlib.h>

#include "tc_common.cuh"

namespace navc {

constexpr float kMaskFillT = -10e6f;   // models/bert.py:161
constexpr int AB_THREADS = 128;
// shared-memory image (bytes; every operand: hi part, then lo part)
constexpr int AB_Q = 0;          // [32 rows x 64]  4 KB + 4 KB
constexpr int AB_DO = 8192;      // [32 rows x 64]
constexpr int AB_K = 16384;      // [128 rows x 64] 16 KB + 16 KB   (later: dK staging, 128 x 64 fp32)
constexpr int AB_V = 49152;      // [128 rows x 64]                 (later: P / dS, then dV staging)
constexpr int AB_P = AB_V;       // P:  2 key blocks x [32 rows x 64] hi (8 KB), lo (8 KB)
constexpr int AB_DS = AB_V + 16384;
constexpr int AB_TAIL = 98304;   // the M = 128 reads of the 32-row A operands end below this offset
constexpr int AB_SMEM = AB_TAIL + 256 + 1024;   // + barriers, TMEM slot, 32 row dots

struct AbParams {
    const float* q; int ldq;
    const float* k; const float* v; int ldkv;
    const float* d_ctx; int ld_dctx;
    const float* ctx; int ld_ctx;     // forward context (or NULL): sum_j P_ij dP_ij = dO_i . O_i
    float* dq; int ld_dq;
    float* dk; float* dv; int ld_dkv;
    // optional split output of dK / dV (what the K|V projection's gradient GEMMs consume): bf16 hi (/ lo) rows with leading
    // dimension ld_dkv and the column sums (bias gradient) accumulated with atomicAdd; then dk / dv (fp32) are not written
    uint16_t* dk_hi; uint16_t* dk_lo; uint16_t* dv_hi; uint16_t* dv_lo;
    float* dk_colsum; float* dv_colsum;
    const int32_t* seq_off;
    int is_self, E, S, mask_kind, watch;
};

// optional in-kernel timeline (navc_debug_trace_attn_bwd, tools/attn_bwd_trace.py): thread 0 of every CTA stamps clock64()
__device__ unsigned long long* ab_trace_buf = nullptr;
struct AbTrace {
    unsigned long long* p;
    int n;
    __device__ AbTrace(int cta) : p(nullptr), n(0) {
        if (ab_trace_buf && threadIdx.x == 0) p = ab_trace_buf + (size_t)cta * 16;
    }
    __device__ __forceinline__ void ev() {
        if (p && n < 16) p[n++] = (unsigned long long)clock64();
    }
};

// MN-major operand whose 64-wide MN blocks are `lbo` bytes apart (tc_common.cuh make_smem_desc_mn: 8192)
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ROWS x 64 fp32 (global, row stride ld) -> bf16 hi / lo tiles in the swizzled image; rows >= n_valid become zero.
// All loads of a tile are issued before the first conversion (a load per loop iteration, each waiting for the shared-memory
// store of the previous one, cost 40 serial DRAM round trips per CTA: 26 us per item).
template <int ROWS>
__device__ __forceinline__ void ab_fetch(const float* __restrict__ src, int ld, int n_valid, int tid, float4 (&v)[ROWS / 8]) {
#pragma unroll
    for (int it = 0; it < ROWS / 8; ++it) {
        const int idx = it * AB_THREADS + tid;
        const int row = idx >> 4, c4 = idx & 15;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n_valid) v[it] = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * ld + c4 * 4));
    }
}
template <bool kX3, int ROWS>
__device__ __forceinline__ void ab_put(const float4 (&v)[ROWS / 8], uint8_t* hi, uint8_t* lo, int tid) {
#pragma unroll
    for (int it = 0; it < ROWS / 8; ++it) {
        const int idx = it * AB_THREADS + tid;
        const int row = idx >> 4, c4 = idx & 15;
        uint2 h, l;
        split_bf16x4(v[it], h, l);
        const int off = row * 128 + ((((c4 >> 1) ^ (row & 7))) << 4) + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(hi + off) = h;
        if (kX3) *reinterpret_cast<uint2*>(lo + off) = l;
    }
}

template <bool kX3>
__device__ __forceinline__ void ab_mma3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, bool first) {
    if (kX3) {
        tc_mma_bf16(d, a_lo, b_hi, idesc, first ? 0u : 1u);
        tc_mma_bf16(d, a_hi, b_lo, idesc, 1u);
        tc_mma_bf16(d, a_hi, b_hi, idesc, 1u);
    } else {
        tc_mma_bf16(d, a_hi, b_hi, idesc, first ? 0u : 1u);
    }
}

template <bool kX3>
__global__ void __launch_bounds__(AB_THREADS, 2) attn_bwd_tc_kernel(AbParams p) {
    extern __shared__ uint8_t ab_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(sm);
    const uint32_t bar1 = sb + AB_TAIL, bar2 = sb + AB_TAIL + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + AB_TAIL + 16);
    float* delta_s = reinterpret_cast<float*>(sm + AB_TAIL + 64);
    const int g = blockIdx.x, h = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    AbTrace tr(blockIdx.y * gridDim.x + blockIdx.x);
    tr.ev();   // 0 start
    const int r0 = __ldg(p.seq_off + g);
    int nq = __ldg(p.seq_off + g + 1) - r0;
    if (nq > 32) nq = 32;
    const size_t kv_row0 = p.is_self ? (size_t)r0 : (size_t)g * p.E;
    const int nk = p.is_self ? nq : p.E;
    float* dkb = p.dk + kv_row0 * p.ld_dkv + h * 64;
    float* dvb = p.dv + kv_row0 * p.ld_dkv + h * 64;
    const size_t hrow0 = kv_row0 * p.ld_dkv + h * 64;   // split output: element offset of this head's first row
    if (nq <= 0) {   // no query rows: the owner's keys get zero gradient (text -> video attention only)
        if (!p.is_self)
            for (int idx = tid; idx < nk * 16; idx += AB_THREADS) {
                const int row = idx >> 4, c = idx & 15;
                if (p.dk_hi) {
                    const size_t o = hrow0 + (size_t)row * p.ld_dkv + c * 4;
                    *reinterpret_cast<uint2*>(p.dk_hi + o) = make_uint2(0u, 0u);
                    *reinterpret_cast<uint2*>(p.dv_hi + o) = make_uint2(0u, 0u);
                    if (p.dk_lo) { *reinterpret_cast<uint2*>(p.dk_lo + o) = make_uint2(0u, 0u); *reinterpret_cast<uint2*>(p.dv_lo + o) = make_uint2(0u, 0u); }
                } else {
                    *reinterpret_cast<float4*>(dkb + (size_t)row * p.ld_dkv + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(dvb + (size_t)row * p.ld_dkv + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        return;
    }

    if (tid == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- operands: fp32 -> bf16 hi / lo, swizzled ----
    {
        float4 vq[4], vo[4], vk[16], vv[16];
        ab_fetch<32>(p.q + (size_t)r0 * p.ldq + h * 64, p.ldq, nq, tid, vq);
        ab_fetch<32>(p.d_ctx + (size_t)r0 * p.ld_dctx + h * 64, p.ld_dctx, nq, tid, vo);
        ab_fetch<128>(p.k + kv_row0 * p.ldkv + h * 64, p.ldkv, nk, tid, vk);
        ab_fetch<128>(p.v + kv_row0 * p.ldkv + h * 64, p.ldkv, nk, tid, vv);
        tr.ev();   // 1 barriers / TMEM allocated, loads issued
        if (p.ctx) {
            float4 vc[4];
            ab_fetch<32>(p.ctx + (size_t)r0 * p.ld_ctx + h * 64, p.ld_ctx, nq, tid, vc);
#pragma unroll
            for (int it = 0; it < 4; ++it) {     // row it * 8 + tid / 16: sixteen lanes hold its 64 columns
                float part = vo[it].x * vc[it].x + vo[it].y * vc[it].y + vo[it].z * vc[it].z + vo[it].w * vc[it].w;
#pragma unroll
                for (int sh = 8; sh > 0; sh >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sh);
                if ((tid & 15) == 0) delta_s[it * 8 + (tid >> 4)] = part;
            }
        }
        ab_put<kX3, 32>(vq, sm + AB_Q, sm + AB_Q + 4096, tid);
        ab_put<kX3, 32>(vo, sm + AB_DO, sm + AB_DO + 4096, tid);
        ab_put<kX3, 128>(vk, sm + AB_K, sm + AB_K + 16384, tid);
        ab_put<kX3, 128>(vv, sm + AB_V, sm + AB_V + 16384, tid);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    tr.ev();       // 2 operands in shared memory

    if (tid == 0) {
        constexpr uint32_t idesc_s = make_idesc(128, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ab_mma3<kX3>(tmem, make_smem_desc(sb + AB_Q) + ko, make_smem_desc(sb + AB_Q + 4096) + ko,
                         make_smem_desc(sb + AB_K) + ko, make_smem_desc(sb + AB_K + 16384) + ko, idesc_s, k == 0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ab_mma3<kX3>(tmem + 128, make_smem_desc(sb + AB_DO) + ko, make_smem_desc(sb + AB_DO + 4096) + ko,
                         make_smem_desc(sb + AB_V) + ko, make_smem_desc(sb + AB_V + 16384) + ko, idesc_s, k == 0);
        }
        tc_commit(bar1);
    }

    if (warp == 0) {
        // ---- softmax statistics, P and dS: lane = query row (backward.cu phase B) ----
        __syncwarp();
        mbar_wait(bar1, 0);
        tc_fence_after();
        tr.ev();   // 3 S, dP ready
        const int i = lane;
        const bool row_ok = i < nq;
        const bool use_watch = (p.mask_kind == NAVC_MASK_CAUSAL) && p.watch != 0 && p.S >= p.watch;
        const int nch = (nk + 31) >> 5;
        // masks as lane-level key bounds (branch-free: the per-element form compiled to ~400 divergent branches per row)
        const bool causal = p.is_self && p.mask_kind == NAVC_MASK_CAUSAL;
        const int key_lo = (causal && use_watch) ? i - p.watch + 1 : 0;     // keys below are filled
        const int key_hi = causal ? i : 127;                                // keys above are filled
        const int key_diag = (p.is_self && p.mask_kind == NAVC_MASK_SELF) ? i : -1;
        auto is_masked = [&](int j) -> bool { return (j < key_lo) | (j > key_hi) | (j == key_diag); };
        // The scores of the row stay in registers (ONE tensor-memory round trip for S; every tcgen05.ld + wait costs
        // ~0.25 us of exposed latency and the first version made 20 of them): sr = scaled / filled score, then exp(score - max).
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < nch) tc_ld32(tmem + (uint32_t)(c * 32), sr[c]);
        tc_wait_ld();
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= nch) break;                         // warp-uniform: self-attention has one 32-key chunk
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int key = c * 32 + j;
                float sv = is_masked(key) ? kMaskFillT : __uint_as_float(sr[c][j]) * 0.125f;
                sv = key < nk ? sv : -INFINITY;
                sr[c][j] = __float_as_uint(sv);
                m = fmaxf(m, sv);
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= nch) break;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e = __expf(__uint_as_float(sr[c][j]) - m);   // exp(-inf) = 0 for the keys that do not exist
                sum += e;
                sr[c][j] = __float_as_uint(e);
            }
        }
        const float inv_sum = 1.0f / sum;
        // sum_j P_ij dP_ij: equals dO_i . O_i (computed while the operands were loaded) when the forward context is given
        float dot = 0.f;
        if (p.ctx) {
            dot = delta_s[i];
        } else {
            for (int c = 0; c < nch; ++c) {
                uint32_t d[32];
                tc_ld32(tmem + (uint32_t)(128 + c * 32), d);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float e = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) e = (cc == c) ? __uint_as_float(sr[cc][j]) : e;
                    dot = fmaf(e, __uint_as_float(d[j]), dot);
                }
            }
            dot *= inv_sum;
        }
        // P and dS, 32 keys at a time; the dP chunk of step c + 1 travels while step c is computed
        uint32_t dbuf[2][32];
        if (nch > 0) tc_ld32(tmem + 128u, dbuf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < nch) {
                tc_wait_ld();
                if (c + 1 < nch) tc_ld32(tmem + (uint32_t)(128 + (c + 1) * 32), dbuf[(c + 1) & 1]);
            }
            uint8_t* pb = sm + AB_P + (c >> 1) * 4096 + i * 128;
            uint8_t* db = sm + AB_DS + (c >> 1) * 4096 + i * 128;
            if (c >= nch) {                        // keys that do not exist: zero columns of P / dS
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int off = ((((c & 1) * 4 + t) ^ (i & 7))) << 4;
                    *reinterpret_cast<uint4*>(pb + off) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(db + off) = make_uint4(0u, 0u, 0u, 0u);
                    if (kX3) {
                        *reinterpret_cast<uint4*>(pb + 8192 + off) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(db + 8192 + off) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                continue;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {          // 8 keys = one 16-byte chunk of the row
                uint32_t ph[4], pl[4], sh[4], sl[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float pv[2], dsv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int j = t * 8 + jj * 2 + u;
                        const int key = c * 32 + j;
                        const bool live = (key < nk) & row_ok;
                        const float pp = live ? __uint_as_float(sr[c][j]) * inv_sum : 0.f;
                        const float ds = pp * (__uint_as_float(dbuf[c & 1][j]) - dot) * 0.125f;
                        pv[u] = pp;
                        dsv[u] = (live & !is_masked(key)) ? ds : 0.f;
                    }
                    split_bf16x2(pv[0], pv[1], ph[jj], pl[jj]);
                    split_bf16x2(dsv[0], dsv[1], sh[jj], sl[jj]);
                }
                const int off = ((((c & 1) * 4 + t) ^ (i & 7))) << 4;
                *reinterpret_cast<uint4*>(pb + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(db + off) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
                if (kX3) {
                    *reinterpret_cast<uint4*>(pb + 8192 + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    *reinterpret_cast<uint4*>(db + 8192 + off) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        tr.ev();   // 4 P, dS written
        if (lane == 0) {
            tc_fence_after();
            constexpr uint32_t idesc_t = make_idesc(128, 64) | (1u << 15) | (1u << 16);   // A and B MN-major
            constexpr uint32_t idesc_q = make_idesc_bmn(128, 64);                         // A K-major, B MN-major
            // dV = P^T dO, dK = dS^T Q: the reduction runs over the 32 query rows (two k-steps of 16 rows)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint64_t ko = (uint64_t)((k * 2048) >> 4);
                ab_mma3<kX3>(tmem, ab_desc_mn(sb + AB_P, 4096) + ko, ab_desc_mn(sb + AB_P + 8192, 4096) + ko,
                             make_smem_desc_mn(sb + AB_DO) + ko, make_smem_desc_mn(sb + AB_DO + 4096) + ko, idesc_t, k == 0);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint64_t ko = (uint64_t)((k * 2048) >> 4);
                ab_mma3<kX3>(tmem + 64, ab_desc_mn(sb + AB_DS, 4096) + ko, ab_desc_mn(sb + AB_DS + 8192, 4096) + ko,
                             make_smem_desc_mn(sb + AB_Q) + ko, make_smem_desc_mn(sb + AB_Q + 4096) + ko, idesc_t, k == 0);
            }
            // dQ = dS K: the reduction runs over the 128 keys (two 64-key blocks of dS x four k-steps; K rows 16 at a time)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t a_off = (uint32_t)((k >> 2) * 4096 + (k & 3) * 32);
                const uint32_t b_off = (uint32_t)(k * 2048);
                ab_mma3<kX3>(tmem + 128, make_smem_desc(sb + AB_DS + a_off), make_smem_desc(sb + AB_DS + 8192 + a_off),
                             make_smem_desc_mn(sb + AB_K + b_off), make_smem_desc_mn(sb + AB_K + 16384 + b_off), idesc_q, k == 0);
            }
            tc_commit(bar2);
        }
        __syncwarp();
    }

    // ---- results: dV / dK rows are the keys (one per thread), dQ rows the queries (warp 0) ----
    mbar_wait(bar2, 0);
    tc_fence_after();
    tr.ev();       // 5 dV, dK, dQ ready
    {
        const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
        const int t = tid;
#pragma unroll
        for (int part = 0; part < 2; ++part) {          // 0: dV -> V buffer, 1: dK -> K buffer
            uint8_t* stg = sm + (part == 0 ? AB_V : AB_K) + t * 256;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tc_ld32(t_lane + (uint32_t)(part * 64 + half * 32), r);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(stg + ((((half * 8 + c) ^ (t & 15))) << 4)) = make_uint4(r[c * 4], r[c * 4 + 1], r[c * 4 + 2], r[c * 4 + 3]);
            }
        }
        if (warp == 0) {
            float* dqr = p.dq + (size_t)(r0 + lane) * p.ld_dq + h * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tc_ld32(t_lane + (uint32_t)(128 + half * 32), r);
                tc_wait_ld();
                if (lane < nq) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(dqr + half * 32 + c * 4) = make_uint4(r[c * 4], r[c * 4 + 1], r[c * 4 + 2], r[c * 4 + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tr.ev();       // 6 staged
    if (p.dk_hi) {
        // split output: the staged fp32 rows leave as bf16 hi (/ lo) -- exactly what navc_transpose_pack would make of them
        for (int idx = tid; idx < nk * 16; idx += AB_THREADS) {
            const int row = idx >> 4, c = idx & 15;
            const int off = row * 256 + (((c ^ (row & 15))) << 4);
            const size_t o = hrow0 + (size_t)row * p.ld_dkv + c * 4;
            uint2 hh, ll;
            split_bf16x4(*reinterpret_cast<const float4*>(sm + AB_V + off), hh, ll);
            *reinterpret_cast<uint2*>(p.dv_hi + o) = hh;
            if (p.dv_lo) *reinterpret_cast<uint2*>(p.dv_lo + o) = ll;
            split_bf16x4(*reinterpret_cast<const float4*>(sm + AB_K + off), hh, ll);
            *reinterpret_cast<uint2*>(p.dk_hi + o) = hh;
            if (p.dk_lo) *reinterpret_cast<uint2*>(p.dk_lo + o) = ll;
        }
        // column sums of this item's rows: thread t < 64 owns column t of dK, thread 64 + t column t of dV
        if (p.dk_colsum) {
            const int col = tid & 63;
            const uint8_t* base = sm + (tid < 64 ? AB_K : AB_V);
            float acc = 0.f;
            for (int row = 0; row < nk; ++row)
                acc += *reinterpret_cast<const float*>(base + row * 256 + ((((col >> 2) ^ (row & 15))) << 4) + (col & 3) * 4);
            atomicAdd((tid < 64 ? p.dk_colsum : p.dv_colsum) + h * 64 + col, acc);
        }
    } else {
        for (int idx = tid; idx < nk * 16; idx += AB_THREADS) {
            const int row = idx >> 4, c = idx & 15;
            const int off = row * 256 + (((c ^ (row & 15))) << 4);
            *reinterpret_cast<float4*>(dvb + (size_t)row * p.ld_dkv + c * 4) = *reinterpret_cast<const float4*>(sm + AB_V + off);
            *reinterpret_cast<float4*>(dkb + (size_t)row * p.ld_dkv + c * 4) = *reinterpret_cast<const float4*>(sm + AB_K + off);
        }
    }
    tr.ev();       // 7 stores issued
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

static int ab_launch(int mode, const AbParams& p, int G, int H, cudaStream_t st, const char* what) {
    NAVC_REQUIRE(tc_ready(), "%s: navc_init() has not been called", what);
    NAVC_REQUIRE(mode == NAVC_TC_BF16 || mode == NAVC_TC_BF16X3, "%s: bad mode %d", what, mode);
    NAVC_REQUIRE(p.ldq % 4 == 0 && p.ldkv % 4 == 0 && p.ld_dctx % 4 == 0 && p.ld_dq % 4 == 0 && p.ld_dkv % 4 == 0 &&
                     ((((uintptr_t)p.q) | ((uintptr_t)p.k) | ((uintptr_t)p.v) | ((uintptr_t)p.d_ctx) | ((uintptr_t)p.dq) |
                       ((uintptr_t)p.dk) | ((uintptr_t)p.dv)) & 15) == 0 &&
                     ((((uintptr_t)p.dk_hi) | ((uintptr_t)p.dk_lo) | ((uintptr_t)p.dv_hi) | ((uintptr_t)p.dv_lo)) & 7) == 0,
                 "%s: operands must be 16-byte aligned with leading dimensions multiple of 4", what);
    static bool ready = false;
    if (!ready) {
        NAVC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        NAVC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        NAVC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        NAVC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ready = true;
    }
    if (mode == NAVC_TC_BF16X3) attn_bwd_tc_kernel<true><<<dim3(G, H), AB_THREADS, AB_SMEM, st>>>(p);
    else attn_bwd_tc_kernel<false><<<dim3(G, H), AB_THREADS, AB_SMEM, st>>>(p);
    return check_launch(what);
}

}  // namespace navc

using namespace navc;

// Debug: install (or remove, buf == NULL) the timeline buffer: 16 uint64 per CTA of the grid (N x H CTAs).
extern "C" int navc_debug_trace_attn_bwd(void* buf) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(navc::ab_trace_buf, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}

extern "C" int navc_self_attention_bwd_tc(int mode, const float* qkv, int ld, const int32_t* seq_off, int N, int S, int D, int H,
                                          int mask_kind, int watch, const float* d_ctx, const float* ctx, float* d_qkv, void* stream) {
    NAVC_REQUIRE(qkv && seq_off && d_ctx && d_qkv, "navc_self_attention_bwd_tc: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 32 && H > 0 && D == H * 64 && ld >= 3 * D, "navc_self_attention_bwd_tc: needs dk == 64 and S <= 32");
    AbParams p = {};
    p.q = qkv; p.ldq = ld; p.k = qkv + D; p.v = qkv + 2 * D; p.ldkv = ld;
    p.d_ctx = d_ctx; p.ld_dctx = D; p.ctx = ctx; p.ld_ctx = D;
    p.dq = d_qkv; p.ld_dq = ld; p.dk = d_qkv + D; p.dv = d_qkv + 2 * D; p.ld_dkv = ld;
    p.seq_off = seq_off; p.is_self = 1; p.E = 0; p.S = S; p.mask_kind = mask_kind; p.watch = watch;
    return ab_launch(mode, p, N, H, as_stream(stream), "navc_self_attention_bwd_tc");
}

extern "C" int navc_cross_attention_bwd_tc(int mode, const float* q, int ldq, const float* kv, int ldkv, const int32_t* seq_off,
                                           int N, int S, int E, int D, int H, const float* d_ctx, const float* ctx, float* d_q,
                                           int ld_dq, float* d_kv, int ld_dkv, void* stream) {
    NAVC_REQUIRE(q && kv && seq_off && d_ctx && d_q && d_kv, "navc_cross_attention_bwd_tc: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 32 && E > 0 && E <= 128 && H > 0 && D == H * 64,
                 "navc_cross_attention_bwd_tc: needs dk == 64, S <= 32 and E <= 128");
    AbParams p = {};
    p.q = q; p.ldq = ldq; p.k = kv; p.v = kv + D; p.ldkv = ldkv;
    p.d_ctx = d_ctx; p.ld_dctx = D; p.ctx = ctx; p.ld_ctx = D;
    p.dq = d_q; p.ld_dq = ld_dq; p.dk = d_kv; p.dv = d_kv + D; p.ld_dkv = ld_dkv;
    p.seq_off = seq_off; p.is_self = 0; p.E = E; p.S = S; p.mask_kind = 0; p.watch = 0;
    return ab_launch(mode, p, N, H, as_stream(stream), "navc_cross_attention_bwd_tc");
}

// Text -> video gradients with dK / dV written as the bf16 hi (/ lo) operand pair of the K|V projection's gradient GEMMs
// (d_kv_hi / d_kv_lo [N * E, ld_dkv], K at column 0 and V at column D of the given pointers) plus their column sums
// (kv_colsum [2 D]: the bias gradient, accumulated): the fp32 d_kv round trip (write, re-read, split) does not exist.
extern "C" int navc_cross_attention_bwd_tc_split(int mode, const float* q, int ldq, const float* kv, int ldkv, const int32_t* seq_off,
                                                 int N, int S, int E, int D, int H, const float* d_ctx, const float* ctx, float* d_q,
                                                 int ld_dq, uint16_t* d_kv_hi, uint16_t* d_kv_lo, int ld_dkv, float* kv_colsum,
                                                 void* stream) {
    NAVC_REQUIRE(q && kv && seq_off && d_ctx && d_q && d_kv_hi, "navc_cross_attention_bwd_tc_split: null pointer");
    NAVC_REQUIRE(mode == NAVC_TC_BF16 || d_kv_lo, "navc_cross_attention_bwd_tc_split: the split mode needs the lo output");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 32 && E > 0 && E <= 128 && H > 0 && D == H * 64 && ld_dkv % 4 == 0,
                 "navc_cross_attention_bwd_tc_split: needs dk == 64, S <= 32, E <= 128 and ld_dkv %% 4 == 0");
    AbParams p = {};
    p.q = q; p.ldq = ldq; p.k = kv; p.v = kv + D; p.ldkv = ldkv;
    p.d_ctx = d_ctx; p.ld_dctx = D; p.ctx = ctx; p.ld_ctx = D;
    p.dq = d_q; p.ld_dq = ld_dq; p.dk = nullptr; p.dv = nullptr; p.ld_dkv = ld_dkv;
    p.dk_hi = d_kv_hi; p.dv_hi = d_kv_hi + D;
    p.dk_lo = d_kv_lo; p.dv_lo = d_kv_lo ? d_kv_lo + D : nullptr;
    p.dk_colsum = kv_colsum; p.dv_colsum = kv_colsum ? kv_colsum + D : nullptr;
    p.seq_off = seq_off; p.is_self = 0; p.E = E; p.S = S; p.mask_kind = 0; p.watch = 0;
    return ab_launch(mode, p, N, H, as_stream(stream), "navc_cross_attention_bwd_tc_split");
}
