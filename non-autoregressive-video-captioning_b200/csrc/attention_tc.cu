// tcgen05 attention cores (dk = 64): token self-attention and text->video cross-attention.
//
// One CTA = one 128-query-row tile of one head:
//   TMA (128B swizzle) loads Q[128,64], K[128,64], V[128,64] (bf16 hi, and lo in the split mode)
//   tcgen05.mma  S[128,128] = Q K^T          (K-major operands)            -> TMEM columns [0,128)
//   4 softmax warps: tcgen05.ld S row per thread, 1/sqrt(dk) scale, masks derived in-kernel
//     (-1e7 fill as models/bert.py:157-161; keys of other sequences / beyond E excluded),
//     e = exp(s - max) split to bf16 hi/lo and written to shared memory in the canonical
//     K-major 128B-swizzle layout (the A operand of the second MMA; aliases the Q/K buffers)
//   tcgen05.mma  O[128,64] = P V             (V consumed MN-major, as TMA loaded it) -> TMEM [0,64) (reuses S)
//   epilogue: tcgen05.ld O row per thread, * 1/sum, bf16 hi/lo (+fp32) context rows.
// Split mode issues hi*hi + hi*lo + lo*hi for both products (fp32-like accuracy).
//
// Self-attention packs 4 sequences (S <= 32) into one tile: rows z*4S + [0,128) of qkv[R,3D];
// the block-diagonal structure is enforced by the mask.  Cross-attention tiles the group*S query
// rows of one video by 128 and reads that video's E <= 128 keys (K/V projected once per video).
#include "tc_common.cuh"

namespace navc {

constexpr int kAtThreads = 192;
constexpr int kAtTile = 128 * 64 * 2;  // one [128,64] bf16 tile = 16 KB
constexpr float kMaskFillTc = -10e6f;   // models/bert.py:161

template <bool kX3> struct AtCfg {
    static constexpr int kParts = kX3 ? 2 : 1;
    // [Q hi, (Q lo), K hi, (K lo)] then [V hi, (V lo)]; P (2 panels per part) aliases the Q/K region
    static constexpr int kQKBytes = 2 * kParts * kAtTile;
    static constexpr int kVBytes = kParts * kAtTile;
    static constexpr int kUsedBytes = kQKBytes + kVBytes + 1024 /*align*/ + 256 /*barriers, flags*/;
    // at most 4 CTAs per SM (4 x 128 TMEM columns): keep the footprint above a fifth of shared memory
    static constexpr int kSmemBytes = kUsedBytes > 48 * 1024 ? kUsedBytes : 48 * 1024;
};

struct AtParams {
    int q_col, k_col, v_col;      // column offsets (elements) of this head's Q/K/V inside the maps: + h*64
    int rows_per_tile;            // query rows owned by a tile (4S for self, 128 for cross)
    int nq_per_owner;             // cross: group*S query rows per video; self: unused
    int tiles_per_owner;          // cross: ceil(group*S/128); self: 1
    int keys_per_owner;           // cross: E; self: unused
    int total_q_rows;             // R
    int S;                        // tokens per sequence
    int is_self, mask_kind, watch;
    const int64_t* tokens;        // self: [N,S]
    const int32_t* seq_off;       // packed rows: [n_seq + 1] row offsets (NULL = padded layout)
    int n_seq, group;             // packed rows: sequences in total, sequences per video (cross)
    int D;
    float* ctx_f32; uint16_t* ctx_hi; uint16_t* ctx_lo;
};

template <bool kX3>
__global__ void __launch_bounds__(kAtThreads, kX3 ? 2 : 4)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
               const __grid_constant__ CUtensorMap map_kv_hi, const __grid_constant__ CUtensorMap map_kv_lo,
               AtParams p) {
    // Packed self-attention (p.seq_off && p.is_self): a tile holds 4 sequences in 32-row slots; the maps
    // then have 32-row boxes and every operand tile is four TMA loads at the sequences' row offsets.
    using Cfg = AtCfg<kX3>;
    constexpr int P = Cfg::kParts;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // tile addresses
    const uint32_t sQ_hi = smem_base, sQ_lo = smem_base + kAtTile;                    // lo only if kX3
    const uint32_t sK_hi = smem_base + P * kAtTile, sK_lo = sK_hi + kAtTile;
    const uint32_t sV_hi = smem_base + Cfg::kQKBytes, sV_lo = sV_hi + kAtTile;
    // P panels (alias Q/K): part x panel, each [128 rows, 64 keys] bf16
    const uint32_t sP = smem_base;
    uint8_t* gP = smem_gen;
    const uint32_t bar_base = smem_base + Cfg::kQKBytes + Cfg::kVBytes;
    const uint32_t bar_qk = bar_base, bar_v = bar_base + 8, bar_s = bar_base + 16, bar_p = bar_base + 24, bar_o = bar_base + 32;
    uint8_t* tail = smem_gen + Cfg::kQKBytes + Cfg::kVBytes;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 40);
    uint8_t* keypad = tail + 64;  // [128] 1 = key is PAD (self-attention)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y;

    // tile geometry
    const bool packed = p.seq_off != nullptr;
    int q_row0, k_row0, nq_valid, n_keys;
    int seq0 = 0;              // packed self: first sequence of this tile
    int slot_len = 0;          // packed self: length of the sequence whose 32-row slot this thread's row is in
    int slot_row0 = 0;         // packed self: first packed row of that sequence
    if (p.is_self && packed) {
        seq0 = blockIdx.x * 4;
        q_row0 = k_row0 = 0;   // unused
        nq_valid = 128;
        n_keys = 128;
    } else if (p.is_self) {
        q_row0 = blockIdx.x * p.rows_per_tile;
        k_row0 = q_row0;
        nq_valid = min(p.rows_per_tile, p.total_q_rows - q_row0);
        n_keys = nq_valid;
    } else {
        const int g = blockIdx.x / p.tiles_per_owner, z = blockIdx.x % p.tiles_per_owner;
        if (packed) {
            const int r0 = __ldg(p.seq_off + g * p.group), r1 = __ldg(p.seq_off + min((g + 1) * p.group, p.n_seq));
            q_row0 = r0 + z * 128;
            nq_valid = min(128, r1 - q_row0);
            if (nq_valid <= 0) return;  // this video has fewer query rows than the launch maximum (whole CTA exits)
        } else {
            q_row0 = g * p.nq_per_owner + z * 128;
            nq_valid = min(128, p.nq_per_owner - z * 128);
        }
        k_row0 = g * p.keys_per_owner;
        n_keys = p.keys_per_owner;
    }

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();  // tokens / Q / K / V below may still be in flight in the preceding kernel
    if (warp >= 2) {
        const int r = (warp - 2) * 32 + lane;
        uint8_t kp = 0;
        if (p.is_self && packed) {
            const int seq = seq0 + (r >> 5), j = r & 31;
            int len = 0;
            if (seq < p.n_seq) len = __ldg(p.seq_off + seq + 1) - __ldg(p.seq_off + seq);
            if (p.tokens && j < len) kp = p.tokens[(size_t)seq * p.S + j] == NAVC_PAD ? 1 : 0;
            // the slot this thread serves in the softmax / epilogue phases is its TMEM lane quarter (warp & 3)
            const int myseq = seq0 + (warp & 3);
            if (myseq < p.n_seq) {
                slot_row0 = __ldg(p.seq_off + myseq);
                slot_len = __ldg(p.seq_off + myseq + 1) - slot_row0;
            }
        } else if (p.is_self && p.tokens && r < nq_valid) {
            kp = p.tokens[(size_t)q_row0 + r] == NAVC_PAD ? 1 : 0;
        }
        keypad[r] = kp;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base;  // O reuses the S columns (S is dead once P is in smem)

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_qk, 2 * P * kAtTile);
            if (p.is_self && packed) {
                // four 32-row boxes per operand tile: slot i <- rows of sequence seq0 + i (4 KB apart: 8-row swizzle atoms)
                for (int i = 0; i < 4; ++i) {
                    const int row = __ldg(p.seq_off + min(seq0 + i, p.n_seq - 1));
                    const uint32_t so = (uint32_t)(i * 4096);
                    tma_load_2d(sQ_hi + so, &map_q_hi, bar_qk, p.q_col + h * 64, row);
                    tma_load_2d(sK_hi + so, &map_kv_hi, bar_qk, p.k_col + h * 64, row);
                    if (kX3) {
                        tma_load_2d(sQ_lo + so, &map_q_lo, bar_qk, p.q_col + h * 64, row);
                        tma_load_2d(sK_lo + so, &map_kv_lo, bar_qk, p.k_col + h * 64, row);
                    }
                }
                mbar_expect_tx(bar_v, P * kAtTile);
                for (int i = 0; i < 4; ++i) {
                    const int row = __ldg(p.seq_off + min(seq0 + i, p.n_seq - 1));
                    const uint32_t so = (uint32_t)(i * 4096);
                    tma_load_2d(sV_hi + so, &map_kv_hi, bar_v, p.v_col + h * 64, row);
                    if (kX3) tma_load_2d(sV_lo + so, &map_kv_lo, bar_v, p.v_col + h * 64, row);
                }
            } else {
                tma_load_2d(sQ_hi, &map_q_hi, bar_qk, p.q_col + h * 64, q_row0);
                tma_load_2d(sK_hi, &map_kv_hi, bar_qk, p.k_col + h * 64, k_row0);
                if (kX3) {
                    tma_load_2d(sQ_lo, &map_q_lo, bar_qk, p.q_col + h * 64, q_row0);
                    tma_load_2d(sK_lo, &map_kv_lo, bar_qk, p.k_col + h * 64, k_row0);
                }
                mbar_expect_tx(bar_v, P * kAtTile);
                tma_load_2d(sV_hi, &map_kv_hi, bar_v, p.v_col + h * 64, k_row0);
                if (kX3) tma_load_2d(sV_lo, &map_kv_lo, bar_v, p.v_col + h * 64, k_row0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- S = Q K^T ----
            mbar_wait(bar_qk, 0);
            tc_fence_after();
            constexpr uint32_t idesc_s = make_idesc(128, 128);
            const uint64_t dq_hi = make_smem_desc(sQ_hi), dq_lo = make_smem_desc(sQ_lo);
            const uint64_t dk_hi = make_smem_desc(sK_hi), dk_lo = make_smem_desc(sK_lo);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
                if (kX3) {
                    tc_mma_bf16(tS, dq_lo + koff, dk_hi + koff, idesc_s, k ? 1u : 0u);
                    tc_mma_bf16(tS, dq_hi + koff, dk_lo + koff, idesc_s, 1u);
                    tc_mma_bf16(tS, dq_hi + koff, dk_hi + koff, idesc_s, 1u);
                } else {
                    tc_mma_bf16(tS, dq_hi + koff, dk_hi + koff, idesc_s, k ? 1u : 0u);
                }
            }
            tc_commit(bar_s);
            // ---- O = P V ----
            mbar_wait(bar_v, 0);
            mbar_wait(bar_p, 0);
            tc_fence_after();
            constexpr uint32_t idesc_o = make_idesc_bmn(128, 64);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // A: P panel k/4 (64 keys each, 16 KB), 32 B per k-step inside the panel
                const uint32_t pa = sP + (uint32_t)((k >> 2) * kAtTile + (k & 3) * 32);
                const uint64_t dp_hi = make_smem_desc(pa);
                const uint64_t dp_lo = make_smem_desc(pa + 2 * kAtTile);
                // B: V rows 16k..16k+15 (MN-major): 2 KB per k-step
                const uint64_t dv_hi = make_smem_desc_mn(sV_hi + (uint32_t)(k * 2048));
                const uint64_t dv_lo = make_smem_desc_mn(sV_lo + (uint32_t)(k * 2048));
                if (kX3) {
                    tc_mma_bf16(tO, dp_lo, dv_hi, idesc_o, k ? 1u : 0u);
                    tc_mma_bf16(tO, dp_hi, dv_lo, idesc_o, 1u);
                    tc_mma_bf16(tO, dp_hi, dv_hi, idesc_o, 1u);
                } else {
                    tc_mma_bf16(tO, dp_hi, dv_hi, idesc_o, k ? 1u : 0u);
                }
            }
            tc_commit(bar_o);
        }
    } else {
        // ===================== softmax + epilogue (warps 2..5) =====================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                 // tile row = TMEM lane
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        // Per-thread 32-bit key masks per 32-key chunk: `vis` = keys this query may see at all (its own
        // sequence / the owner's E keys), `fill` = visible keys whose score is replaced by -1e7
        // (PAD keys, causal / diagonal masks; models/bert.py:157-161).  Chunks [c_lo, c_hi] (warp
        // uniform) are the only ones with a visible key for some row of this warp; P is zero elsewhere.
        int own_lo = 0, own_hi = n_keys, ipos = 0, c_lo = 0, c_hi = 3;
        if (p.is_self && packed) {
            own_lo = quarter * 32;             // one sequence per warp / per 32-key chunk
            own_hi = own_lo + slot_len;
            ipos = lane;
            c_lo = c_hi = quarter;
        } else if (p.is_self) {
            const int si = r / p.S;
            own_lo = si * p.S;
            own_hi = min(own_lo + p.S, n_keys);
            ipos = r - own_lo;
            const int s_first = (quarter * 32) / p.S, s_last = (quarter * 32 + 31) / p.S;
            c_lo = min(3, (s_first * p.S) >> 5);
            c_hi = min(3, ((s_last + 1) * p.S - 1) >> 5);
        } else {
            c_hi = min(3, (n_keys - 1) >> 5);
        }
        const bool use_watch = (p.mask_kind == NAVC_MASK_CAUSAL) && p.watch != 0 && p.S >= p.watch;
        auto range_mask = [](int lo, int hi, int c) -> uint32_t {  // bits of keys [lo, hi) inside chunk c
            const int a = max(lo - c * 32, 0), b = min(hi - c * 32, 32);
            if (b <= a) return 0u;
            const uint32_t upto_b = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
            return upto_b & ~((1u << a) - 1u);
        };
        uint32_t* padw = reinterpret_cast<uint32_t*>(keypad + 128);  // PAD-key bits per 32-key chunk
        if (p.is_self) {
            const uint32_t mine = __ballot_sync(0xffffffffu, keypad[r] != 0);
            if (lane == 0) padw[quarter] = mine;
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four softmax warps only
        }
        auto chunk_masks = [&](int c, uint32_t& vm, uint32_t& fm) {
            vm = range_mask(own_lo, own_hi, c);
            fm = 0u;
            if (p.is_self) {
                uint32_t f = padw[c];
                if (p.mask_kind == NAVC_MASK_CAUSAL) {
                    f |= range_mask(own_lo + ipos + 1, own_hi, c);                           // future keys
                    if (use_watch) f |= range_mask(own_lo, own_lo + ipos - p.watch + 1, c);  // keys older than the window
                }
                if (p.mask_kind == NAVC_MASK_SELF) f |= range_mask(own_lo + ipos, own_lo + ipos + 1, c);
                fm = f & vm;
            }
        };
        const float scale = 0.125f;  // 1/sqrt(dk), dk = 64: exact power of two == the reference's division by sqrt(dk)
        constexpr float kLog2e = 1.4426950408889634f;

        mbar_wait(bar_s, 0);
        tc_fence_after();
        // pass 1: row maximum over the visible keys
        float m = -INFINITY;
#pragma unroll 1
        for (int c = c_lo; c <= c_hi; ++c) {
            uint32_t v[32];
            tc_ld32(tS + t_lane + (uint32_t)(c * 32), v);
            tc_wait_ld();
            uint32_t vis, fill;
            chunk_masks(c, vis, fill);
            const uint32_t vm = vis & ~fill;
            float mc = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if ((vm >> j) & 1u) mc = fmaxf(mc, __uint_as_float(v[j]));
            m = fmaxf(m, mc * scale);               // scale > 0: max commutes with the scaling
            if (fill) m = fmaxf(m, kMaskFillTc);
        }
        // pass 2: e = exp(s - m), row sum, P -> shared memory (bf16 hi/lo, K-major 128B swizzle)
        float sum = 0.f;
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const float m2 = m * kLog2e, sc2 = scale * kLog2e;
        const float e_fill = fast_exp2((kMaskFillTc - m) * kLog2e);  // 0 unless every visible key is filled
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint8_t* panel_hi = gP + (c >> 1) * kAtTile + row_off;
            if (c < c_lo || c > c_hi) {  // warp-uniform: no visible key in this chunk for any row of the warp
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t ch = (uint32_t)(((c & 1) * 4 + i) ^ (r & 7)) * 16;
                    *reinterpret_cast<uint4*>(panel_hi + ch) = make_uint4(0u, 0u, 0u, 0u);
                    if (kX3) *reinterpret_cast<uint4*>(panel_hi + 2 * kAtTile + ch) = make_uint4(0u, 0u, 0u, 0u);
                }
                continue;
            }
            uint32_t v[32];
            tc_ld32(tS + t_lane + (uint32_t)(c * 32), v);
            tc_wait_ld();
            uint32_t vm, fm;
            chunk_masks(c, vm, fm);
            uint32_t hi_w[16], lo_w[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float e2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float e = fast_exp2(fmaf(__uint_as_float(v[j + u]), sc2, -m2));
                    if ((fm >> (j + u)) & 1u) e = e_fill;
                    if (!((vm >> (j + u)) & 1u)) e = 0.f;
                    e2[u] = e;
                    sum += e;
                }
                if (kX3) {
                    split_bf16x2(e2[0], e2[1], hi_w[j >> 1], lo_w[j >> 1]);
                } else {
                    const __nv_bfloat162 hb = __floats2bfloat162_rn(e2[0], e2[1]);
                    hi_w[j >> 1] = *reinterpret_cast<const uint32_t*>(&hb);
                }
            }
            // 32 keys = 64 B = four 16-byte chunks of panel c/2, chunk index (c%2)*4 + i
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t ch = (uint32_t)(((c & 1) * 4 + i) ^ (r & 7)) * 16;
                *reinterpret_cast<uint4*>(panel_hi + ch) = make_uint4(hi_w[i * 4], hi_w[i * 4 + 1], hi_w[i * 4 + 2], hi_w[i * 4 + 3]);
                if (kX3)
                    *reinterpret_cast<uint4*>(panel_hi + 2 * kAtTile + ch) = make_uint4(lo_w[i * 4], lo_w[i * 4 + 1], lo_w[i * 4 + 2], lo_w[i * 4 + 3]);
            }
        }
        if (p.is_self && packed) {
            // slot rows beyond the sequence hold whatever follows it in memory: P is 0 there, but 0 * NaN
            // would poison O, so those V rows are cleared (each key row = one 128-byte swizzled row)
            mbar_wait(bar_v, 0);
            if (lane >= slot_len) {
                uint8_t* vrow = smem_gen + Cfg::kQKBytes + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    *reinterpret_cast<uint4*>(vrow + i * 16) = make_uint4(0u, 0u, 0u, 0u);
                    if (kX3) *reinterpret_cast<uint4*>(vrow + kAtTile + i * 16) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);

        // epilogue: O row * 1/sum
        mbar_wait(bar_o, 0);
        tc_fence_after();
        const float inv = 1.0f / sum;
        const bool store = (p.is_self && packed) ? (lane < slot_len) : (r < nq_valid);
        const size_t o = ((p.is_self && packed) ? (size_t)(slot_row0 + lane) : ((size_t)q_row0 + r)) * p.D + h * 64;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tc_ld32(tO + t_lane + (uint32_t)(c * 32), v);
            tc_wait_ld();
            if (store) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {  // 8 columns per step: 16-byte bf16 stores
                    const float4 f0 = make_float4(__uint_as_float(v[g8 * 8]) * inv, __uint_as_float(v[g8 * 8 + 1]) * inv,
                                                  __uint_as_float(v[g8 * 8 + 2]) * inv, __uint_as_float(v[g8 * 8 + 3]) * inv);
                    const float4 f1 = make_float4(__uint_as_float(v[g8 * 8 + 4]) * inv, __uint_as_float(v[g8 * 8 + 5]) * inv,
                                                  __uint_as_float(v[g8 * 8 + 6]) * inv, __uint_as_float(v[g8 * 8 + 7]) * inv);
                    const size_t oo = o + c * 32 + g8 * 8;
                    if (p.ctx_f32) {
                        *reinterpret_cast<float4*>(p.ctx_f32 + oo) = f0;
                        *reinterpret_cast<float4*>(p.ctx_f32 + oo + 4) = f1;
                    }
                    if (p.ctx_hi) {
                        uint2 h0, l0, h1, l1;
                        split_bf16x4(f0, h0, l0);
                        split_bf16x4(f1, h1, l1);
                        *reinterpret_cast<uint4*>(p.ctx_hi + oo) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                        if (p.ctx_lo) *reinterpret_cast<uint4*>(p.ctx_lo + oo) = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

static bool g_at_ready = false;
static int at_init() {
    if (g_at_ready) return 0;
    NAVC_CUDA(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<false>::kSmemBytes));
    NAVC_CUDA(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<true>::kSmemBytes));
    g_at_ready = true;
    return 0;
}

static int launch_attn_tc(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq, int q_cols, int q_rows,
                          const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv, int kv_cols, int kv_rows,
                          const AtParams& p, int n_tiles, int H, cudaStream_t st, const char* what, int box_rows = 128) {
    NAVC_REQUIRE(tc_ready(), "%s: navc_init() has not been called", what);
    NAVC_REQUIRE(mode == NAVC_TC_BF16 || mode == NAVC_TC_BF16X3, "%s: bad mode %d", what, mode);
    NAVC_REQUIRE(q_hi && kv_hi && (mode == NAVC_TC_BF16 || (q_lo && kv_lo)), "%s: null operand", what);
    NAVC_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && p.D % 8 == 0, "%s: leading dimensions must be multiples of 8", what);
    NAVC_REQUIRE((((uintptr_t)q_hi | (uintptr_t)q_lo | (uintptr_t)kv_hi | (uintptr_t)kv_lo) & 15) == 0,
                 "%s: operands must be 16-byte aligned", what);
    if (at_init()) return 2;
    CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo;
    const int kv_box = p.is_self ? box_rows : 128;
    if (tc_make_map(&mq_hi, q_hi, q_rows, q_cols, ldq, box_rows)) return 1;
    if (tc_make_map(&mk_hi, kv_hi, kv_rows, kv_cols, ldkv, kv_box)) return 1;
    if (mode == NAVC_TC_BF16X3) {
        if (tc_make_map(&mq_lo, q_lo, q_rows, q_cols, ldq, box_rows)) return 1;
        if (tc_make_map(&mk_lo, kv_lo, kv_rows, kv_cols, ldkv, kv_box)) return 1;
    } else {
        mq_lo = mq_hi;
        mk_lo = mk_hi;
    }
    dim3 grid(n_tiles, H);
    if (mode == NAVC_TC_BF16X3)
        NAVC_CUDA(launch_pdl(attn_tc_kernel<true>, grid, dim3(kAtThreads), AtCfg<true>::kSmemBytes, st, mq_hi, mq_lo, mk_hi, mk_lo, p));
    else
        NAVC_CUDA(launch_pdl(attn_tc_kernel<false>, grid, dim3(kAtThreads), AtCfg<false>::kSmemBytes, st, mq_hi, mq_lo, mk_hi, mk_lo, p));
    return check_launch(what);
}

}  // namespace navc

using namespace navc;

extern "C" int navc_self_attention_tc(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                      const int64_t* tokens, int N, int S, int D, int H, int mask_kind, int watch,
                                      float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream) {
    NAVC_REQUIRE(tokens && (ctx_f32 || ctx_hi), "navc_self_attention_tc: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 32 && H > 0 && D == H * 64 && ld >= 3 * D,
                 "navc_self_attention_tc: needs dk == 64 and S <= 32 (N=%d S=%d D=%d H=%d)", N, S, D, H);
    NAVC_REQUIRE(mask_kind >= 0 && mask_kind <= 2, "navc_self_attention_tc: bad mask kind");
    const int R = N * S, rpt = 4 * S;
    AtParams p = {};
    p.q_col = 0; p.k_col = D; p.v_col = 2 * D;
    p.rows_per_tile = rpt; p.total_q_rows = R; p.S = S; p.is_self = 1; p.mask_kind = mask_kind; p.watch = watch;
    p.tokens = tokens; p.D = D; p.ctx_f32 = ctx_f32; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo;
    return launch_attn_tc(mode, qkv_hi, qkv_lo, ld, 3 * D, R, qkv_hi, qkv_lo, ld, 3 * D, R, p, (R + rpt - 1) / rpt, H,
                          as_stream(stream), "navc_self_attention_tc");
}

extern "C" int navc_cross_attention_tc(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                       const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv, int N, int S, int E,
                                       int D, int H, int group, float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo,
                                       void* stream) {
    NAVC_REQUIRE(ctx_f32 || ctx_hi, "navc_cross_attention_tc: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && E > 0 && E <= 128 && H > 0 && D == H * 64 && group >= 1 && N % group == 0 &&
                     ldq >= D && ldkv >= 2 * D,
                 "navc_cross_attention_tc: needs dk == 64 and E <= 128 (N=%d S=%d E=%d D=%d H=%d)", N, S, E, D, H);
    const int G = N / group, nq = group * S, tpo = (nq + 127) / 128;
    AtParams p = {};
    p.q_col = 0; p.k_col = 0; p.v_col = D;
    p.rows_per_tile = 128; p.nq_per_owner = nq; p.tiles_per_owner = tpo; p.keys_per_owner = E;
    p.total_q_rows = N * S; p.S = S; p.is_self = 0; p.D = D;
    p.ctx_f32 = ctx_f32; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo;
    return launch_attn_tc(mode, q_hi, q_lo, ldq, D, N * S, kv_hi, kv_lo, ldkv, 2 * D, G * E, p, G * tpo, H,
                          as_stream(stream), "navc_cross_attention_tc");
}

extern "C" int navc_self_attention_tc_packed(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                             const int64_t* tokens, const int32_t* seq_off, int N, int S, int D, int H,
                                             int mask_kind, int watch, float* ctx_f32, uint16_t* ctx_hi,
                                             uint16_t* ctx_lo, void* stream) {
    return navc_self_attention_tc_rows(mode, qkv_hi, qkv_lo, ld, tokens, seq_off, N * S, N, S, D, H, mask_kind, watch,
                                       ctx_f32, ctx_hi, ctx_lo, stream);
}

extern "C" int navc_self_attention_tc_rows(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                           const int64_t* tokens, const int32_t* seq_off, int rows, int N, int S, int D,
                                           int H, int mask_kind, int watch, float* ctx_f32, uint16_t* ctx_hi,
                                           uint16_t* ctx_lo, void* stream) {
    NAVC_REQUIRE(tokens && seq_off && (ctx_f32 || ctx_hi) && rows > 0, "navc_self_attention_tc_packed: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 32 && H > 0 && D == H * 64 && ld >= 3 * D,
                 "navc_self_attention_tc_packed: needs dk == 64 and S <= 32 (N=%d S=%d D=%d H=%d)", N, S, D, H);
    NAVC_REQUIRE(mask_kind >= 0 && mask_kind <= 2, "navc_self_attention_tc_packed: bad mask kind");
    const int R = rows;  // rows the packed buffers hold (TMA zero-fills beyond)
    AtParams p = {};
    p.q_col = 0; p.k_col = D; p.v_col = 2 * D;
    p.rows_per_tile = 128; p.total_q_rows = R; p.S = S; p.is_self = 1; p.mask_kind = mask_kind; p.watch = watch;
    p.tokens = tokens; p.seq_off = seq_off; p.n_seq = N; p.group = 1; p.D = D;
    p.ctx_f32 = ctx_f32; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo;
    return launch_attn_tc(mode, qkv_hi, qkv_lo, ld, 3 * D, R, qkv_hi, qkv_lo, ld, 3 * D, R, p, (N + 3) / 4, H,
                          as_stream(stream), "navc_self_attention_tc_packed", 32);
}

extern "C" int navc_cross_attention_tc_packed(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                              const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv,
                                              const int32_t* seq_off, int N, int S, int E, int D, int H, int group,
                                              float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream) {
    return navc_cross_attention_tc_rows(mode, q_hi, q_lo, ldq, kv_hi, kv_lo, ldkv, seq_off, N * S, N, S, E, D, H, group,
                                        ctx_f32, ctx_hi, ctx_lo, stream);
}

extern "C" int navc_cross_attention_tc_rows(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                            const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv,
                                            const int32_t* seq_off, int rows, int N, int S, int E, int D, int H, int group,
                                            float* ctx_f32, uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream) {
    NAVC_REQUIRE(seq_off && (ctx_f32 || ctx_hi) && rows > 0, "navc_cross_attention_tc_packed: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && E > 0 && E <= 128 && H > 0 && D == H * 64 && group >= 1 && N % group == 0 &&
                     ldq >= D && ldkv >= 2 * D,
                 "navc_cross_attention_tc_packed: needs dk == 64 and E <= 128 (N=%d S=%d E=%d D=%d H=%d)", N, S, E, D, H);
    const int G = N / group, nq = group * S, tpo = (nq + 127) / 128;
    AtParams p = {};
    p.q_col = 0; p.k_col = 0; p.v_col = D;
    p.rows_per_tile = 128; p.nq_per_owner = nq; p.tiles_per_owner = tpo; p.keys_per_owner = E;
    p.total_q_rows = N * S; p.S = S; p.is_self = 0; p.D = D;
    p.seq_off = seq_off; p.n_seq = N; p.group = group;
    p.ctx_f32 = ctx_f32; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo;
    return launch_attn_tc(mode, q_hi, q_lo, ldq, D, rows, kv_hi, kv_lo, ldkv, 2 * D, G * E, p, G * tpo, H,
                          as_stream(stream), "navc_cross_attention_tc_packed");
}
