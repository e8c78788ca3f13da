// tcgen05 attention cores over packed rows, second generation (dk = 64): token self-attention and text -> video
// cross-attention (models/bert.py:139-179) as ONE persistent, warp-specialised, two-stage pipelined kernel.
//
// What the first-generation kernel (attention_tc.cu) did per CTA, serially: TMA loads -> QK^T -> softmax -> PV ->
// stores, one (128-row tile, head) per CTA, 1536-2048 CTAs per launch, two of them resident per SM: 43-48 us per launch
// at config 2 (21 % of the step) at 16-21 % tensor-pipe activity and 0.28-0.37 of the HBM peak, bound by neither roof.
// The self-attention tiles held 4 sequences in 32-row slots (13.6 real rows on average: 2.3x over-fetch) and every
// context row was written row-per-lane (32 distinct lines per store instruction).
//
// Here one CTA per SM stays resident and walks the (tile, head) work items; a work item moves through a two-stage ring
// (shared memory AND tensor memory), so the loads of item i+1 / i+2, the QK^T of item i+1, the softmax of item i and
// the stores of item i-1 overlap:
//   warp 0      producer: derives each item's geometry, publishes it, issues its TMA boxes (32 rows x 64 columns,
//               only as many as the item has rows / keys)
//   warp 1      MMA issuer: S = Q K^T (UMMA 128x128x16) of item i, then O = P V (UMMA 128x64x16, V consumed MN-major) of
//               item i-1
//   warps 2-5 / 6-9   two softmax + epilogue groups, alternating items: row maximum, then exp / row sum over S
//               (32-key chunks straight from tensor memory), P -> TENSOR memory (tcgen05.st, packed bf16 hi / lo: the
//               A operand of the second product, which never touches shared memory), then O * 1/sum
//               -> swizzled staging rows -> full 128-byte lines to global memory (4 rows per store instruction)
// Self-attention tiles are 96-row windows of the packed row space (every sequence that STARTS in the window; at most
// 127 rows): block-diagonal masks from the per-row sequence bounds, no slot padding.  Cross-attention tiles are a
// video's candidate rows (contiguous in the packed layout) against that video's E <= 128 keys.
// Split mode (bf16x3) issues hi*hi + hi*lo + lo*hi for both products.  Masks as models/bert.py:157-161 (-1e7 fill after
// the 1/sqrt(dk) scale) and models/Decoder.py:9-39 (key padding, causal (+ watch), diagonal).
#include <stdlib.h>

#include "tc_common.cuh"

namespace navc {

constexpr int A2_THREADS = 320;
constexpr int A2_TILE = 128 * 64 * 2;   // one [128, 64] bf16 tile = 16 KB
constexpr int A2_BOX = 32 * 64 * 2;     // one TMA box: 32 rows x 64 columns = 4 KB
constexpr float kMaskFill2 = -10e6f;    // models/bert.py:161

template <bool kX3> struct A2Cfg {
    static constexpr int P = kX3 ? 2 : 1;
    static constexpr int kQK = 2 * P * A2_TILE;               // [Q hi, (Q lo), K hi, (K lo)]
    static constexpr int kStage = 3 * P * A2_TILE;            // + [V hi, (V lo)]: 96 KB (48 KB plain bf16)
    static constexpr int kStg = P * A2_TILE;                  // context staging rows, hi (+ lo)
    static constexpr int kTail = 1024;                        // barriers, tmem slot, geometry, pad-key words
    static constexpr int kSmemBytes = 2 * kStage + kStg + kTail + 1024 /*align*/;
};

// optional in-kernel timeline (navc_debug_trace_attn, tools/attn2_trace.py): roles 0 producer, 1 MMA issuer, 2 softmax group 0
__device__ unsigned long long* a2_trace_buf = nullptr;
struct A2Trace {
    unsigned long long* p;
    int n;
    __device__ A2Trace(int role) : p(nullptr), n(0) {
        if (a2_trace_buf) p = a2_trace_buf + ((size_t)blockIdx.x * 3 + role) * 64;
    }
    __device__ __forceinline__ void ev(unsigned tag) {
        if (p && n < 64) p[n++] = ((unsigned long long)tag << 56) | ((unsigned long long)clock64() & 0x00ffffffffffffffull);
    }
};

struct A2Geom { int row0, nq, krow0, nkeys, h, s0, s1, pad_; };

struct A2Params {
    int is_self, mask_kind, watch;
    int q_col, k_col, v_col;          // column offsets (elements) of head 0 inside the maps
    int n_tiles, tiles_per_owner;     // launch maximum of tiles; cross: tiles per video
    int E, group, n_seq, S, H, D;
    const int64_t* tokens;            // self: [n_seq, S]
    const int32_t* seq_off;           // [n_seq + 1] packed row offsets
    const int32_t* tile_seq;          // self: [n_tiles + 1] first sequence of every tile (navc_pack_tiles)
    uint16_t* ctx_hi; uint16_t* ctx_lo;
};

template <bool kX3>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn2_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                const __grid_constant__ CUtensorMap map_kv_hi, const __grid_constant__ CUtensorMap map_kv_lo,
                A2Params p) {
    using Cfg = A2Cfg<kX3>;
    constexpr int P = Cfg::P;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    auto stage_s = [&](int s) { return smem_base + (uint32_t)(s * Cfg::kStage); };
    auto stage_g = [&](int s) { return smem_gen + s * Cfg::kStage; };
    uint8_t* stg_g = smem_gen + 2 * Cfg::kStage;
    uint8_t* tail = stg_g + Cfg::kStg;
    const uint32_t bar_base = smem_base + 2 * Cfg::kStage + Cfg::kStg;
    // per stage s: qk, v, s, p, o, t ; then the four staging-quarter locks
    auto bar_qk = [&](int s) { return bar_base + 8u * (0 + s); };
    auto bar_v = [&](int s) { return bar_base + 8u * (2 + s); };
    auto bar_s = [&](int s) { return bar_base + 8u * (4 + s); };
    auto bar_p = [&](int s) { return bar_base + 8u * (6 + s); };
    auto bar_o = [&](int s) { return bar_base + 8u * (8 + s); };
    auto bar_t = [&](int s) { return bar_base + 8u * (10 + s); };
    auto bar_stg = [&](int q) { return bar_base + 8u * (12 + q); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
    A2Geom* geom = reinterpret_cast<A2Geom*>(tail + 192);             // [4]: item i uses slot i & 3
    uint32_t* padw_all = reinterpret_cast<uint32_t*>(tail + 320);     // [2 groups][4] PAD-key bits per 32-key chunk
    uint8_t* keypad_all = tail + 384;                                  // [2 groups][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();
    // (stale shared memory: Q / K rows that no box covers only reach rows / key columns of S that are not stored resp.
    // are masked by selects, never arithmetically; the V rows beyond an item's keys are cleared per item below)
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_qk(s), 1); mbar_init(bar_v(s), 1); mbar_init(bar_s(s), 1); mbar_init(bar_p(s), 4);
            mbar_init(bar_o(s), 1); mbar_init(bar_t(s), 4);
        }
        for (int q = 0; q < 4; ++q) mbar_init(bar_stg(q), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // tokens / Q / K / V may still be in flight in the preceding kernel

    if (warp == 0) {
        // ===================== producer =====================
        // The geometry of a work item hangs on two dependent global loads (~1 us each under load), and half of the
        // enumerated cross-attention tiles are empty: looked up one item at a time this paced the whole kernel (timeline:
        // 2.8 us between issues).  So the 32 lanes look up 32 consecutive items of this CTA at once, then lane 0 walks them.
        {
            int i = 0;
            A2Trace tr(0);
            if (lane != 0) tr.p = nullptr;
            tr.ev(1);
            const int total = p.n_tiles * p.H;
            bool done = false;
            for (int base = 0; !done; base += 32) {
                const int w = (int)blockIdx.x + (base + lane) * (int)gridDim.x;
                A2Geom mine;
                mine.row0 = mine.krow0 = mine.nkeys = mine.s0 = mine.s1 = mine.pad_ = 0;
                mine.nq = w < total ? 0 : -1;   // -1: past the end
                mine.h = 0;
                if (w < total) {
                    const int tile = w / p.H;   // head fastest: concurrent CTAs share a tile's rows in L2
                    mine.h = w - tile * p.H;
                    if (p.is_self) {
                        mine.s0 = __ldg(p.tile_seq + tile);
                        mine.s1 = __ldg(p.tile_seq + tile + 1);
                        if (mine.s1 > mine.s0) {
                            mine.row0 = __ldg(p.seq_off + mine.s0);
                            mine.nq = __ldg(p.seq_off + mine.s1) - mine.row0;
                            mine.krow0 = mine.row0;
                            mine.nkeys = mine.nq;
                        }
                    } else {
                        const int v = tile / p.tiles_per_owner, z = tile - v * p.tiles_per_owner;
                        const int r0 = __ldg(p.seq_off + v * p.group), r1 = __ldg(p.seq_off + min((v + 1) * p.group, p.n_seq));
                        mine.row0 = r0 + z * 128;
                        mine.nq = max(0, min(128, r1 - mine.row0));
                        mine.krow0 = v * p.E;
                        mine.nkeys = p.E;
                    }
                }
                for (int l = 0; l < 32; ++l) {
                    A2Geom g;
                    g.row0 = __shfl_sync(0xffffffffu, mine.row0, l); g.nq = __shfl_sync(0xffffffffu, mine.nq, l);
                    g.krow0 = __shfl_sync(0xffffffffu, mine.krow0, l); g.nkeys = __shfl_sync(0xffffffffu, mine.nkeys, l);
                    g.h = __shfl_sync(0xffffffffu, mine.h, l); g.s0 = __shfl_sync(0xffffffffu, mine.s0, l);
                    g.s1 = __shfl_sync(0xffffffffu, mine.s1, l); g.pad_ = 0;
                    if (g.nq < 0) { done = true; break; }
                    if (g.nq == 0) continue;
                    if (lane == 0) {
                        const int s = i & 1, k = i >> 1, h = g.h;
                        tr.ev(2);   // next item
                        // Q / K of stage s are free as soon as the S = Q K^T of item i-2 has retired (P lives in tensor memory)
                        if (i >= 2) mbar_wait_relaxed(bar_s(s), (uint32_t)((k - 1) & 1));
                        tr.ev(3);   // Q / K loads issued
                        geom[i & 3] = g;
                        const int nbq = (g.nq + 31) >> 5, nbk = (g.nkeys + 31) >> 5;
                        const uint32_t sq = stage_s(s), sk = sq + P * A2_TILE, sv = sq + Cfg::kQK;
                        mbar_expect_tx(bar_qk(s), (uint32_t)((nbq + nbk) * A2_BOX * P));
                        for (int b2 = 0; b2 < nbq; ++b2) {
                            tma_load_2d(sq + b2 * A2_BOX, &map_q_hi, bar_qk(s), p.q_col + h * 64, g.row0 + b2 * 32);
                            if (kX3) tma_load_2d(sq + A2_TILE + b2 * A2_BOX, &map_q_lo, bar_qk(s), p.q_col + h * 64, g.row0 + b2 * 32);
                        }
                        for (int b2 = 0; b2 < nbk; ++b2) {
                            tma_load_2d(sk + b2 * A2_BOX, &map_kv_hi, bar_qk(s), p.k_col + h * 64, g.krow0 + b2 * 32);
                            if (kX3) tma_load_2d(sk + A2_TILE + b2 * A2_BOX, &map_kv_lo, bar_qk(s), p.k_col + h * 64, g.krow0 + b2 * 32);
                        }
                        if (i >= 2) mbar_wait_relaxed(bar_o(s), (uint32_t)((k - 1) & 1));   // V of stage s: free once the P V of item i-2 has retired
                        mbar_expect_tx(bar_v(s), (uint32_t)(nbk * A2_BOX * P));
                        for (int b2 = 0; b2 < nbk; ++b2) {
                            tma_load_2d(sv + b2 * A2_BOX, &map_kv_hi, bar_v(s), p.v_col + h * 64, g.krow0 + b2 * 32);
                            if (kX3) tma_load_2d(sv + A2_TILE + b2 * A2_BOX, &map_kv_lo, bar_v(s), p.v_col + h * 64, g.krow0 + b2 * 32);
                        }
                    }
                    ++i;
                    __syncwarp();
                }
            }
            // one terminator per stage (each softmax group and the MMA warp stop on theirs)
            if (lane == 0) {
                for (int t = 0; t < 2; ++t, ++i) {
                    const int s = i & 1, k = i >> 1;
                    if (i >= 2) mbar_wait_relaxed(bar_o(s), (uint32_t)((k - 1) & 1));
                    geom[i & 3].nq = 0;
                    mbar_arrive(bar_qk(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(128, 128);
            constexpr uint32_t idesc_o = make_idesc_bmn(128, 64);
            auto do_pv = [&](int j) {
                const int s = j & 1;
                const uint32_t par = (uint32_t)((j >> 1) & 1);
                mbar_wait_relaxed(bar_v(s), par);
                mbar_wait_relaxed(bar_p(s), par);
                tc_fence_after();
                const uint32_t sv = stage_s(s) + Cfg::kQK;
                const uint32_t tO = tmem_base + (uint32_t)(s * 256), tP = tO + 128u;   // P hi at +128, P lo at +192 (64 columns each)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // A: P columns 8k..8k+7 (16 keys as packed bf16 pairs); B: V rows 16k..16k+15 (MN-major)
                    const uint64_t dv_hi = make_smem_desc_mn(sv + (uint32_t)(k * 2048)), dv_lo = make_smem_desc_mn(sv + A2_TILE + (uint32_t)(k * 2048));
                    if (kX3) {
                        tc_mma_bf16_ts(tO, tP + 64u + (uint32_t)(k * 8), dv_hi, idesc_o, k ? 1u : 0u);
                        tc_mma_bf16_ts(tO, tP + (uint32_t)(k * 8), dv_lo, idesc_o, 1u);
                        tc_mma_bf16_ts(tO, tP + (uint32_t)(k * 8), dv_hi, idesc_o, 1u);
                    } else {
                        tc_mma_bf16_ts(tO, tP + (uint32_t)(k * 8), dv_hi, idesc_o, k ? 1u : 0u);
                    }
                }
                tc_commit(bar_o(s));
            };
            int pending = -1, stops = 0;
            A2Trace tr(1);
            for (int i = 0; stops < 2; ++i) {
                const int s = i & 1, k = i >> 1;
                mbar_wait_relaxed(bar_qk(s), (uint32_t)(k & 1));
                tr.ev(4);   // Q / K landed
                if (geom[i & 3].nq == 0) {
                    if (pending >= 0) { do_pv(pending); pending = -1; }
                    mbar_arrive(bar_s(s));   // lets this stage's softmax group see its terminator
                    ++stops;
                    continue;
                }
                if (i >= 2) mbar_wait_relaxed(bar_t(s), (uint32_t)((k - 1) & 1));   // the epilogue of item i-2 has read its O: TMEM stage free
                tc_fence_after();
                const uint32_t sq = stage_s(s), sk = sq + P * A2_TILE;
                const uint32_t tS = tmem_base + (uint32_t)(s * 256);
                const uint64_t dq_hi = make_smem_desc(sq), dq_lo = make_smem_desc(sq + A2_TILE);
                const uint64_t dk_hi = make_smem_desc(sk), dk_lo = make_smem_desc(sk + A2_TILE);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t koff = (uint64_t)((kk * UMMA_K * 2) >> 4);
                    if (kX3) {
                        tc_mma_bf16(tS, dq_lo + koff, dk_hi + koff, idesc_s, kk ? 1u : 0u);
                        tc_mma_bf16(tS, dq_hi + koff, dk_lo + koff, idesc_s, 1u);
                        tc_mma_bf16(tS, dq_hi + koff, dk_hi + koff, idesc_s, 1u);
                    } else {
                        tc_mma_bf16(tS, dq_hi + koff, dk_hi + koff, idesc_s, kk ? 1u : 0u);
                    }
                }
                tc_commit(bar_s(s));
                tr.ev(5);   // S issued
                if (pending >= 0) do_pv(pending);   // O = P V of the previous item, while this item's softmax runs
                tr.ev(6);   // previous item's PV issued
                pending = i;
            }
        }
    } else {
        // ===================== softmax + epilogue groups (warps 2..5 and 6..9) =====================
        const int grp = (warp - 2) >> 2;            // group g serves the items of stage g
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;          // tile row = TMEM lane
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const int s = grp;
        uint8_t* gV = stage_g(s) + Cfg::kQK;
        const uint32_t tS = tmem_base + (uint32_t)(s * 256), tP = tS + 128u;
        uint32_t* padw = padw_all + grp * 4;
        uint8_t* keypad = keypad_all + grp * 128;
        const float scale = 0.125f;                 // 1/sqrt(dk), dk = 64: exact power of two == the reference's division
        constexpr float kLog2e = 1.4426950408889634f;
        auto range_mask = [](int lo, int hi, int c) -> uint32_t {  // bits of keys [lo, hi) inside chunk c
            const int a = max(lo - c * 32, 0), b = min(hi - c * 32, 32);
            if (b <= a) return 0u;
            const uint32_t upto_b = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
            return upto_b & ~((1u << a) - 1u);
        };
        A2Trace tr(2);
        if (warp != 2 || lane != 0) tr.p = nullptr;
        for (int k = 0;; ++k) {
            const uint32_t par = (uint32_t)(k & 1);
            const int item = 2 * k + grp;           // position in this CTA's item sequence
            mbar_wait(bar_s(s), par);
            const A2Geom g = geom[item & 3];
            if (g.nq == 0) break;
            tc_fence_after();
            tr.ev(7);   // S ready
            // ---- this row's visible keys [own_lo, own_hi) and, for self-attention, its position and the PAD keys ----
            int own_lo = 0, own_hi = g.nkeys, ipos = 0;
            if (p.is_self) {
                uint8_t kp = 0;
                if (r < g.nq) {
                    // the sequence of packed row row0 + r: the last i in [s0, s1) with seq_off[i] <= row
                    const int row = g.row0 + r;
                    int lo = g.s0, hi = g.s1 - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (__ldg(p.seq_off + mid) <= row) lo = mid; else hi = mid - 1;
                    }
                    const int b0 = __ldg(p.seq_off + lo), b1 = __ldg(p.seq_off + lo + 1);
                    own_lo = b0 - g.row0;
                    own_hi = b1 - g.row0;
                    ipos = row - b0;
                    if (p.tokens) kp = p.tokens[(size_t)lo * p.S + ipos] == NAVC_PAD ? 1 : 0;
                } else {
                    own_lo = own_hi = 0;   // rows beyond the tile see nothing (their P row is zero)
                }
                keypad[r] = kp;
                const uint32_t mine = __ballot_sync(0xffffffffu, kp != 0);
                if (lane == 0) padw[quarter] = mine;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the four warps of this group only
            }
            // chunks with a visible key for some row of this warp (warp uniform)
            int c_lo = own_hi > own_lo ? (own_lo >> 5) : 4, c_hi = own_hi > own_lo ? ((own_hi - 1) >> 5) : -1;
            c_lo = __reduce_min_sync(0xffffffffu, c_lo);
            c_hi = __reduce_max_sync(0xffffffffu, c_hi);
            const bool use_watch = (p.mask_kind == NAVC_MASK_CAUSAL) && p.watch != 0 && p.S >= p.watch;
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
            // ---- two passes over S, one 32-key chunk at a time.  The chunk loops are deliberately NOT unrolled: the first
            // version kept the whole row in registers with everything unrolled (4 chunks x 3 mask specialisations) and spent
            // 22 % of its issue slots waiting for instruction fetch (ncu: stall_no_inst) -- the code no longer fit the
            // instruction caches.  Re-reading a chunk from tensor memory is far cheaper than that. ----
            float m = -INFINITY;
#pragma unroll 1
            for (int c = c_lo; c <= c_hi; ++c) {
                uint32_t v[32];
                tc_ld32(tS + t_lane + (uint32_t)(c * 32), v);
                uint32_t vis, fill;
                chunk_masks(c, vis, fill);
                const uint32_t vm = vis & ~fill;
                const bool plain = __all_sync(0xffffffffu, vm == 0xffffffffu);
                tc_wait_ld();
                float mc = -INFINITY;
                if (plain) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mc = fmaxf(mc, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if ((vm >> j) & 1u) mc = fmaxf(mc, __uint_as_float(v[j]));
                }
                m = fmaxf(m, mc * scale);               // scale > 0: max commutes with the scaling
                if (fill) m = fmaxf(m, kMaskFill2);
            }
            tr.ev(8);   // row maxima known
            float sum = 0.f;
            const float m2 = m * kLog2e, sc2 = scale * kLog2e;
            const float e_fill = fast_exp2((kMaskFill2 - m) * kLog2e);  // 0 unless every visible key is filled
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t hi_w[16], lo_w[16];
                if (c < c_lo || c > c_hi) {  // warp-uniform: no visible key in this chunk for any row of the warp
#pragma unroll
                    for (int i = 0; i < 16; ++i) { hi_w[i] = 0u; lo_w[i] = 0u; }
                } else {
                    uint32_t v[32];
                    tc_ld32(tS + t_lane + (uint32_t)(c * 32), v);
                    uint32_t vm, fm;
                    chunk_masks(c, vm, fm);
                    // warp-uniform: every key of the chunk visible and unfilled for every row (cross-attention chunks below
                    // E, interior chunks of long sequences): the per-key selects are half of the issue slots of this loop
                    const bool plain = __all_sync(0xffffffffu, vm == 0xffffffffu && fm == 0u);
                    tc_wait_ld();
                    if (plain) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float e0 = fast_exp2(fmaf(__uint_as_float(v[j]), sc2, -m2));
                            const float e1 = fast_exp2(fmaf(__uint_as_float(v[j + 1]), sc2, -m2));
                            sum += e0 + e1;
                            if (kX3) {
                                split_bf16x2(e0, e1, hi_w[j >> 1], lo_w[j >> 1]);
                            } else {
                                const __nv_bfloat162 hb = __floats2bfloat162_rn(e0, e1);
                                hi_w[j >> 1] = *reinterpret_cast<const uint32_t*>(&hb);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float e2[2];
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                float e = fast_exp2(fmaf(__uint_as_float(v[j + q]), sc2, -m2));
                                if ((fm >> (j + q)) & 1u) e = e_fill;
                                if (!((vm >> (j + q)) & 1u)) e = 0.f;
                                e2[q] = e;
                                sum += e;
                            }
                            if (kX3) {
                                split_bf16x2(e2[0], e2[1], hi_w[j >> 1], lo_w[j >> 1]);
                            } else {
                                const __nv_bfloat162 hb = __floats2bfloat162_rn(e2[0], e2[1]);
                                hi_w[j >> 1] = *reinterpret_cast<const uint32_t*>(&hb);
                            }
                        }
                    }
                }
                // 32 keys = 16 packed columns of this row's P (hi at tP, lo at tP + 64)
                tc_st16(tP + t_lane + (uint32_t)(c * 16), hi_w);
                if (kX3) tc_st16(tP + 64u + t_lane + (uint32_t)(c * 16), lo_w);
            }
            tc_wait_st();
            // key rows beyond the item's keys that a box brought in anyway hold whatever follows in memory (rows past the
            // packed count are never written: possibly NaN bit patterns): P is 0 there, but 0 * NaN would poison O, so
            // those V rows are cleared (a key row = one 128-byte swizzled row of the hi / lo tile)
            if (r >= g.nkeys) {
                mbar_wait(bar_v(s), par);
                uint8_t* vrow = gV + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    *reinterpret_cast<uint4*>(vrow + i * 16) = make_uint4(0u, 0u, 0u, 0u);
                    if (kX3) *reinterpret_cast<uint4*>(vrow + A2_TILE + i * 16) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p(s));
            tr.ev(9);   // P written

            // ---- epilogue: O row * 1/sum -> staging rows -> full lines to global memory ----
            mbar_wait(bar_o(s), par);
            tc_fence_after();
            tr.ev(10);  // O ready
            const float inv = sum > 0.f ? 1.0f / sum : 0.f;
            uint32_t o[2][32];
            tc_ld32(tS + t_lane, o[0]);
            tc_ld32(tS + t_lane + 32u, o[1]);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t(s));            // TMEM stage s may take the S of item i+2
            // this warp's quarter of the staging rows is shared with the other group's warp of the same quarter
            tr.ev(12);  // O in registers
            if (item >= 1) mbar_wait(bar_stg(quarter), (uint32_t)((item - 1) & 1));
            tr.ev(13);  // staging quarter ours
            uint8_t* srow = stg_g + r * 128;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {   // 8 columns = one 16-byte chunk of the row; chunk j lives at j ^ (r & 7)
                const int cc = c8 >> 2, j0 = (c8 & 3) * 8;
                const float4 f0 = make_float4(__uint_as_float(o[cc][j0]) * inv, __uint_as_float(o[cc][j0 + 1]) * inv,
                                              __uint_as_float(o[cc][j0 + 2]) * inv, __uint_as_float(o[cc][j0 + 3]) * inv);
                const float4 f1 = make_float4(__uint_as_float(o[cc][j0 + 4]) * inv, __uint_as_float(o[cc][j0 + 5]) * inv,
                                              __uint_as_float(o[cc][j0 + 6]) * inv, __uint_as_float(o[cc][j0 + 7]) * inv);
                uint2 h0, l0, h1, l1;
                split_bf16x4(f0, h0, l0);
                split_bf16x4(f1, h1, l1);
                const uint32_t ch = (uint32_t)((c8 ^ (r & 7)) * 16);
                *reinterpret_cast<uint4*>(srow + ch) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                if (kX3) *reinterpret_cast<uint4*>(srow + A2_TILE + ch) = make_uint4(l0.x, l0.y, l1.x, l1.y);
            }
            __syncwarp();
            // 8 lanes per row, 4 rows per instruction: every store covers whole 128-byte lines of ctx[:, h*64 .. h*64+64)
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = quarter * 32 + it * 4 + (lane >> 3), cj = lane & 7;
                const uint8_t* src = stg_g + rr * 128 + ((cj ^ (rr & 7)) * 16);
                const uint4 hv = *reinterpret_cast<const uint4*>(src);
                uint4 lv = make_uint4(0u, 0u, 0u, 0u);
                if (kX3) lv = *reinterpret_cast<const uint4*>(src + A2_TILE);
                if (rr < g.nq) {
                    const size_t oo = (size_t)(g.row0 + rr) * p.D + g.h * 64 + cj * 8;
                    *reinterpret_cast<uint4*>(p.ctx_hi + oo) = hv;
                    if (kX3 && p.ctx_lo) *reinterpret_cast<uint4*>(p.ctx_lo + oo) = lv;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stg(quarter));    // staging quarter free for the next item
            tr.ev(11);  // context rows stored
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// first sequence of every 96-row window of the packed row space: tile_seq[t] = first n with seq_off[n] >= 96 t
// (t = 0 .. n_tiles; sequences are at most 32 rows, so a window's sequences span at most 127 rows)
__global__ void pack_tiles_kernel(const int32_t* __restrict__ seq_off, int N, int window, int32_t* __restrict__ tile_seq, int n_tiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const int target = t * window;
    int lo = 0, hi = N;   // first n in [0, N] with seq_off[n] >= target (seq_off[N] = row count; N if none)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (seq_off[mid] >= target) hi = mid; else lo = mid + 1;
    }
    tile_seq[t] = lo;
}

static bool g_a2_ready = false;
static int a2_init() {
    if (g_a2_ready) return 0;
    NAVC_CUDA(cudaFuncSetAttribute(attn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A2Cfg<false>::kSmemBytes));
    NAVC_CUDA(cudaFuncSetAttribute(attn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A2Cfg<true>::kSmemBytes));
    g_a2_ready = true;
    return 0;
}

static int launch_attn2(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq, int q_cols, int q_rows,
                        const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv, int kv_cols, int kv_rows,
                        const A2Params& p, cudaStream_t st, const char* what) {
    NAVC_REQUIRE(tc_ready(), "%s: navc_init() has not been called", what);
    NAVC_REQUIRE(mode == NAVC_TC_BF16 || mode == NAVC_TC_BF16X3, "%s: bad mode %d", what, mode);
    NAVC_REQUIRE(q_hi && kv_hi && (mode == NAVC_TC_BF16 || (q_lo && kv_lo)), "%s: null operand", what);
    NAVC_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && p.D % 8 == 0, "%s: leading dimensions must be multiples of 8", what);
    NAVC_REQUIRE((((uintptr_t)q_hi | (uintptr_t)q_lo | (uintptr_t)kv_hi | (uintptr_t)kv_lo | (uintptr_t)p.ctx_hi | (uintptr_t)p.ctx_lo) & 15) == 0,
                 "%s: operands must be 16-byte aligned", what);
    if (a2_init()) return 2;
    CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo;
    if (tc_make_map(&mq_hi, q_hi, q_rows, q_cols, ldq, 32) || tc_make_map(&mk_hi, kv_hi, kv_rows, kv_cols, ldkv, 32)) return 1;
    mq_lo = mq_hi;
    mk_lo = mk_hi;
    if (mode == NAVC_TC_BF16X3) {
        if (tc_make_map(&mq_lo, q_lo, q_rows, q_cols, ldq, 32) || tc_make_map(&mk_lo, kv_lo, kv_rows, kv_cols, ldkv, 32)) return 1;
    }
    int sms = navc_sm_count();
    if (sms <= 0) sms = 148;
    const int items = p.n_tiles * p.H;
    dim3 grid(items < sms ? items : sms);
    if (mode == NAVC_TC_BF16X3)
        NAVC_CUDA(launch_pdl(attn2_tc_kernel<true>, grid, dim3(A2_THREADS), A2Cfg<true>::kSmemBytes, st, mq_hi, mq_lo, mk_hi, mk_lo, p));
    else
        NAVC_CUDA(launch_pdl(attn2_tc_kernel<false>, grid, dim3(A2_THREADS), A2Cfg<false>::kSmemBytes, st, mq_hi, mq_lo, mk_hi, mk_lo, p));
    return check_launch(what);
}

}  // namespace navc

using namespace navc;

extern "C" int navc_attention_window(void) { return 96; }

// Debug: install (or remove, buf == NULL) the in-kernel timeline buffer: 3 x 64 uint64 per CTA of the grid.
extern "C" int navc_debug_trace_attn(void* buf) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(navc::a2_trace_buf, &q, sizeof(q)) == cudaSuccess ? 0 : 1;
}

extern "C" int navc_pack_tiles(const int32_t* seq_off, int N, int32_t* tile_seq, int n_tiles, void* stream) {
    NAVC_REQUIRE(seq_off && tile_seq && N > 0 && n_tiles > 0, "navc_pack_tiles: bad arguments");
    pack_tiles_kernel<<<(n_tiles + 1 + 127) / 128, 128, 0, as_stream(stream)>>>(seq_off, N, 96, tile_seq, n_tiles);
    return check_launch("navc_pack_tiles");
}

extern "C" int navc_self_attention_tc_tiles(int mode, const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld,
                                            const int64_t* tokens, const int32_t* seq_off, const int32_t* tile_seq,
                                            int n_tiles, int rows, int N, int S, int D, int H, int mask_kind, int watch,
                                            uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream) {
    NAVC_REQUIRE(tokens && seq_off && tile_seq && ctx_hi && rows > 0 && n_tiles > 0, "navc_self_attention_tc_tiles: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && S <= 32 && H > 0 && D == H * 64 && ld >= 3 * D,
                 "navc_self_attention_tc_tiles: needs dk == 64 and S <= 32 (N=%d S=%d D=%d H=%d)", N, S, D, H);
    NAVC_REQUIRE(mask_kind >= 0 && mask_kind <= 2, "navc_self_attention_tc_tiles: bad mask kind");
    A2Params p = {};
    p.is_self = 1; p.mask_kind = mask_kind; p.watch = watch;
    p.q_col = 0; p.k_col = D; p.v_col = 2 * D;
    p.n_tiles = n_tiles; p.tiles_per_owner = 1; p.n_seq = N; p.S = S; p.H = H; p.D = D; p.group = 1;
    p.tokens = tokens; p.seq_off = seq_off; p.tile_seq = tile_seq; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo;
    return launch_attn2(mode, qkv_hi, qkv_lo, ld, 3 * D, rows, qkv_hi, qkv_lo, ld, 3 * D, rows, p, as_stream(stream),
                        "navc_self_attention_tc_tiles");
}

extern "C" int navc_cross_attention_tc_tiles(int mode, const uint16_t* q_hi, const uint16_t* q_lo, int ldq,
                                             const uint16_t* kv_hi, const uint16_t* kv_lo, int ldkv,
                                             const int32_t* seq_off, int rows, int N, int S, int E, int D, int H, int group,
                                             uint16_t* ctx_hi, uint16_t* ctx_lo, void* stream) {
    NAVC_REQUIRE(seq_off && ctx_hi && rows > 0, "navc_cross_attention_tc_tiles: null pointer");
    NAVC_REQUIRE(N > 0 && S > 0 && E > 0 && E <= 128 && H > 0 && D == H * 64 && group >= 1 && N % group == 0 &&
                     ldq >= D && ldkv >= 2 * D,
                 "navc_cross_attention_tc_tiles: needs dk == 64 and E <= 128 (N=%d S=%d E=%d D=%d H=%d)", N, S, E, D, H);
    const int G = N / group, nq = group * S, tpo = (nq + 127) / 128;
    A2Params p = {};
    p.is_self = 0;
    p.q_col = 0; p.k_col = 0; p.v_col = D;
    p.n_tiles = G * tpo; p.tiles_per_owner = tpo; p.E = E; p.group = group; p.n_seq = N; p.S = S; p.H = H; p.D = D;
    p.seq_off = seq_off; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo;
    return launch_attn2(mode, q_hi, q_lo, ldq, D, rows, kv_hi, kv_lo, ldkv, 2 * D, G * E, p, as_stream(stream),
                        "navc_cross_attention_tc_tiles");
}
