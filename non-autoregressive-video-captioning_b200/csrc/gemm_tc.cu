// tcgen05 GEMM for sm_100a:  Y = epilogue(X[M,K] * W[N,K]^T)
//
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (UMMA 128x256x16,
//   kind::f16, bf16 operands, fp32 accumulators in TMEM, double buffered) -> tcgen05.ld -> fused
//   epilogue (bias / activation / residual / non-pad row mask / bf16 hi-lo split, or on-the-fly
//   softmax statistics for the vocabulary projection).
//
// Warp roles (320 threads, persistent, one CTA per SM):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  epilogue: warp w reads TMEM lane quarter w%4 (row = TMEM lane) and column half
//               (w-2)/4 of the 256-column accumulator; 32x32 chunks are transposed through a
//               swizzled shared-memory staging buffer so that every global load/store of the
//               epilogue (residual, fp32 / bf16 hi / bf16 lo outputs) is a coalesced row segment.
//
// Operand modes: NAVC_TC_BF16 issues one product per k-step (hi*hi); NAVC_TC_BF16X3 issues three
// (hi*hi + hi*lo + lo*hi) which recovers ~fp32 accuracy from bf16 tensor cores.
#include <stdlib.h>

#include "tc_common.cuh"

namespace navc {

constexpr int TBM = 128, TBK = 64;   // CTA tile rows; TBK bf16 = one 128-byte swizzle row
constexpr int kVocabBN = 256;        // the vocabulary epilogue always uses 256-wide tiles
// Tile width TBN is a template parameter: 256 by default, 128 when the 256-wide tiling would leave
// a large part of the last wave of the persistent grid empty (e.g. N = 512: 336 tiles on 148 SMs).
constexpr int kEpiWarps = 8;                         // generic / vocabulary epilogues
constexpr int kPairEpiWarps = 16;                    // pair epilogue: 4 warps per scheduler to hide its latencies
constexpr int kStageFloats = 32 * 32;                // per-epilogue-warp transpose buffer (4 KB); 2 KB per warp in pair mode
__host__ __device__ constexpr int tc_epi_warps(int epi) { return (epi == 2 || epi == 3) ? kPairEpiWarps : kEpiWarps; }
constexpr int kResBufBytes = 2048;                   // per warp and buffer: residual hi box (1 KB) + lo box (1 KB)
__host__ __device__ constexpr int tc_threads(int epi) { return 64 + 32 * tc_epi_warps(epi); }
constexpr int kTileABytes = TBM * TBK * 2;           // 16 KB
constexpr int kAccStages = 2;                        // 2 x 256 TMEM columns
constexpr int kTmemCols = 512;

template <bool kX3, int TBN, int kEpi = 0> struct TcCfg {
    static constexpr int kTileBBytes = TBN * TBK * 2;    // 32 KB (TBN = 256) / 16 KB (TBN = 128)
    static constexpr int kStageBytes = (kX3 ? 2 : 1) * (kTileABytes + kTileBBytes);
    // 192 KB operand ring; the residual-prefetching epilogue (kEpi == 3) trades 64 KB of it for its
    // per-warp double-buffered residual boxes
    static constexpr int kResBytes = kEpi == 3 ? kPairEpiWarps * 2 * kResBufBytes : 0;
    static constexpr int kStages = (196608 - kResBytes) / kStageBytes;
    static constexpr int kRingBytes = kStages * kStageBytes;
    static constexpr int kSmemBytes = kRingBytes + kEpiWarps * kStageFloats * 4 + kResBytes + 1024 /*align slack*/ + 512 /*barriers*/;
    static_assert(kStages >= 2, "operand ring needs at least two stages");
};

struct TcVocab {
    const float* bias;
    float* part_max;
    float* part_sum;
    int32_t* part_idx;
    const int64_t* target;
    float* target_logit;
    int n_tiles;
};

// kEpi: 0 = generic epilogue (fp32 / bf16 outputs, fp32 residual; smem-transposed coalesced stores),
//       1 = vocabulary softmax statistics, 2 = "pair" epilogue (lane = row, bf16 hi/lo outputs through
//       TMA stores, bf16 hi+lo residual): the inference fast path.
constexpr int kEpiGeneric = 0, kEpiVocab = 1, kEpiPair = 2, kEpiPairRes = 3;  // 3 = pair + TMA-prefetched residual
// 4 = generic epilogue over MN-major operands: Y[n, k] = sum_m A[m, n] * B[m, k] with A [rows, N_out] and
// B [rows, K_in] both row-major (the weight gradient dW = dY^T X straight from dY and X: no transposed copies).
// A k-block is 64 reduction rows; each operand tile is loaded as 64-column TMA boxes of 64 rows (8 KB, 128B
// swizzle), i.e. the canonical MN-major layout with 8 KB between 64-wide MN blocks and 1 KB between 8-row groups.
constexpr int kEpiWgrad = 4;
// 5 = generic epilogue, A K-major as usual but B MN-major: Y[m, k] = sum_n A[m, n] * B[n, k] with B [N, K_out]
// row-major (the data gradient dX = dY W straight from the forward's weight copies: no transposed weights).
constexpr int kEpiDgrad = 5;

// kTf32: the operands are fp32 matrices consumed as TF32 (kind::tf32; generic epilogue only): a k-block is 32 elements =
// the same 128-byte swizzle rows and tile sizes as 64 bf16, four k-steps of 8 elements = the same 32-byte descriptor steps.
template <bool kX3, int kEpi, int TBN, bool kTf32 = false>
__global__ void __launch_bounds__(tc_threads(kEpi), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
               const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
               int M_max, int N, int K, EpiParams epi, TcVocab vep) {
    // device-side row count (packed-row decoding): only the first *m_dev rows are computed
    const int M = epi.m_dev ? min(M_max, __ldg(epi.m_dev)) : M_max;
    constexpr bool kVocab = kEpi == kEpiVocab;
    constexpr bool kPairAny = kEpi == kEpiPair || kEpi == kEpiPairRes;
    constexpr bool kMNA = kEpi == kEpiWgrad;                        // A tile MN-major (reduction rows x 64-wide M boxes)
    constexpr bool kMNB = kEpi == kEpiWgrad || kEpi == kEpiDgrad;   // B tile MN-major
    constexpr bool kResTma = kEpi == kEpiPairRes;
    using Cfg = TcCfg<kX3, TBN, kEpi>;
    constexpr int kTileBBytes = Cfg::kTileBBytes;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // barriers live after the tile ring
    float* stage_all = reinterpret_cast<float*>(smem_gen + Cfg::kRingBytes);
    const uint32_t res_off = Cfg::kRingBytes + kEpiWarps * kStageFloats * 4;   // residual boxes (kEpi == 3)
    const uint32_t bar_off = res_off + Cfg::kResBytes;
    const uint32_t bar_base = smem_base + bar_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + kAccStages + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + bar_off + 8 * (2 * Cfg::kStages + 2 * kAccStages));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + TBM - 1) / TBM, n_blocks = (N + TBN - 1) / TBN;
    constexpr int kKB = kTf32 ? TBK / 2 : TBK;   // K elements per k-block
    const int k_blocks = (K + kKB - 1) / kKB;  // TMA zero-fills the K tail of the last block
    const int split = kVocab ? 1 : epi.split_k;
    const int kpb = (k_blocks + split - 1) / split;  // k-blocks per split (host guarantees every split is non-empty)
    const int n_tiles = m_blocks * n_blocks * split;
    // Tail split ("stream-K lite", pair epilogue): the tiles of the last, partly filled wave of the persistent grid are
    // cut into sk_parts K-ranges so that every SM works during that wave; part 0 of a tile adds the other parts'
    // partial accumulators (handed over through sk_ws) in its epilogue.  Derived in-kernel because M may be device-side.
    int sk_full = n_tiles, sk_tail = 0, sk_parts = 1, sk_per = k_blocks;
    if (kEpi == kEpiPair && epi.sk_ws != nullptr && split == 1) {
        const int G = (int)gridDim.x;
        sk_full = (n_tiles / G) * G;
        sk_tail = n_tiles - sk_full;
        // a part must keep >= 8 k-blocks (the hand-over costs about as much as 4-6 k-blocks of MMA) and the finishing
        // CTA sums at most 3 partner tiles
        int p = sk_tail > 0 ? min(min(G / sk_tail, k_blocks / 8), 4) : 1;
        if (p >= 2) {
            sk_per = (k_blocks + p - 1) / p;
            sk_parts = (k_blocks + sk_per - 1) / sk_per;
        }
        if (sk_parts < 2) { sk_full = n_tiles; sk_tail = 0; sk_parts = 1; sk_per = k_blocks; }
    }
    struct Unit { int mn, kb0, kb1, part, tl; };
    auto get_unit = [&](int it, Unit& u) -> bool {
        const int t = (int)blockIdx.x + it * (int)gridDim.x;
        if (sk_parts > 1) {
            if (t < sk_full) { u.mn = t; u.kb0 = 0; u.kb1 = k_blocks; u.part = -1; u.tl = 0; return true; }
            const int x = t - sk_full;
            if (x >= sk_tail * sk_parts) return false;
            u.tl = x / sk_parts;
            u.part = x - u.tl * sk_parts;
            u.mn = sk_full + u.tl;
            u.kb0 = u.part * sk_per;
            u.kb1 = min(k_blocks, u.kb0 + sk_per);
            return true;
        }
        if (t >= n_tiles) return false;
        u.mn = t / split;
        u.kb0 = (t % split) * kpb;
        u.kb1 = min(k_blocks, u.kb0 + kpb);
        u.part = -1;
        u.tl = 0;
        return true;
    };

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), tc_epi_warps(kEpi)); }
        if constexpr (kResTma) {
            for (int s = 0; s < 2 * kPairEpiWarps; ++s) mbar_init(bar_base + 256u + 8u * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // everything below reads / writes memory the preceding kernel may still be producing

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            Unit u;
            for (int it = 0; get_unit(it, u); ++it) {
                const int mn = u.mn;
                const int mb = mn / n_blocks, nb = mn % n_blocks;
                const int kb0 = u.kb0, kb1 = u.kb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                    const int row = kb * TBK;  // MN-major tiles: reduction rows [row, row + 64); rows beyond the tensor arrive as zeros
                    if constexpr (kMNA) {
#pragma unroll
                        for (int j = 0; j < TBM / 64; ++j) {
                            tma_load_2d(sa + j * 8192, &map_a_hi, full_bar(stage), mb * TBM + j * 64, row);
                            if (kX3) tma_load_2d(sa + kTileABytes + kTileBBytes + j * 8192, &map_a_lo, full_bar(stage), mb * TBM + j * 64, row);
                        }
                    } else {
                        tma_load_2d(sa, &map_a_hi, full_bar(stage), kb * kKB, mb * TBM);
                        if (kX3) tma_load_2d(sa + kTileABytes + kTileBBytes, &map_a_lo, full_bar(stage), kb * TBK, mb * TBM);
                    }
                    if constexpr (kMNB) {
#pragma unroll
                        for (int j = 0; j < TBN / 64; ++j) {
                            tma_load_2d(sa + kTileABytes + j * 8192, &map_b_hi, full_bar(stage), nb * TBN + j * 64, row);
                            if (kX3) tma_load_2d(sa + 2 * kTileABytes + kTileBBytes + j * 8192, &map_b_lo, full_bar(stage), nb * TBN + j * 64, row);
                        }
                    } else {
                        tma_load_2d(sa + kTileABytes, &map_b_hi, full_bar(stage), kb * kKB, nb * TBN);
                        if (kX3) tma_load_2d(sa + 2 * kTileABytes + kTileBBytes, &map_b_lo, full_bar(stage), kb * TBK, nb * TBN);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = kTf32 ? make_idesc_tf32(TBM, TBN)
                                             : (make_idesc(TBM, TBN) | (kMNA ? (1u << 15) : 0u) | (kMNB ? (1u << 16) : 0u));  // A / B MN-major bits
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            Unit u;
            for (int it = 0; get_unit(it, u); ++it) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TBN);
                const int kb0 = u.kb0, kb1 = u.kb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    const uint64_t da_hi = kMNA ? make_smem_desc_mn(sa) : make_smem_desc(sa);
                    const uint64_t db_hi = kMNB ? make_smem_desc_mn(sa + kTileABytes) : make_smem_desc(sa + kTileABytes);
                    const uint64_t da_lo = kMNA ? make_smem_desc_mn(sa + kTileABytes + kTileBBytes) : make_smem_desc(sa + kTileABytes + kTileBBytes);
                    const uint64_t db_lo = kMNB ? make_smem_desc_mn(sa + 2 * kTileABytes + kTileBBytes) : make_smem_desc(sa + 2 * kTileABytes + kTileBBytes);
#pragma unroll
                    for (int k = 0; k < TBK / UMMA_K; ++k) {
                        // K-major: 32 bytes per k-step inside the 128-byte rows; MN-major: 16 rows of 128 bytes
                        const uint64_t koff_a = kMNA ? (uint64_t)((k * UMMA_K * 128) >> 4) : (uint64_t)((k * UMMA_K * 2) >> 4);
                        const uint64_t koff_b = kMNB ? (uint64_t)((k * UMMA_K * 128) >> 4) : (uint64_t)((k * UMMA_K * 2) >> 4);
                        if (kX3) {
                            // small cross terms first, the dominant hi*hi product last
                            tc_mma_bf16(d_tmem, da_lo + koff_a, db_hi + koff_b, idesc, ((kb - kb0) | k) ? 1u : 0u);
                            tc_mma_bf16(d_tmem, da_hi + koff_a, db_lo + koff_b, idesc, 1u);
                            tc_mma_bf16(d_tmem, da_hi + koff_a, db_hi + koff_b, idesc, 1u);
                        } else if (kTf32) {
                            tc_mma_tf32(d_tmem, da_hi + koff_a, db_hi + koff_b, idesc, ((kb - kb0) | k) ? 1u : 0u);
                        } else {
                            tc_mma_bf16(d_tmem, da_hi + koff_a, db_hi + koff_b, idesc, ((kb - kb0) | k) ? 1u : 0u);
                        }
                    }
                    tc_commit(empty_bar(stage));                       // frees the smem slot when the MMAs retire
                    if (kb == kb1 - 1) tc_commit(tfull_bar(acc));  // accumulator ready for the epilogue
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int half = ew >> 2;      // column half [half*128, half*128+128) of the accumulator
        float* stage = stage_all + ew * kStageFloats;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t rstep = 0;  // residual boxes consumed so far by this warp (kEpi == 3)
        (void)rstep;
        Unit u;
        for (int it = 0; get_unit(it, u); ++it) {
            const int mn = u.mn;
            const int mb = mn / n_blocks, nb = mn % n_blocks;
            const int row0 = mb * TBM + quarter * 32;
            const int row = row0 + lane;
            const int n0 = nb * TBN + half * (TBN / 2);

            if constexpr (kPairAny) {
                // ---- pair epilogue: 16 warps; warp = (TMEM lane quarter, column group of TBN/4); lane = row;
                //      16-column steps; outputs leave through TMA stores of [32 rows x 16 cols] bf16 boxes ----
                constexpr int GW = TBN / 4;
                const int grp = ew >> 2;
                const int rowp = row0 + lane;
                const bool row_ok = rowp < M;
                const bool rz = (row_ok && epi.row_tokens) ? (epi.row_tokens[rowp] == NAVC_PAD) : false;
                uint8_t* stg = reinterpret_cast<uint8_t*>(stage_all) + ew * 2048;   // hi box at +0, lo box at +1024
                const uint32_t stg_s = smem_u32(stg);
                const int ng0 = nb * TBN + grp * GW;
                // residual boxes of step `rstep` (per warp, double buffered) arrive through TMA: no LSU traffic
                const uint32_t rbuf_s = smem_base + res_off + (uint32_t)(ew * 2 * kResBufBytes);
                const uint8_t* rbuf_g = smem_gen + res_off + ew * 2 * kResBufBytes;
                auto res_issue = [&](uint32_t step, int col) {
                    if (lane == 0) {
                        const uint32_t b = step & 1u, bar = bar_base + 256u + 8u * (uint32_t)(ew * 2 + (int)b);
                        mbar_expect_tx(bar, epi.res_lo ? 2048u : 1024u);
                        tma_load_2d(rbuf_s + b * kResBufBytes, &map_r_hi, bar, col, row0);
                        if (epi.res_lo) tma_load_2d(rbuf_s + b * kResBufBytes + 1024u, &map_r_lo, bar, col, row0);
                    }
                };
                if constexpr (kResTma) {
                    if (ng0 < N) res_issue(rstep, ng0);
                }
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TBN + grp * GW);
                // tail split: this thread's row of the partial tiles, [slot][128 rows][TBN] fp32
                float* sk_row = nullptr;
                if (u.part >= 0)
                    sk_row = epi.sk_ws + ((size_t)(u.tl * (sk_parts - 1) + (u.part > 0 ? u.part - 1 : 0)) * TBM + quarter * 32 + lane) * TBN + grp * GW;
                if (u.part > 0) {
                    // a partner part: park the raw accumulators, signal, done
#pragma unroll 1
                    for (int c = 0; c < GW / 16; ++c) {
                        if (ng0 + c * 16 >= N) break;
                        uint32_t r[16];
                        tc_ld16(t_row + (uint32_t)(c * 16), r);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            __stcg(reinterpret_cast<uint4*>(sk_row + c * 16 + i * 4), make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]));
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicAdd(epi.sk_cnt + u.tl, 1);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                    if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
                    continue;
                }
                if (u.part == 0) {
                    // the finishing part: all warps of all partners must have parked their partials.  Every unit of the
                    // tail wave owns a CTA of the (fully resident) grid, so the partners are running; the wait is bounded
                    // anyway so that a bug shows up as a flagged wrong result, not as a hung GPU.
                    if (lane == 0) {
                        const int want = (sk_parts - 1) * kPairEpiWarps;
                        int* cnt = epi.sk_cnt + u.tl;
                        unsigned spins = 0;
                        while (atomicAdd(cnt, 0) < want) {   // read through L2, where the partners' increments land
                            __nanosleep(64);
                            if (++spins > (1u << 24)) { *epi.sk_err = 1; break; }
                        }
                    }
                    __syncwarp();
                    __threadfence();
                }
#pragma unroll 1
                for (int c = 0; c < GW / 16; ++c) {
                    const int col0 = ng0 + c * 16;
                    if (col0 >= N) break;  // warp-uniform
                    uint32_t r[16];
                    tc_ld16(t_row + (uint32_t)(c * 16), r);
                    if constexpr (kResTma) {
                        if (c + 1 < GW / 16 && col0 + 16 < N) res_issue(rstep + 1u, col0 + 16);  // prefetch the next step
                    }
                    // operands of this step (requested before the TMEM wait)
                    float4 bv[4];
                    uint4 rh[2], rl[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (epi.bias && col0 + i * 4 < N) bv[i] = __ldg(reinterpret_cast<const float4*>(epi.bias + col0 + i * 4));
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        rh[i] = make_uint4(0u, 0u, 0u, 0u);
                        rl[i] = make_uint4(0u, 0u, 0u, 0u);
                        if constexpr (!kResTma) {
                            if (epi.res_hi && row_ok && col0 + i * 8 < N) {
                                const size_t ro = (size_t)rowp * epi.ld_res + col0 + i * 8;
                                rh[i] = __ldg(reinterpret_cast<const uint4*>(epi.res_hi + ro));
                                if (epi.res_lo) rl[i] = __ldg(reinterpret_cast<const uint4*>(epi.res_lo + ro));
                            }
                        }
                    }
                    if constexpr (kResTma) {
                        const uint32_t b = rstep & 1u;
                        mbar_wait(bar_base + 256u + 8u * (uint32_t)(ew * 2 + (int)b), (rstep >> 1) & 1u);
                        const uint8_t* rb = rbuf_g + b * kResBufBytes + lane * 32;   // rows/cols beyond M/N arrive as zeros
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            rh[i] = *reinterpret_cast<const uint4*>(rb + i * 16);
                            if (epi.res_lo) rl[i] = *reinterpret_cast<const uint4*>(rb + 1024 + i * 16);
                        }
                        ++rstep;
                    }
                    tc_wait_ld();
                    if (u.part == 0) {
                        // add the partners' partial accumulators (parked in L2): all loads in flight, then the adds
                        float4 w[3][4];
#pragma unroll
                        for (int p = 0; p < 3; ++p)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                w[p][i] = (p + 1 < sk_parts) ? __ldcg(reinterpret_cast<const float4*>(sk_row + (size_t)p * TBM * TBN + c * 16 + i * 4))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int p = 0; p < 3; ++p)
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                r[i * 4 + 0] = __float_as_uint(__uint_as_float(r[i * 4 + 0]) + w[p][i].x);
                                r[i * 4 + 1] = __float_as_uint(__uint_as_float(r[i * 4 + 1]) + w[p][i].y);
                                r[i * 4 + 2] = __float_as_uint(__uint_as_float(r[i * 4 + 2]) + w[p][i].z);
                                r[i * 4 + 3] = __float_as_uint(__uint_as_float(r[i * 4 + 3]) + w[p][i].w);
                            }
                    }
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i * 4 + 0] = __uint_as_float(r[i * 4 + 0]) + bv[i].x;
                        v[i * 4 + 1] = __uint_as_float(r[i * 4 + 1]) + bv[i].y;
                        v[i * 4 + 2] = __uint_as_float(r[i * 4 + 2]) + bv[i].z;
                        v[i * 4 + 3] = __uint_as_float(r[i * 4 + 3]) + bv[i].w;
                    }
                    if (epi.act == NAVC_ACT_GELU_NEW) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = act_apply_fast(v[j], NAVC_ACT_GELU_NEW);
                    } else if (epi.act != NAVC_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = act_apply_fast(v[j], epi.act);
                    }
                    if (kResTma || epi.res_hi) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint32_t hw_[4] = {rh[i].x, rh[i].y, rh[i].z, rh[i].w};
                            const uint32_t lw_[4] = {rl[i].x, rl[i].y, rl[i].z, rl[i].w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                v[i * 8 + u * 2 + 0] += __uint_as_float(hw_[u] << 16) + __uint_as_float(lw_[u] << 16);
                                v[i * 8 + u * 2 + 1] += __uint_as_float(hw_[u] & 0xffff0000u) + __uint_as_float(lw_[u] & 0xffff0000u);
                            }
                        }
                    }
                    if (rz) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = 0.f;
                    }
                    uint32_t hw[8], lw[8];
                    if (epi.out_lo) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            hw[j] = *reinterpret_cast<const uint32_t*>(&hb);
                        }
                    }
                    // the previous step's TMA stores must have finished reading the staging boxes
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                    *reinterpret_cast<uint4*>(stg + lane * 32) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(stg + lane * 32 + 16) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                    if (epi.out_lo) {
                        *reinterpret_cast<uint4*>(stg + 1024 + lane * 32) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        *reinterpret_cast<uint4*>(stg + 1024 + lane * 32 + 16) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_o_hi, stg_s, col0, row0);
                        if (epi.out_lo) tma_store_2d(&map_o_lo, stg_s + 1024u, col0, row0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (u.part == 0 && lane == 0) {
                    // the last finishing warp to leave re-arms the counters for the next launch
                    if (atomicAdd(epi.sk_cnt + gridDim.x + u.tl, 1) == kPairEpiWarps - 1) {
                        epi.sk_cnt[u.tl] = 0;
                        epi.sk_cnt[gridDim.x + u.tl] = 0;
                        __threadfence();
                    }
                }
            } else if constexpr (!kVocab) {
                const bool row_ok = row < M;
                const bool my_rz = (row_ok && epi.row_tokens) ? (epi.row_tokens[row] == NAVC_PAD) : false;
                const uint32_t rz_mask = __ballot_sync(0xffffffffu, my_rz);
                const bool vec = (epi.ld_out % 4 == 0) && (N % 4 == 0) && (!epi.residual || epi.ld_res % 4 == 0) &&
                                 ((((uintptr_t)epi.bias) & 15) == 0) &&
                                 (epi.act == NAVC_ACT_NONE || epi.act == NAVC_ACT_GELU_NEW);  // other activations: scalar path
                const int cc = lane & 7, rr = lane >> 3;  // phase-2 mapping: 8 lanes per row, 4 rows per step
                // number of 32-column chunks of this warp's half that exist (warp-uniform)
                int n_chunks = (N - n0 + 31) / 32;
                n_chunks = n_chunks < 0 ? 0 : (n_chunks > TBN / 64 ? TBN / 64 : n_chunks);
                if (epi.dbg == 1) n_chunks = 0;
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TBN + half * (TBN / 2));
                uint32_t r[32];
                if (n_chunks > 0) tc_ld32(t_row, r);
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c) {
                    const int col0 = n0 + c * 32;
                    const int col2 = col0 + cc * 4;
                    // global operands of this chunk are requested before the TMEM wait so their latency overlaps
                    float4 q[8], bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (vec) {
                        if (epi.bias && col2 < N) bv = __ldg(reinterpret_cast<const float4*>(epi.bias + col2));
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int grow = row0 + it * 4 + rr;
                            q[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (epi.residual && grow < M && col2 < N)
                                q[it] = __ldg(reinterpret_cast<const float4*>(epi.residual + (size_t)grow * epi.ld_res + col2));
                        }
                    }
                    tc_wait_ld();
                    if (vec) {
                        // phase 1 (lane = row): park the raw 32x32 accumulator chunk in shared memory.
                        // float4 slot g of row `lane` lives at slot g ^ (lane & 7): conflict-free both ways.
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            *reinterpret_cast<uint4*>(stage + lane * 32 + ((g ^ (lane & 7)) << 2)) =
                                make_uint4(r[g * 4 + 0], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
                        __syncwarp();
                        if (c + 1 < n_chunks) tc_ld32(t_row + (uint32_t)((c + 1) * 32), r);  // overlaps phase 2
                        // phase 2: residual, row mask, coalesced stores.  Branch-free so the eight row
                        // groups overlap: all shared loads first, then the math, then predicated stores.
                        if (epi.dbg == 2) continue;
                        float4 v[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rl = it * 4 + rr;
                            v[it] = *reinterpret_cast<const float4*>(stage + rl * 32 + ((cc ^ (rl & 7)) << 2));
                        }
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            v[it].x += bv.x; v[it].y += bv.y; v[it].z += bv.z; v[it].w += bv.w;
                        }
                        if (epi.act == NAVC_ACT_GELU_NEW) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                v[it].x = act_apply_fast(v[it].x, NAVC_ACT_GELU_NEW); v[it].y = act_apply_fast(v[it].y, NAVC_ACT_GELU_NEW);
                                v[it].z = act_apply_fast(v[it].z, NAVC_ACT_GELU_NEW); v[it].w = act_apply_fast(v[it].w, NAVC_ACT_GELU_NEW);
                            }
                        }
                        const bool col_ok = col2 < N && epi.dbg != 3;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rl = it * 4 + rr;
                            const float keep = ((rz_mask >> rl) & 1u) ? 0.f : 1.f;
                            v[it].x = (v[it].x + q[it].x) * keep; v[it].y = (v[it].y + q[it].y) * keep;
                            v[it].z = (v[it].z + q[it].z) * keep; v[it].w = (v[it].w + q[it].w) * keep;
                        }
                        if (epi.accumulate) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int grow = row0 + it * 4 + rr;
                                if (grow < M && col_ok) atomicAdd(reinterpret_cast<float4*>(epi.out_f32 + (size_t)grow * epi.ld_out + col2), v[it]);
                            }
                        } else if (epi.out_f32) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int grow = row0 + it * 4 + rr;
                                if (grow < M && col_ok) *reinterpret_cast<float4*>(epi.out_f32 + (size_t)grow * epi.ld_out + col2) = v[it];
                            }
                        }
                        if (epi.out_hi && !epi.accumulate) {
                            uint2 hv[8], lv[8];
#pragma unroll
                            for (int it = 0; it < 8; ++it) split_bf16x4(v[it], hv[it], lv[it]);
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int grow = row0 + it * 4 + rr;
                                if (grow < M && col_ok) *reinterpret_cast<uint2*>(epi.out_hi + (size_t)grow * epi.ld_out + col2) = hv[it];
                            }
                            if (epi.out_lo) {
#pragma unroll
                                for (int it = 0; it < 8; ++it) {
                                    const int grow = row0 + it * 4 + rr;
                                    if (grow < M && col_ok) *reinterpret_cast<uint2*>(epi.out_lo + (size_t)grow * epi.ld_out + col2) = lv[it];
                                }
                            }
                        }
                        __syncwarp();
                    } else {
                        // odd leading dimensions / N: scalar fallback through the staging buffer (lane = row)
#pragma unroll
                        for (int j = 0; j < 32; ++j) stage[lane * 32 + ((j + lane) & 31)] = __uint_as_float(r[j]);
                        __syncwarp();
                        if (c + 1 < n_chunks) tc_ld32(t_row + (uint32_t)((c + 1) * 32), r);
                        if (row_ok) {
#pragma unroll 1
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) epi_store1(epi, row, col0 + j, stage[lane * 32 + ((j + lane) & 31)], my_rz);
                        }
                        __syncwarp();
                    }
                }
            } else {
                const int64_t tgt = (vep.target && row < M) ? vep.target[row] : -1;
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TBN + half * (TBN / 2));
                float run_m = -INFINITY, run_s = 0.f;
                int run_i = 0x7fffffff;
                constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll 1
                for (int c = 0; c < TBN / 64; ++c) {
                    const int col0 = n0 + c * 32;
                    if (col0 >= N) break;
                    uint32_t r[32];
                    tc_ld32(t_row + (uint32_t)(c * 32), r);
                    tc_wait_ld();
                    float cm = -INFINITY;
                    int ci = 0x7fffffff;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = col0 + j;
                        float v = __uint_as_float(r[j]);
                        if (col < N) {
                            if (vep.bias) v += __ldg(vep.bias + col);
                            if (v > cm) { cm = v; ci = col; }
                            if ((int64_t)col == tgt) vep.target_logit[row] = v;
                        } else {
                            v = -INFINITY;
                        }
                        r[j] = __float_as_uint(v);
                    }
                    const float nm = fmaxf(run_m, cm);
                    float s = (run_s == 0.f) ? 0.f : run_s * fast_exp2((run_m - nm) * kLog2e);
                    const float nm2 = nm * kLog2e;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j]);
                        s += fast_exp2(fmaf(v, kLog2e, -nm2));  // exp2(-inf) = 0 for the padded columns
                    }
                    if (cm > run_m) run_i = ci;  // strict: earlier chunk keeps ties (lowest column)
                    run_m = nm;
                    run_s = s;
                }
                if (row < M && n0 < N) {
                    const size_t o = (size_t)row * vep.n_tiles + (nb * 2 + half);
                    vep.part_max[o] = run_m; vep.part_sum[o] = run_s; vep.part_idx[o] = run_i;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
    }

    if constexpr (kPairAny) {
        if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // staging stays valid until read
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        // measured on config 2 (CUDA-graph replay): 14.9 ms with, 14.8 ms without -> off unless NAVC_PDL=1
        const char* e = getenv("NAVC_PDL");
        g_pdl = (e && (e[0] == '1' || e[0] == 'y' || e[0] == 't')) ? 1 : 0;
    }
    return g_pdl != 0;
}
static bool g_tc_ready = false;
bool tc_ready() { return g_tc_ready; }

// tail split workspace: (#SMs) partial tiles of 128 x 256 fp32 + 2 counters per CTA + an error flag, allocated once
static float* g_sk_ws = nullptr;
static int* g_sk_cnt = nullptr;
static int g_sk_slots = 0;
static int g_sk_on = -1;
static bool sk_enabled() {
    if (g_sk_on < 0) {
        // measured (round 1): config-2 step 14.59 ms with, 14.19 ms without; AR beam batch 33.5 vs 30.5 ms -- the
        // hand-over through L2 and the partner -> finisher serialisation cost more than the idle SMs of the last
        // wave.  Off unless NAVC_STREAMK=1 / navc_set_streamk(1).
        const char* e = getenv("NAVC_STREAMK");
        g_sk_on = (e && (e[0] == '1' || e[0] == 'y' || e[0] == 't')) ? 1 : 0;
    }
    return g_sk_on != 0 && g_sk_ws != nullptr;
}

int tc_init() {
    if (g_tc_ready) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NAVC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    NAVC_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "navc_init: cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (!g_sk_ws) {
        int sms = navc_sm_count();
        if (sms <= 0) sms = 148;
        g_sk_slots = sms;
        NAVC_CUDA(cudaMalloc(&g_sk_ws, (size_t)g_sk_slots * TBM * 256 * sizeof(float)));
        NAVC_CUDA(cudaMalloc(&g_sk_cnt, (size_t)(2 * g_sk_slots + 1) * sizeof(int)));
        NAVC_CUDA(cudaMemset(g_sk_cnt, 0, (size_t)(2 * g_sk_slots + 1) * sizeof(int)));
    }
#define NAVC_TC_ATTR(X3, EPI, BN) \
    NAVC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<X3, EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<X3, BN, EPI>::kSmemBytes))
    NAVC_TC_ATTR(false, kEpiGeneric, 256); NAVC_TC_ATTR(true, kEpiGeneric, 256);
    NAVC_TC_ATTR(false, kEpiGeneric, 128); NAVC_TC_ATTR(true, kEpiGeneric, 128);
    NAVC_TC_ATTR(false, kEpiPair, 256); NAVC_TC_ATTR(true, kEpiPair, 256);
    NAVC_TC_ATTR(false, kEpiPair, 128); NAVC_TC_ATTR(true, kEpiPair, 128);
    NAVC_TC_ATTR(false, kEpiPairRes, 128); NAVC_TC_ATTR(true, kEpiPairRes, 128);
    NAVC_TC_ATTR(false, kEpiVocab, 256); NAVC_TC_ATTR(true, kEpiVocab, 256);
    NAVC_TC_ATTR(false, kEpiWgrad, 256); NAVC_TC_ATTR(true, kEpiWgrad, 256);
    NAVC_TC_ATTR(false, kEpiWgrad, 128); NAVC_TC_ATTR(true, kEpiWgrad, 128);
    NAVC_TC_ATTR(false, kEpiDgrad, 256); NAVC_TC_ATTR(true, kEpiDgrad, 256);
    NAVC_TC_ATTR(false, kEpiDgrad, 128); NAVC_TC_ATTR(true, kEpiDgrad, 128);
#undef NAVC_TC_ATTR
    NAVC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, kEpiGeneric, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<false, 256, kEpiGeneric>::kSmemBytes));
    NAVC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, kEpiGeneric, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<false, 128, kEpiGeneric>::kSmemBytes));
    g_tc_ready = true;
    return 0;
}

// Encoded tensor maps are pure functions of (pointer, shape, leading dimension, box): a small per-thread direct-mapped
// cache saves the ~1 us driver call per map (a GEMM launch builds 6-8 of them: 10+ us of host time per launch, which
// paced the eager paths -- training steps, the per-class timing of bench.py -- more than the kernels did).
int tc_make_map_uncached(CUtensorMap* map, const uint16_t* ptr, int rows, int K, int ld, int box_rows);
struct MapKey { const void* ptr; int rows, cols, ld, box, kind; };
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
static thread_local MapSlot g_map_cache[512];
static inline bool map_cache_get(const MapKey& k, CUtensorMap* out, MapSlot** slot) {
    const uint64_t h = ((uint64_t)(uintptr_t)k.ptr >> 4) * 0x9E3779B97F4A7C15ull ^ ((uint64_t)k.rows * 1315423911u) ^ ((uint64_t)k.cols << 20) ^
                       ((uint64_t)k.ld << 7) ^ ((uint64_t)k.box << 3) ^ (uint64_t)k.kind;
    MapSlot* s = &g_map_cache[(h >> 17) & 511];
    *slot = s;
    if (s->used && s->key.ptr == k.ptr && s->key.rows == k.rows && s->key.cols == k.cols && s->key.ld == k.ld && s->key.box == k.box &&
        s->key.kind == k.kind) {
        *out = s->map;
        return true;
    }
    return false;
}

// 2-D bf16 tensor map over a row-major [rows, K] matrix with leading dimension ld, box = [box_rows, 64].
int tc_make_map(CUtensorMap* map, const uint16_t* ptr, int rows, int K, int ld, int box_rows) {
    const MapKey key = {ptr, rows, K, ld, box_rows, 0};
    MapSlot* slot;
    if (map_cache_get(key, map, &slot)) return 0;
    if (tc_make_map_uncached(map, ptr, rows, K, ld, box_rows)) return 1;
    slot->key = key; slot->map = *map; slot->used = true;
    return 0;
}
int tc_make_map_uncached(CUtensorMap* map, const uint16_t* ptr, int rows, int K, int ld, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(ptr), gdim, gstride, box,
                          estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NAVC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
    return 0;
}

// 2-D fp32 tensor map (kind::tf32 operands): box = [box_rows, 32 floats] = the same 128-byte swizzle rows
static int tc_make_map_f32(CUtensorMap* map, const float* ptr, int rows, int K, int ld, int box_rows) {
    const MapKey key = {ptr, rows, K, ld, box_rows, 2};
    MapSlot* slot;
    if (map_cache_get(key, map, &slot)) return 0;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)(TBK / 2), (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estride,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NAVC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
    slot->key = key; slot->map = *map; slot->used = true;
    return 0;
}

// 2-D bf16 tensor map for the epilogue's TMA stores: box = [32 rows, 16 columns], no swizzle.
static int tc_make_store_map_uncached(CUtensorMap* map, const uint16_t* ptr, int rows, int cols, int ld);
static int tc_make_store_map(CUtensorMap* map, const uint16_t* ptr, int rows, int cols, int ld) {
    const MapKey key = {ptr, rows, cols, ld, 32, 1};
    MapSlot* slot;
    if (map_cache_get(key, map, &slot)) return 0;
    if (tc_make_store_map_uncached(map, ptr, rows, cols, ld)) return 1;
    slot->key = key; slot->map = *map; slot->used = true;
    return 0;
}
static int tc_make_store_map_uncached(CUtensorMap* map, const uint16_t* ptr, int rows, int cols, int ld) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {16u, 32u};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(ptr), gdim, gstride, box,
                          estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NAVC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return 0;
}

int tc_make_store_map16(CUtensorMap* map, const uint16_t* ptr, int rows, int cols, int ld) {   // gemm2_tc.cu
    return tc_make_store_map(map, ptr, rows, cols, ld);
}
int g2_linear(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w_hi, const uint16_t* w_lo, int ldw,
              int M, int N, int K, const EpiParams& epi, cudaStream_t st);   // gemm2_tc.cu
bool g2_enabled();
bool g2_forced();

static thread_local int g_dgrad_w_rows = 0;  // true row count of the dgrad B operand (launch_tc_dgrad)

static int tile_waste_pct(int tiles, int sms) {  // idle share of the last wave of a persistent grid, in percent
    const int waves = (tiles + sms - 1) / sms;
    return 100 - (100 * tiles) / (waves * sms);
}

template <int kEpi, int TBN>
static int launch_tc_bn(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w_hi,
                        const uint16_t* w_lo, int ldw, int M, int N, int K, const EpiParams& epi, const TcVocab& vep,
                        cudaStream_t st, const char* what) {
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    if (kEpi == kEpiWgrad) {
        // operands are [K reduction rows, M] and [K reduction rows, N] row-major: 64 x 64 boxes
        if (tc_make_map(&ma_hi, x_hi, K, M, ldx, 64)) return 1;
        if (tc_make_map(&mb_hi, w_hi, K, N, ldw, 64)) return 1;
        if (mode == NAVC_TC_BF16X3) {
            if (tc_make_map(&ma_lo, x_lo, K, M, ldx, 64)) return 1;
            if (tc_make_map(&mb_lo, w_lo, K, N, ldw, 64)) return 1;
        } else {
            ma_lo = ma_hi;
            mb_lo = mb_hi;
        }
    } else if (kEpi == kEpiDgrad) {
        // A [M, K reduction] K-major as usual; B [K reduction rows, N] row-major: 64 x 64 boxes
        if (tc_make_map(&ma_hi, x_hi, M, K, ldx, TBM)) return 1;
        const int w_rows = g_dgrad_w_rows > 0 ? g_dgrad_w_rows : K;
        if (tc_make_map(&mb_hi, w_hi, w_rows, N, ldw, 64)) return 1;
        if (mode == NAVC_TC_BF16X3) {
            if (tc_make_map(&ma_lo, x_lo, M, K, ldx, TBM)) return 1;
            if (tc_make_map(&mb_lo, w_lo, w_rows, N, ldw, 64)) return 1;
        } else {
            ma_lo = ma_hi;
            mb_lo = mb_hi;
        }
    } else if (tc_make_map(&ma_hi, x_hi, M, K, ldx, TBM) || tc_make_map(&mb_hi, w_hi, N, K, ldw, TBN)) {
        return 1;
    } else if (mode == NAVC_TC_BF16X3) {
        if (tc_make_map(&ma_lo, x_lo, M, K, ldx, TBM)) return 1;
        if (tc_make_map(&mb_lo, w_lo, N, K, ldw, TBN)) return 1;
    } else {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }
    CUtensorMap mo_hi = ma_hi, mo_lo = ma_hi, mr_hi = ma_hi, mr_lo = ma_hi;
    if (kEpi == kEpiPair || kEpi == kEpiPairRes) {
        if (tc_make_store_map(&mo_hi, epi.out_hi, M, N, epi.ld_out)) return 1;
        if (epi.out_lo && tc_make_store_map(&mo_lo, epi.out_lo, M, N, epi.ld_out)) return 1;
    }
    if (kEpi == kEpiPairRes) {
        if (tc_make_store_map(&mr_hi, epi.res_hi, M, N, epi.ld_res)) return 1;
        if (epi.res_lo && tc_make_store_map(&mr_lo, epi.res_lo, M, N, epi.ld_res)) return 1;
    }
    const int tiles = ((M + TBM - 1) / TBM) * ((N + TBN - 1) / TBN) * epi.split_k;
    int sms = navc_sm_count();
    if (sms <= 0) sms = 148;
    int grid = tiles < sms ? tiles : sms;
    EpiParams epi_sk = epi;
    if (kEpi == kEpiPair && epi.split_k == 1 && sk_enabled() && sms <= g_sk_slots && (K + TBK - 1) / TBK >= 16) {
        // tail split: needs the whole grid resident (one CTA per SM) -- the finishing part waits for its partners
        grid = sms;
        epi_sk.sk_ws = g_sk_ws;
        epi_sk.sk_cnt = g_sk_cnt;
        epi_sk.sk_err = g_sk_cnt + 2 * g_sk_slots;
    }
    if (mode == NAVC_TC_BF16X3) {
        NAVC_CUDA(launch_pdl(gemm_tc_kernel<true, kEpi, TBN>, dim3(grid), dim3(tc_threads(kEpi)), TcCfg<true, TBN, kEpi>::kSmemBytes, st,
                             ma_hi, ma_lo, mb_hi, mb_lo, mo_hi, mo_lo, mr_hi, mr_lo, M, N, K, epi_sk, vep));
    } else {
        NAVC_CUDA(launch_pdl(gemm_tc_kernel<false, kEpi, TBN>, dim3(grid), dim3(tc_threads(kEpi)), TcCfg<false, TBN, kEpi>::kSmemBytes, st,
                             ma_hi, ma_lo, mb_hi, mb_lo, mo_hi, mo_lo, mr_hi, mr_lo, M, N, K, epi_sk, vep));
    }
    return check_launch(what);
}

template <int kEpi>
static int launch_tc(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w_hi,
                     const uint16_t* w_lo, int ldw, int M, int N, int K, const EpiParams& epi, const TcVocab& vep,
                     cudaStream_t st, const char* what) {
    NAVC_REQUIRE(g_tc_ready, "%s: navc_init() has not been called", what);
    NAVC_REQUIRE(mode == NAVC_TC_BF16 || mode == NAVC_TC_BF16X3, "%s: bad mode %d", what, mode);
    NAVC_REQUIRE(x_hi && w_hi && (mode == NAVC_TC_BF16 || (x_lo && w_lo)), "%s: null operand", what);
    NAVC_REQUIRE(M > 0 && N > 0 && K > 0 && (kEpi == kEpiWgrad || K % 8 == 0) && ldx % 8 == 0 && ldw % 8 == 0,
                 "%s: need K%%8==0 and ld%%8==0 (M=%d N=%d K=%d ldx=%d ldw=%d)", what, M, N, K, ldx, ldw);
    NAVC_REQUIRE((((uintptr_t)x_hi | (uintptr_t)w_hi | (uintptr_t)x_lo | (uintptr_t)w_lo) & 15) == 0,
                 "%s: operands must be 16-byte aligned", what);
    if constexpr (kEpi != kEpiVocab) {
        // 128-wide tiles when they fill the last wave of the persistent grid markedly better
        int sms = navc_sm_count();
        if (sms <= 0) sms = 148;
        const int mt = (M + TBM - 1) / TBM;
        const int t256 = mt * ((N + 255) / 256) * epi.split_k, t128 = mt * ((N + 127) / 128) * epi.split_k;
        const int forced = epi.dbg >= 128 ? epi.dbg : 0;  // profiling aid: reserved = 128 / 256 forces a tile width
        // (fewer 256-wide tiles than SMs: the narrow tiling doubles the CTAs that share the operand streaming)
        const bool narrow = forced ? forced == 128
                                   : (N <= 128 || (t256 < sms && t128 > t256) || tile_waste_pct(t256, sms) >= tile_waste_pct(t128, sms) + 8);
        if constexpr (kEpi == kEpiPair) {
            // a bf16 hi/lo residual is prefetched by TMA (128-wide tiles only: the boxes need 64 KB of the ring)
            // (plain bf16 only: the split mode cannot spare two of its three 64 KB stages' worth of ring)
            if (mode == NAVC_TC_BF16 && epi.res_hi && (narrow || N <= 1024) && forced != 256)
                return launch_tc_bn<kEpiPairRes, 128>(mode, x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, epi, vep, st, what);
        }
        if (narrow) return launch_tc_bn<kEpi, 128>(mode, x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, epi, vep, st, what);
    }
    return launch_tc_bn<kEpi, 256>(mode, x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, epi, vep, st, what);
}

// Y = epilogue(X W^T) with fp32 operands consumed as TF32 (generic epilogue: fp32 / bf16 hi / lo outputs, fp32 residual)
template <int TBN>
static int launch_tf32_bn(const float* x, int ldx, const float* w, int ldw, int M, int N, int K, const EpiParams& epi, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (tc_make_map_f32(&ma, x, M, K, ldx, TBM) || tc_make_map_f32(&mb, w, N, K, ldw, TBN)) return 1;
    const int tiles = ((M + TBM - 1) / TBM) * ((N + TBN - 1) / TBN);
    int sms = navc_sm_count();
    if (sms <= 0) sms = 148;
    const int grid = tiles < sms ? tiles : sms;
    TcVocab v = {};
    NAVC_CUDA(launch_pdl(gemm_tc_kernel<false, kEpiGeneric, TBN, true>, dim3(grid), dim3(tc_threads(kEpiGeneric)),
                         TcCfg<false, TBN, kEpiGeneric>::kSmemBytes, st, ma, ma, mb, mb, ma, ma, ma, ma, M, N, K, epi, v));
    return check_launch("navc_linear_tf32");
}

// dX[rows, k_in] = dY[rows, kred] W[n_out, k_in]: the B map is declared with its true n_out rows so that the
// reduction tail (kred > n_out, and the k-block tail) reads zeros.
static int launch_tc_dgrad(int mode, const uint16_t* dy_hi, const uint16_t* dy_lo, int ld_dy, const uint16_t* w_hi,
                           const uint16_t* w_lo, int ld_w, int rows, int n_out, int k_in, int kred, const EpiParams& epi,
                           const TcVocab& vep, cudaStream_t st) {
    g_dgrad_w_rows = n_out;
    const int rc = launch_tc<kEpiDgrad>(mode, dy_hi, dy_lo, ld_dy, w_hi, w_lo, ld_w, rows, k_in, kred, epi, vep, st, "navc_dgrad_tc");
    g_dgrad_w_rows = 0;
    return rc;
}

}  // namespace navc

using namespace navc;

// 1 if a finishing CTA of a tail split ever gave up waiting for its partners (a bug; results of that launch are
// wrong), else 0; -1 without the workspace.  Synchronises the device: tests only.
extern "C" int navc_set_streamk(int on) {
    const int prev = g_sk_on > 0 ? 1 : 0;
    g_sk_on = on ? 1 : 0;
    return prev;
}

extern "C" int navc_streamk_error(void) {
    if (!g_sk_cnt) return -1;
    int flag = 0;
    if (cudaMemcpy(&flag, g_sk_cnt + 2 * g_sk_slots, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return flag;
}

extern "C" int navc_vocab_tile(int tc) { return tc ? kVocabBN / 2 : 128; }

extern "C" int navc_linear_tc(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w_hi,
                              const uint16_t* w_lo, int ldw, int M, int N, int K, const navc_epilogue_t* e,
                              void* stream) {
    NAVC_REQUIRE(e && (e->out_f32 || e->out_hi), "navc_linear_tc: no output");
    EpiParams p = to_params(e);
    if (p.accumulate) {
        NAVC_REQUIRE(p.out_f32 && !p.bias && !p.residual && !p.row_tokens && p.act == NAVC_ACT_NONE,
                     "navc_linear_tc: accumulate / split_k outputs take no bias, activation, residual or row mask");
        // every split must own at least one 64-wide k-block
        const int k_blocks = (K + TBK - 1) / TBK;
        if (p.split_k > k_blocks) p.split_k = k_blocks;
        const int kpb = (k_blocks + p.split_k - 1) / p.split_k;
        p.split_k = (k_blocks + kpb - 1) / kpb;
    }
    TcVocab v = {};
    // fast path: bf16 pair outputs only, optional bf16 pair residual, everything 16-byte aligned
    const bool pair = !p.accumulate && !p.out_f32 && !p.residual && p.out_hi && p.dbg != 9 && p.ld_out % 8 == 0 && N % 8 == 0 &&
                      ((((uintptr_t)p.out_hi) | ((uintptr_t)p.out_lo)) & 15) == 0 && (((uintptr_t)p.bias) & 15) == 0 &&
                      (!p.res_hi || (p.ld_res % 8 == 0 && ((((uintptr_t)p.res_hi) | ((uintptr_t)p.res_lo)) & 15) == 0));
    NAVC_REQUIRE(pair || !p.res_hi, "navc_linear_tc: a bf16 hi/lo residual needs bf16-only outputs (no out_f32 / fp32 residual), "
                                    "N %% 8 == 0 and 16-byte aligned operands");
    // second-generation kernel (CTA pairs, tail split along N) in the split-bf16 mode, where it measured faster (config-2
    // step 13.98 vs 14.10 ms); plain bf16 keeps the first-generation kernel with its TMA-prefetched residual boxes
    // (7.87 vs 9.19 ms) unless NAVC_GEMM2=force
    // NAVC_GEMM2_SMALL=0 keeps the out / query projections (N <= 512, K <= 512) on the first-generation kernel (A/B runs)
    static const bool g2_small = !(getenv("NAVC_GEMM2_SMALL") && getenv("NAVC_GEMM2_SMALL")[0] == '0');
    const bool g2_shape = (N > 512 || K > 512) || g2_small || g2_forced();
    if (pair && g2_enabled() && g2_shape && (mode == NAVC_TC_BF16X3 || g2_forced()) && !sk_enabled() && (p.dbg == 0 || p.dbg == 7 || (p.dbg >= 11 && p.dbg <= 15) || p.dbg == 128 || p.dbg == 256)) {
        // second-generation kernel (gemm2_tc.cu): cluster multicast of the weight tile, tail split along N
        NAVC_REQUIRE(mode == NAVC_TC_BF16 || mode == NAVC_TC_BF16X3, "navc_linear_tc: bad mode %d", mode);
        NAVC_REQUIRE(x_hi && w_hi && (mode == NAVC_TC_BF16 || (x_lo && w_lo)), "navc_linear_tc: null operand");
        NAVC_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0,
                     "navc_linear_tc: need K%%8==0 and ld%%8==0 (M=%d N=%d K=%d ldx=%d ldw=%d)", M, N, K, ldx, ldw);
        NAVC_REQUIRE((((uintptr_t)x_hi | (uintptr_t)w_hi | (uintptr_t)x_lo | (uintptr_t)w_lo) & 15) == 0,
                     "navc_linear_tc: operands must be 16-byte aligned");
        return g2_linear(mode, x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, p, as_stream(stream));
    }
    if (pair) return launch_tc<kEpiPair>(mode, x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, p, v, as_stream(stream), "navc_linear_tc");
    return launch_tc<kEpiGeneric>(mode, x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, p, v, as_stream(stream), "navc_linear_tc");
}

extern "C" int navc_linear_tf32(const float* x, int ldx, const float* w, int ldw, int M, int N, int K, const navc_epilogue_t* e,
                                void* stream) {
    NAVC_REQUIRE(e && (e->out_f32 || e->out_hi), "navc_linear_tf32: no output");
    NAVC_REQUIRE(g_tc_ready, "navc_linear_tf32: navc_init() has not been called");
    EpiParams p = to_params(e);
    NAVC_REQUIRE(!p.accumulate && p.split_k <= 1 && !p.res_hi, "navc_linear_tf32: no split-K / accumulate / bf16-pair residual");
    p.split_k = 1;
    NAVC_REQUIRE(x && w && M > 0 && N > 0 && K > 0 && K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)w)) & 15) == 0,
                 "navc_linear_tf32: need K %% 4 == 0, ld %% 4 == 0 and 16-byte aligned operands (M=%d N=%d K=%d)", M, N, K);
    int sms = navc_sm_count();
    if (sms <= 0) sms = 148;
    const int mt = (M + TBM - 1) / TBM;
    const int t256 = mt * ((N + 255) / 256), t128 = mt * ((N + 127) / 128);
    const bool narrow = N <= 128 || (t256 < sms && t128 > t256) || tile_waste_pct(t256, sms) >= tile_waste_pct(t128, sms) + 8;
    if (narrow) return launch_tf32_bn<128>(x, ldx, w, ldw, M, N, K, p, as_stream(stream));
    return launch_tf32_bn<256>(x, ldx, w, ldw, M, N, K, p, as_stream(stream));
}

extern "C" int navc_wgrad_tc(int mode, const uint16_t* dy_hi, const uint16_t* dy_lo, int ld_dy, const uint16_t* x_hi,
                             const uint16_t* x_lo, int ld_x, int rows, int n_out, int k_in, const navc_epilogue_t* e,
                             void* stream) {
    NAVC_REQUIRE(e && e->out_f32, "navc_wgrad_tc: needs out_f32");
    EpiParams p = to_params(e);
    NAVC_REQUIRE(!p.bias && !p.residual && !p.row_tokens && p.act == NAVC_ACT_NONE && !p.res_hi && !p.m_dev,
                 "navc_wgrad_tc: the weight-gradient GEMM takes no bias / activation / residual / row mask");
    NAVC_REQUIRE(rows > 0 && n_out % 8 == 0 && k_in % 8 == 0, "navc_wgrad_tc: n_out and k_in must be multiples of 8");
    p.accumulate = 1;
    const int k_blocks = (rows + TBK - 1) / TBK;
    if (p.split_k > k_blocks) p.split_k = k_blocks;
    const int kpb = (k_blocks + p.split_k - 1) / p.split_k;
    p.split_k = (k_blocks + kpb - 1) / kpb;
    TcVocab v = {};
    return launch_tc<kEpiWgrad>(mode, dy_hi, dy_lo, ld_dy, x_hi, x_lo, ld_x, n_out, k_in, rows, p, v, as_stream(stream),
                                "navc_wgrad_tc");
}

extern "C" int navc_dgrad_tc(int mode, const uint16_t* dy_hi, const uint16_t* dy_lo, int ld_dy, const uint16_t* w_hi,
                             const uint16_t* w_lo, int ld_w, int rows, int n_out, int k_in, const navc_epilogue_t* e,
                             void* stream) {
    NAVC_REQUIRE(e && e->out_f32, "navc_dgrad_tc: needs out_f32");
    EpiParams p = to_params(e);
    NAVC_REQUIRE(!p.res_hi && !p.m_dev && !p.accumulate && p.split_k <= 1, "navc_dgrad_tc: no bf16 residual / device row count / split-K");
    NAVC_REQUIRE(rows > 0 && n_out > 0 && k_in % 8 == 0 && ld_dy >= n_out, "navc_dgrad_tc: k_in must be a multiple of 8");
    p.split_k = 1;
    // reduction over the n_out weight rows; the A map is declared 8-aligned wide (dY's pad columns, if any, meet
    // weight rows beyond n_out, which TMA delivers as zeros)
    int kred = (n_out + 7) / 8 * 8;
    if (kred > ld_dy) kred = ld_dy;
    NAVC_REQUIRE(kred % 8 == 0, "navc_dgrad_tc: ld_dy must be a multiple of 8");
    TcVocab v = {};
    return launch_tc_dgrad(mode, dy_hi, dy_lo, ld_dy, w_hi, w_lo, ld_w, rows, n_out, k_in, kred, p, v, as_stream(stream));
}

extern "C" int navc_vocab_partials_tc(int mode, const uint16_t* h_hi, const uint16_t* h_lo, int ldh,
                                      const uint16_t* w_hi, const uint16_t* w_lo, int ldw, const float* bias, int M,
                                      int V, int K, float* part_max, float* part_sum, int32_t* part_idx,
                                      const int64_t* target, float* target_logit, void* stream) {
    NAVC_REQUIRE(part_max && part_sum && part_idx, "navc_vocab_partials_tc: null output");
    NAVC_REQUIRE(!target || target_logit, "navc_vocab_partials_tc: target without target_logit");
    TcVocab v = {bias, part_max, part_sum, part_idx, target, target_logit, (V + kVocabBN / 2 - 1) / (kVocabBN / 2)};
    EpiParams e = {};
    e.split_k = 1;
    return launch_tc<kEpiVocab>(mode, h_hi, h_lo, ldh, w_hi, w_lo, ldw, M, V, K, e, v, as_stream(stream),
                           "navc_vocab_partials_tc");
}

extern "C" int navc_vocab_partials_tc_dyn(int mode, const uint16_t* h_hi, const uint16_t* h_lo, int ldh,
                                          const uint16_t* w_hi, const uint16_t* w_lo, int ldw, const float* bias, int M,
                                          int V, int K, const int32_t* m_dev, float* part_max, float* part_sum,
                                          int32_t* part_idx, void* stream) {
    NAVC_REQUIRE(part_max && part_sum && part_idx, "navc_vocab_partials_tc_dyn: null output");
    TcVocab v = {bias, part_max, part_sum, part_idx, nullptr, nullptr, (V + kVocabBN / 2 - 1) / (kVocabBN / 2)};
    EpiParams e = {};
    e.split_k = 1;
    e.m_dev = m_dev;
    return launch_tc<kEpiVocab>(mode, h_hi, h_lo, ldh, w_hi, w_lo, ldw, M, V, K, e, v, as_stream(stream),
                                "navc_vocab_partials_tc_dyn");
}
