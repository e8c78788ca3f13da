// tcgen05 GEMM of the inference hot path, second generation:  Y = epilogue(X[M,K] * W[N,K]^T), bf16 hi(/lo) outputs.
//
// Same building blocks as gemm_tc.cu (TMA 128B swizzle -> shared-memory ring -> tcgen05.mma, fp32 accumulators in
// TMEM, double buffered -> tcgen05.ld -> fused "pair" epilogue: bias / activation / bf16 hi+lo residual / non-pad row
// mask / bf16 hi+lo split, TMA stores), rebuilt around an in-kernel timeline of the round-1 kernel
// (navc_debug_trace / tools/gemm2_trace.py, profiles/r2f_*):
//
//  * CTA pairs (`tcgen05.mma.cta_group::2`, clusters of 2 CTAs on one TPC) compute 256-row tiles: each CTA stages its
//    own 128 A rows and only HALF of the B (weight) rows of the tile, the leader CTA issues UMMA 256 x N x 16 and the
//    tensor cores read the other half from the partner's shared memory: 64 KB instead of 96 KB per k-block enter each
//    SM for its 128 x 256 share of the tile (three ring stages instead of two).  TMA loads of both CTAs complete on
//    the leader's barrier, the leader's tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs, the
//    partner's epilogue warps arrive remotely on the leader's "accumulator drained" barrier.
//    (Tried first and dropped: TMA multicast of the B tile inside a 2-CTA cluster -- no gain, the bytes still enter
//    both SMs -- and 256-bit global stores instead of TMA store boxes -- the row-per-lane stores cost ~120 cycles each.)
//  * The bf16 hi/lo residual is read with 256-bit loads (a row-per-lane access costs the LSU per distinct line, not
//    per byte), the first step's before the accumulator barrier is awaited and every later step's one step ahead.
//    (Tried and dropped: preloading the residual into the TMEM accumulators with tcgen05.st so that the MMAs add onto
//    it -- correct, but the two preloads ahead of a CTA's first MMA cost 10 us at kernel start, and with 1-2 tiles per
//    CTA at the config-2 shapes there is no steady state to win it back: so 29.8 vs 30.8 us, f2 65.6 vs 61.2 us.)
//  * Epilogue steps are software pipelined: the tcgen05.ld of step c+1 is issued before step c is computed, and the hi
//    and lo boxes leave as separate bulk groups (`cp.async.bulk.wait_group.read 1`), so a staging box is only waited
//    for when the store issued a full half-step earlier has not drained yet.
//  * Tail split along N: the tiles of the last, partly filled wave of the persistent grid are cut into 2 or 4 column
//    parts (UMMA N = TBN/2, TBN/4; B arrives as 32-row boxes so a part loads only its own weight rows) that run on the
//    otherwise idle SMs.  Unlike a split along K (gemm_tc.cu's tail split, measured slower) nothing is handed over.
//
// Warp roles (576 threads, persistent, one CTA per SM): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer,
// warps 2..17 epilogue (warp = TMEM lane quarter x column group of TBN/4; lane = row; 16-column steps).
#include <stdlib.h>

#include "tc_common.cuh"

namespace navc {

constexpr int G2_BM = 128, G2_BK = 64;
constexpr int G2_EPI_WARPS = 16;
constexpr int G2_THREADS = 64 + 32 * G2_EPI_WARPS;
constexpr int G2_ACC = 2;            // TMEM accumulator stages (2 x TBN columns)
constexpr int G2_TILE_A = G2_BM * G2_BK * 2;   // 16 KB
constexpr int G2_PBOX = 32;          // rows of the B boxes used by the column parts of the tail wave

template <bool kX3, int TBN, int kCl = 1> struct G2Cfg {
    static constexpr int kTileB = (TBN / kCl) * G2_BK * 2;         // a CTA of a pair stages half of the B rows
    static constexpr int kStageBytes = (kX3 ? 2 : 1) * (G2_TILE_A + kTileB);
    static constexpr int kStagesRaw = 196608 / kStageBytes;        // 1 CTA: 2 (x3, 256) / 3 (x3, 128) / 4 (bf16, 256) / 6 (bf16, 128)
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw; // pair:  3 (x3, 256) / 4 (x3, 128) / 6 (bf16, 256) / 8 (bf16, 128)
    static constexpr int kRingBytes = kStages * kStageBytes;
    static constexpr int kStgBytes = G2_EPI_WARPS * 2048;          // per warp: hi box (1 KB) + lo box (1 KB)
    static constexpr int kSmemBytes = kRingBytes + kStgBytes + 1024 /*align*/ + 512 /*barriers*/;
    static_assert(kStages >= 2, "operand ring needs at least two stages");
};

struct G2Unit { int mb, n0, w; bool valid; int prob; };

// Second problem of a CHAINED launch (kChain): Y1 = epilogue1(Y0 W1^T) right behind Y0 = epilogue0(X W0^T) in ONE
// persistent grid.  Both problems have the same M (device side), N, K and tile configuration; a tile of problem 1 needs
// the whole row block of Y0, so the epilogue warps of problem 0 publish per-row-block completion counts (after their TMA
// stores have completed) and the producer of a problem-1 tile acquires them before its first A load.  cnt: [m_blocks]
// int32, zero at launch (the host clears it).
struct G2Chain {
    CUtensorMap a_hi, a_lo, b_hi, b_lo, p_hi, p_lo, o_hi, o_lo;
    EpiParams epi;
    int* cnt;
};

// Optional in-kernel timeline (tools/gemm2_trace.py): when navc_debug_trace() has installed a buffer, the producer
// lane, the MMA lane and lane 0 of epilogue warp 2 of every CTA record (tag, clock64) pairs -- 3 roles x 64 events.
__device__ unsigned long long* g2_trace_buf = nullptr;
struct G2Trace {
    unsigned long long* p;
    int n;
    __device__ G2Trace(int role) : p(nullptr), n(0) {
        if (g2_trace_buf) p = g2_trace_buf + ((size_t)blockIdx.x * 3 + role) * 64;
    }
    __device__ __forceinline__ void ev(unsigned tag) {
        if (p && n < 64) p[n++] = ((unsigned long long)tag << 56) | ((unsigned long long)clock64() & 0x00ffffffffffffffull);
    }
};

template <bool kX3, int TBN, int kCl, bool kRes, bool kChain>
__device__ __forceinline__ void gemm2_tc_body(const CUtensorMap& map_a_hi, const CUtensorMap& map_a_lo,
                const CUtensorMap& map_b_hi, const CUtensorMap& map_b_lo,
                const CUtensorMap& map_p_hi, const CUtensorMap& map_p_lo,
                const CUtensorMap& map_o_hi, const CUtensorMap& map_o_lo,
                int M_max, int N, int K, const EpiParams& epi, const G2Chain* chain) {
    using Cfg = G2Cfg<kX3, TBN, kCl>;
    constexpr int kTileB = Cfg::kTileB;
    constexpr bool kPair = kCl == 2;
    constexpr int kParts = kX3 ? 2 : 1;
    const int M = epi.m_dev ? min(M_max, __ldg(epi.m_dev)) : M_max;   // device-side row count (packed rows)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    uint8_t* stg_all = smem_gen + Cfg::kRingBytes;
    const uint32_t bar_off = Cfg::kRingBytes + Cfg::kStgBytes;
    const uint32_t bar_base = smem_base + bar_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + G2_ACC + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + bar_off + 8 * (2 * Cfg::kStages + 2 * G2_ACC));
    auto pub_bar = [&](int i) { return bar_base + 256u + 8u * (uint32_t)(i & 7); };   // chained launches: "tile i of problem 0 stored"

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = kPair ? (int)cluster_ctarank() : 0;
    const int cid = (int)blockIdx.x / kCl, G = (int)gridDim.x / kCl;   // cluster index, clusters in the grid

    // ---- schedule: a unit = one (row group of kCl m-blocks, column block) tile per cluster; the units of the last,
    //      partly filled wave are cut into `parts` column parts of TBN / parts columns ----
    const int m_blocks = (M + G2_BM - 1) / G2_BM, n_blocks = (N + TBN - 1) / TBN;
    const int k_blocks = (K + G2_BK - 1) / G2_BK;   // TMA zero-fills the K tail
    const int n_units1 = ((m_blocks + kCl - 1) / kCl) * n_blocks;   // units of one problem
    const int n_units = kChain ? 2 * n_units1 : n_units1;            // chained: problem 0's units, then problem 1's
    const int full = (n_units / G) * G, rem = n_units - full;
    int parts = 1;
    if (rem > 0 && epi.dbg != 7) {
        const int q = G / rem;
        parts = q >= 4 ? 4 : (q >= 2 ? 2 : 1);
        if (parts > TBN / 64) parts = TBN / 64;     // parts keep >= 64 columns
    }
    if (kChain && full < n_units1) parts = 1;   // chained: problem 0's tiles are always whole (their completion is counted per tile)
    const int total = full + rem * parts;
    auto get_unit = [&](int it, G2Unit& u) -> bool {
        const int t = cid + it * G;
        if (t >= total) return false;
        int ui, part = 0;
        u.w = TBN;
        if (t < full) {
            ui = t;
        } else {
            const int x = t - full;
            ui = full + x / parts;
            part = x - (x / parts) * parts;
            u.w = TBN / parts;
        }
        u.prob = 0;
        if (kChain && ui >= n_units1) { ui -= n_units1; u.prob = 1; }
        const int mg = ui / n_blocks, nb = ui - mg * n_blocks;   // column block fastest: concurrent clusters share the A rows in L2
        u.mb = mg * kCl + crank;
        u.n0 = nb * TBN + part * u.w;
        u.valid = u.mb < m_blocks && u.n0 < N;
        return true;
    };

    pdl_launch_dependents();
    if (g2_trace_buf && threadIdx.x == 0)   // timeline: kernel entry (slot 63 of the producer's record)
    {
        g2_trace_buf[((size_t)blockIdx.x * 3) * 64 + 63] = (60ull << 56) | ((unsigned long long)clock64() & 0x00ffffffffffffffull);
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        g2_trace_buf[((size_t)blockIdx.x * 3 + 1) * 64 + 63] = gt;   // wall clock (ns) of the entry: comparable across SMs
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        // pair: the leader's MMA thread waits for the epilogue warps of BOTH CTAs
        for (int s = 0; s < G2_ACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kCl * G2_EPI_WARPS); }
        if (kChain) for (int i = 0; i < 8; ++i) mbar_init(pub_bar(i), G2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kPair) {
        __syncthreads();
        cluster_sync_all();   // both CTAs' barriers are initialised before anything can reach them; both are present for the alloc
    }
    if (warp == 1) {
        if constexpr (kPair) {   // the same warp of both CTAs allocates the pair's columns
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            G2Unit u;
            G2Trace tr(0);
            tr.ev(1);   // prologue done
            for (int it = 0; get_unit(it, u); ++it) {
                // a column part loads only its own B rows, as 32-row boxes
                const bool part = u.w < TBN;
                const int brows = u.w / kCl;                         // B rows this CTA stages
                const uint32_t bytes = (uint32_t)(kParts * (G2_TILE_A + brows * 128));
                const int brow0 = u.n0 + crank * brows;
                const bool p1 = kChain && u.prob;
                const CUtensorMap* ma_hi = p1 ? &chain->a_hi : &map_a_hi; const CUtensorMap* ma_lo = p1 ? &chain->a_lo : &map_a_lo;
                const CUtensorMap* mb_hi = p1 ? &chain->b_hi : &map_b_hi; const CUtensorMap* mb_lo = p1 ? &chain->b_lo : &map_b_lo;
                const CUtensorMap* mp_hi = p1 ? &chain->p_hi : &map_p_hi; const CUtensorMap* mp_lo = p1 ? &chain->p_lo : &map_p_lo;
                if (p1 && u.valid) {
                    // every column tile of this row block of problem 0 has been published
                    const int need = n_blocks;   // one count per column tile of the row block (the publisher warp below)
                    const int* c = chain->cnt + u.mb;
                    int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(c) : "memory");
                        if (seen < need) __nanosleep(100);
                    } while (seen < need);
                    asm volatile("fence.proxy.async;" ::: "memory");   // the TMA loads below read what other SMs' TMA stores wrote
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);   // the MMAs reading this slot (of this CTA) have retired
                    if (kb == 0) tr.ev(2);                     // first load of a unit issued
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    if (epi.dbg == 15) {   // profiling aid: no operand loads (ring handshake only)
                        if (!kPair || crank == 0) mbar_arrive(full_bar(stage));
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    // pair: both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of both
                    const uint32_t fb = kPair ? mapa_u32(full_bar(stage), 0) : full_bar(stage);
                    if (crank == 0) mbar_expect_tx(full_bar(stage), kCl * bytes);
                    auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
                        if constexpr (kPair) tma_load_2d_2cta(dst, map, fb, c0, c1); else tma_load_2d(dst, map, fb, c0, c1);
                    };
                    load(sa, ma_hi, kb * G2_BK, u.mb * G2_BM);
                    if (kX3) load(sa + G2_TILE_A + kTileB, ma_lo, kb * G2_BK, u.mb * G2_BM);
                    if (!part) {
                        load(sa + G2_TILE_A, mb_hi, kb * G2_BK, brow0);
                        if (kX3) load(sa + 2 * G2_TILE_A + kTileB, mb_lo, kb * G2_BK, brow0);
                    } else {
                        for (int j = 0; j < brows / G2_PBOX; ++j) {
                            load(sa + G2_TILE_A + j * (G2_PBOX * 128), mp_hi, kb * G2_BK, brow0 + j * G2_PBOX);
                            if (kX3) load(sa + 2 * G2_TILE_A + kTileB + j * (G2_PBOX * 128), mp_lo, kb * G2_BK, brow0 + j * G2_PBOX);
                        }
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && crank == 0) {   // pair: only the leader CTA issues
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            G2Unit u;
            G2Trace tr(1);
            for (int it = 0; get_unit(it, u); ++it) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // the first use of a stage passes immediately
                tc_fence_after();
                tr.ev(4);   // accumulator stage free
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TBN);
                const uint32_t idesc = make_idesc(kCl * G2_BM, u.w);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (kb == 0) tr.ev(5);                  // first operands of the unit landed
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    const uint64_t da_hi = make_smem_desc(sa), db_hi = make_smem_desc(sa + G2_TILE_A);
                    const uint64_t da_lo = make_smem_desc(sa + G2_TILE_A + kTileB), db_lo = make_smem_desc(sa + 2 * G2_TILE_A + kTileB);
#pragma unroll
                    for (int k = 0; k < G2_BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
                        if (epi.dbg == 14 && (kb | k)) continue;   // profiling aid: one MMA per tile
                        const uint32_t first = (kb | k) ? 1u : 0u;
                        if constexpr (kPair) {
                            if (kX3) {
                                tc_mma_bf16_2cta(d_tmem, da_lo + koff, db_hi + koff, idesc, first);
                                tc_mma_bf16_2cta(d_tmem, da_hi + koff, db_lo + koff, idesc, 1u);
                                tc_mma_bf16_2cta(d_tmem, da_hi + koff, db_hi + koff, idesc, 1u);
                            } else {
                                tc_mma_bf16_2cta(d_tmem, da_hi + koff, db_hi + koff, idesc, first);
                            }
                        } else if (kX3) {
                            // small cross terms first, the dominant hi*hi product last
                            tc_mma_bf16(d_tmem, da_lo + koff, db_hi + koff, idesc, first);
                            tc_mma_bf16(d_tmem, da_hi + koff, db_lo + koff, idesc, 1u);
                            tc_mma_bf16(d_tmem, da_hi + koff, db_hi + koff, idesc, 1u);
                        } else {
                            tc_mma_bf16(d_tmem, da_hi + koff, db_hi + koff, idesc, first);
                        }
                    }
                    if constexpr (kPair) {
                        tc_commit_2cta_mc(empty_bar(stage), (uint16_t)0x3);               // slot free, in both CTAs
                        if (kb == k_blocks - 1) tc_commit_2cta_mc(tfull_bar(acc), (uint16_t)0x3);   // accumulators ready, in both CTAs
                    } else {
                        tc_commit(empty_bar(stage));
                        if (kb == k_blocks - 1) tc_commit(tfull_bar(acc));
                    }
                    if (kb == k_blocks - 1) tr.ev(6);       // last MMAs of the unit issued
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
                if (++acc == G2_ACC) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (kChain && warp == 2 + G2_EPI_WARPS) {
        // ===================== publisher (chained launches only) =====================
        // Making a tile's stores visible to the other SMs costs a gpu-scope fence of ~5 us (measured: in the epilogue warps it
        // doubled their time per tile).  So the epilogue warps only arrive on a CTA-local barrier; this otherwise idle warp
        // fences (cumulative over what the arrivals released) and moves the row block's count.
        if (lane == 0) {
            G2Unit u;
            int i0 = 0;
            for (int it = 0; get_unit(it, u); ++it) {
                if (u.prob || !u.valid) continue;
                mbar_wait_relaxed(pub_bar(i0), (uint32_t)((i0 >> 3) & 1));
                __threadfence();
                asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(chain->cnt + u.mb) : "memory");
                ++i0;
            }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        constexpr int GW = TBN / 4, STEPS = GW / 16;
        const int ew = warp - 2;
        const int quarter = warp & 3;   // TMEM lane quarter this warp may access
        const int grp = ew >> 2;        // column group [grp * GW, grp * GW + GW) of the tile
        const int cl0 = grp * GW;       // first tile-local column of this warp
        uint8_t* stg = stg_all + ew * 2048;   // hi box at +0, lo box at +1024
        const uint32_t stg_s = smem_u32(stg);
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        G2Trace tr(2);
        if (ew != 0 || lane != 0) tr.p = nullptr;

        auto arrive_drained = [&](int a) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar(a), 0));   // the leader's MMA thread waits for both CTAs
                else mbar_arrive(tempty_bar(a));
            }
        };
        G2Unit u;
        int acc = 0;
        uint32_t acc_phase = 0;
        int chain_i0 = 0;   // problem-0 tiles of this CTA so far (index of the publisher's barrier ring)
        for (int it = 0; get_unit(it, u); ++it) {
            const EpiParams& eu = (kChain && u.prob) ? chain->epi : epi;
            const CUtensorMap* mo_hi_u = (kChain && u.prob) ? &chain->o_hi : &map_o_hi;
            const CUtensorMap* mo_lo_u = (kChain && u.prob) ? &chain->o_lo : &map_o_lo;
            const int row0 = u.mb * G2_BM + quarter * 32;
            const int rowp = row0 + lane;
            const bool row_ok = u.valid && rowp < M;
            const bool rz = (row_ok && eu.row_tokens) ? (eu.row_tokens[rowp] == NAVC_PAD) : false;
            const bool active = u.valid && cl0 < u.w && u.n0 + cl0 < N && row0 < M && epi.dbg != 11;   // warp-uniform
            // residual of one 16-column step: 32 bytes of hi (+ lo) per row
            const bool res_wide = (eu.ld_res % 16 == 0) && ((((uintptr_t)eu.res_hi) | ((uintptr_t)eu.res_lo)) & 31) == 0;
            uint32_t rh[kRes ? 2 : 1][8], rl[kRes ? 2 : 1][8];
            auto load_res = [&](int c, uint32_t (&h)[8], uint32_t (&l)[8]) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { h[j] = 0u; l[j] = 0u; }
                const int col = u.n0 + cl0 + c * 16;
                if (row_ok && eu.res_hi && cl0 + c * 16 < u.w && col < N) {
                    const size_t ro = (size_t)rowp * eu.ld_res + col;
                    if (res_wide && col + 16 <= N) {
                        ld_global_nc_256(eu.res_hi + ro, h);
                        if (eu.res_lo) ld_global_nc_256(eu.res_lo + ro, l);
                    } else {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            if (col + i * 8 < N) {
                                const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(eu.res_hi + ro + i * 8));
                                h[i * 4] = a4.x; h[i * 4 + 1] = a4.y; h[i * 4 + 2] = a4.z; h[i * 4 + 3] = a4.w;
                                if (eu.res_lo) {
                                    const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(eu.res_lo + ro + i * 8));
                                    l[i * 4] = b4.x; l[i * 4 + 1] = b4.y; l[i * 4 + 2] = b4.z; l[i * 4 + 3] = b4.w;
                                }
                            }
                    }
                }
            };
            if (kRes && active) {
                // the unit's whole row segment -> L2 now (the layer input was written several launches ago and has
                // mostly left L2): the later steps' loads, issued only one step ahead, then miss no further than L2
                if (row_ok && eu.res_hi && epi.dbg != 13) {
                    const size_t ro = (size_t)rowp * eu.ld_res + u.n0 + cl0;
                    int wcols = u.w - cl0 < GW ? u.w - cl0 : GW;
                    if (u.n0 + cl0 + wcols > N) wcols = N - u.n0 - cl0;
                    prefetch_l2(eu.res_hi + ro);
                    prefetch_l2(eu.res_hi + ro + wcols - 1);
                    if (eu.res_lo) { prefetch_l2(eu.res_lo + ro); prefetch_l2(eu.res_lo + ro + wcols - 1); }
                }
                load_res(0, rh[0], rl[0]);   // overlaps the main loop of this tile
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            tr.ev(7);   // accumulators of the unit ready
            if (active) {
                const uint32_t t_row = t_lane + (uint32_t)(acc * TBN + cl0);
                // without a residual the accumulator registers are double buffered (the next step's tcgen05.ld travels
                // while this step is computed); with one, the registers go to the residual's one-step lookahead instead
                constexpr int RB = kRes ? 1 : 2;
                uint32_t r[RB][16];
                tc_ld16(t_row, r[0]);
#pragma unroll
                for (int c = 0; c < STEPS; ++c) {
                    const int cl = cl0 + c * 16;
                    const int col0 = u.n0 + cl;
                    if (cl >= u.w || col0 >= N) break;      // warp-uniform
                    float4 bv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (eu.bias && col0 + i * 4 < N) bv[i] = __ldg(reinterpret_cast<const float4*>(eu.bias + col0 + i * 4));
                    }
                    tc_wait_ld();
                    const bool more = c + 1 < STEPS && cl + 16 < u.w && col0 + 16 < N;
                    if (!kRes && more) tc_ld16(t_row + (uint32_t)((c + 1) * 16), r[(c + 1) % RB]);
                    if (kRes && more) load_res(c + 1, rh[(c + 1) & 1], rl[(c + 1) & 1]);   // one step ahead
                    tr.ev(20 + c);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i * 4 + 0] = __uint_as_float(r[c % RB][i * 4 + 0]) + bv[i].x;
                        v[i * 4 + 1] = __uint_as_float(r[c % RB][i * 4 + 1]) + bv[i].y;
                        v[i * 4 + 2] = __uint_as_float(r[c % RB][i * 4 + 2]) + bv[i].z;
                        v[i * 4 + 3] = __uint_as_float(r[c % RB][i * 4 + 3]) + bv[i].w;
                    }
                    if (kRes && more) tc_ld16(t_row + (uint32_t)((c + 1) * 16), r[0]);   // r[0] is free again
                    if (eu.act == NAVC_ACT_GELU_NEW) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = act_apply_fast(v[j], NAVC_ACT_GELU_NEW);
                    } else if (eu.act != NAVC_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = act_apply_fast(v[j], eu.act);
                    }
                    if constexpr (kRes) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            v[q * 2 + 0] += __uint_as_float(rh[c & 1][q] << 16) + __uint_as_float(rl[c & 1][q] << 16);
                            v[q * 2 + 1] += __uint_as_float(rh[c & 1][q] & 0xffff0000u) + __uint_as_float(rl[c & 1][q] & 0xffff0000u);
                        }
                    }
                    if (rz) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = 0.f;
                    }
                    uint32_t hw[8], lw[8];
                    if (eu.out_lo) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            hw[j] = *reinterpret_cast<const uint32_t*>(&hb);
                        }
                    }
                    tr.ev(30 + c);
                    if (epi.dbg == 12) continue;   // profiling aid: no stores
                    if (kChain && !u.prob) {
                        // chained, problem 0: generic-proxy stores (32 bytes of hi and of lo per row), see the publication below
                        if (row_ok) {
                            const size_t oo = (size_t)rowp * eu.ld_out + col0;
                            st_global_256(eu.out_hi + oo, hw);
                            if (eu.out_lo) st_global_256(eu.out_lo + oo, lw);
                        }
                        tr.ev(50 + c);
                        continue;
                    }
                    // hi box: its previous store (one bulk group before the most recent one) must have read the staging box
                    if (lane == 0) {
                        if (eu.out_lo) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                    *reinterpret_cast<uint4*>(stg + lane * 32) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(stg + lane * 32 + 16) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(mo_hi_u, stg_s, col0, row0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (eu.out_lo) {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the previous lo box
                        __syncwarp();
                        *reinterpret_cast<uint4*>(stg + 1024 + lane * 32) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        *reinterpret_cast<uint4*>(stg + 1024 + lane * 32 + 16) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(mo_lo_u, stg_s + 1024u, col0, row0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                    tr.ev(50 + c);
                }
            }
            tr.ev(8);   // epilogue of the unit issued
            arrive_drained(acc);   // (before the publication below: the accumulator stage is free once it has been read)
            if (kChain && !u.prob && u.valid) {
                // problem 0's tiles leave through ordinary 256-bit stores (above), not TMA boxes (only the issuing thread can wait
                // for a bulk store to COMPLETE): all lanes' stores -> warp barrier -> arrival (release) on the publisher's barrier
                __syncwarp();
                if (lane == 0) mbar_arrive(pub_bar(chain_i0));
                ++chain_i0;
            }
            if (++acc == G2_ACC) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging stays valid until read
        tr.ev(9);       // stores drained
    }

    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();   // no CTA leaves (or frees TMEM) while the pair's MMAs / barrier arrivals can still reach it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        if (g2_trace_buf && lane == 0)      // timeline: kernel exit (slot 62)
        {
            g2_trace_buf[((size_t)blockIdx.x * 3) * 64 + 62] = (61ull << 56) | ((unsigned long long)clock64() & 0x00ffffffffffffffull);
            unsigned long long gt;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
            g2_trace_buf[((size_t)blockIdx.x * 3 + 1) * 64 + 62] = gt;
        }
    }
}

template <bool kX3, int TBN, int kCl, bool kRes>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const __grid_constant__ CUtensorMap map_p_hi, const __grid_constant__ CUtensorMap map_p_lo,
                const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                int M_max, int N, int K, EpiParams epi) {
    gemm2_tc_body<kX3, TBN, kCl, kRes, false>(map_a_hi, map_a_lo, map_b_hi, map_b_lo, map_p_hi, map_p_lo, map_o_hi, map_o_lo, M_max, N, K,
                                              epi, nullptr);
}

// two chained problems (G2Chain above): single CTAs, 128-wide tiles, the residual-capable epilogue for both
template <bool kX3>
__global__ void __launch_bounds__(G2_THREADS + 32, 1)
gemm2_chain_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const __grid_constant__ CUtensorMap map_p_hi, const __grid_constant__ CUtensorMap map_p_lo,
                   const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                   int M_max, int N, int K, EpiParams epi, const __grid_constant__ G2Chain chain) {
    gemm2_tc_body<kX3, 128, 1, true, true>(map_a_hi, map_a_lo, map_b_hi, map_b_lo, map_p_hi, map_p_lo, map_o_hi, map_o_lo, M_max, N, K,
                                           epi, &chain);
}

// ---- host side ------------------------------------------------------------------------------------
int tc_make_store_map16(CUtensorMap* map, const uint16_t* ptr, int rows, int cols, int ld);   // gemm_tc.cu

}  // namespace navc
// Debug: install (or remove, buf == NULL) the in-kernel timeline buffer: 3 x 64 uint64 per CTA of the grid.
extern "C" int navc_debug_trace(void* buf) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(navc::g2_trace_buf, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
namespace navc {

static bool g_g2_ready = false;
static int* g_chain_cnt = nullptr;           // per-row-block completion counts of a chained launch (cleared per launch)
constexpr int kChainCntInts = 8192;          // row blocks: up to 1M rows
static int g2_init() {
    if (g_g2_ready) return 0;
#define NAVC_G2_ATTR(X3, BN, CL) \
    NAVC_CUDA(cudaFuncSetAttribute(gemm2_tc_kernel<X3, BN, CL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<X3, BN, CL>::kSmemBytes)); \
    NAVC_CUDA(cudaFuncSetAttribute(gemm2_tc_kernel<X3, BN, CL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<X3, BN, CL>::kSmemBytes))
    NAVC_G2_ATTR(false, 128, 1); NAVC_G2_ATTR(false, 256, 1); NAVC_G2_ATTR(true, 128, 1); NAVC_G2_ATTR(true, 256, 1);
    NAVC_G2_ATTR(false, 128, 2); NAVC_G2_ATTR(false, 256, 2); NAVC_G2_ATTR(true, 128, 2); NAVC_G2_ATTR(true, 256, 2);
#undef NAVC_G2_ATTR
    NAVC_CUDA(cudaFuncSetAttribute(gemm2_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<true, 128, 1>::kSmemBytes));
    NAVC_CUDA(cudaFuncSetAttribute(gemm2_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<false, 128, 1>::kSmemBytes));
    NAVC_CUDA(cudaMalloc(&g_chain_cnt, kChainCntInts * sizeof(int)));
    g_g2_ready = true;
    return 0;
}

// cost of the persistent schedule the kernel derives, for choosing TBN (units: the MMA time of a 128-column k-block)
static float g2_cost(int m_blocks, int N, int tbn, int cl, int sms, bool x3) {
    const int units = ((m_blocks + cl - 1) / cl) * ((N + tbn - 1) / tbn);
    int G = sms / cl;
    if (G > units) G = units;
    const int full = units / G, rem = units % G;
    int parts = 1;
    if (rem > 0) {
        const int q = G / rem;
        parts = q >= 4 ? 4 : (q >= 2 ? 2 : 1);
        if (parts > tbn / 64) parts = tbn / 64;
    }
    // bytes entering an SM per k-block: the A rows + (pair: half of) the B rows; a k-block takes the longer of its
    // MMAs (768 cycles split-bf16 / 256 plain per 128 columns) and that stream at ~40 B/clk
    auto kblock = [&](float cols) {
        const float mma = cols / 128.0f, stream = (128.0f + cols / cl) / (x3 ? 120.0f : 80.0f);
        return mma > stream ? mma : stream;
    };
    return full * kblock((float)tbn) + (rem > 0 ? kblock((float)tbn / parts) : 0.0f);
}

template <bool kX3, int TBN, int kCl>
static int g2_launch(const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w_hi, const uint16_t* w_lo, int ldw,
                     int M, int N, int K, const EpiParams& epi, int sms, cudaStream_t st) {
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mp_hi, mp_lo, mo_hi, mo_lo;
    const int bbox = TBN / kCl;
    if (tc_make_map(&ma_hi, x_hi, M, K, ldx, G2_BM) || tc_make_map(&mb_hi, w_hi, N, K, ldw, bbox) ||
        tc_make_map(&mp_hi, w_hi, N, K, ldw, G2_PBOX)) return 1;
    ma_lo = ma_hi;
    mb_lo = mb_hi;
    mp_lo = mp_hi;
    if (kX3) {
        if (tc_make_map(&ma_lo, x_lo, M, K, ldx, G2_BM) || tc_make_map(&mb_lo, w_lo, N, K, ldw, bbox) ||
            tc_make_map(&mp_lo, w_lo, N, K, ldw, G2_PBOX)) return 1;
    }
    if (tc_make_store_map16(&mo_hi, epi.out_hi, M, N, epi.ld_out)) return 1;
    mo_lo = mo_hi;
    if (epi.out_lo && tc_make_store_map16(&mo_lo, epi.out_lo, M, N, epi.ld_out)) return 1;
    const int m_blocks = (M + G2_BM - 1) / G2_BM;
    const int units = ((m_blocks + kCl - 1) / kCl) * ((N + TBN - 1) / TBN);
    int clusters = sms / kCl;
    if (clusters > units * (TBN / 64)) clusters = units * (TBN / 64);   // small problems: room for the column parts
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCl);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = G2Cfg<kX3, TBN, kCl>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = kCl;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = kCl > 1 ? 2 : 1;
    if (epi.res_hi)
        NAVC_CUDA(cudaLaunchKernelEx(&cfg, gemm2_tc_kernel<kX3, TBN, kCl, true>, ma_hi, ma_lo, mb_hi, mb_lo, mp_hi, mp_lo, mo_hi, mo_lo, M, N, K, epi));
    else
        NAVC_CUDA(cudaLaunchKernelEx(&cfg, gemm2_tc_kernel<kX3, TBN, kCl, false>, ma_hi, ma_lo, mb_hi, mb_lo, mp_hi, mp_lo, mo_hi, mo_lo, M, N, K, epi));
    return check_launch("navc_linear_tc (gemm2)");
}

static int g_g2_on = -1;   // 0 off, 1 on (split-bf16 mode), 2 forced (every mode)
bool g2_enabled() {
    if (g_g2_on < 0) {
        // NAVC_GEMM2=0: the first-generation kernel (gemm_tc.cu) for A/B runs; NAVC_GEMM2=force: also in plain bf16
        const char* e = getenv("NAVC_GEMM2");
        g_g2_on = (e && (e[0] == '0' || e[0] == 'n')) ? 0 : ((e && e[0] == 'f') ? 2 : 1);
    }
    return g_g2_on != 0;
}
bool g2_forced() { return g2_enabled() && g_g2_on == 2; }

// Pair-epilogue GEMM (bf16 hi(/lo) outputs only): called by navc_linear_tc once the arguments are validated.
int g2_linear(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w_hi, const uint16_t* w_lo, int ldw,
              int M, int N, int K, const EpiParams& epi, cudaStream_t st) {
    NAVC_REQUIRE(tc_ready(), "navc_linear_tc: navc_init() has not been called");
    if (g2_init()) return 2;
    int sms = navc_sm_count();
    if (sms <= 0) sms = 148;
    // rows the launch is expected to compute (packed rows: the device-side count; the host passes its estimate)
    const int m_est = (epi.m_hint > 0 && epi.m_hint < M) ? epi.m_hint : M;
    const int mbe = (m_est + G2_BM - 1) / G2_BM;
    const char* ecl = getenv("NAVC_GEMM2_CLUSTER");
    // out / query projections (N <= 512, K <= 512: 8 k-blocks per tile, 1-2 tiles per CTA): no steady state to win back a
    // pair's start-up or a 256-wide tile's last, un-overlapped epilogue -- single CTAs, 128-wide tiles (isolated, M = 10 553:
    // so 24.0 us against 29.6 (256 wide), 31.3 (pairs) and 27.4 (first-generation kernel); cq 21.1 / 22.0 / 23.1 / 21.9)
    const bool small = N <= 512 && K <= 512;
    int cl = (mbe >= 8 && sms % 2 == 0 && !small) ? 2 : 1;     // tiny M (AR beam steps): more, smaller tiles instead of CTA pairs
    if (ecl) cl = ecl[0] == '1' ? 1 : 2;
    const bool x3 = mode == NAVC_TC_BF16X3;
    int tbn = (N > 128 && !(small && !ecl) && g2_cost(mbe, N, 256, cl, sms, x3) <= g2_cost(mbe, N, 128, cl, sms, x3)) ? 256 : 128;
    if (epi.dbg == 128 || epi.dbg == 256) tbn = epi.dbg;   // profiling aid: force a tile width
#define NAVC_G2_GO(X3, BN, CL) return g2_launch<X3, BN, CL>(x_hi, x_lo, ldx, w_hi, w_lo, ldw, M, N, K, epi, sms, st)
    if (x3) {
        if (tbn == 256) { if (cl == 2) NAVC_G2_GO(true, 256, 2); NAVC_G2_GO(true, 256, 1); }
        if (cl == 2) NAVC_G2_GO(true, 128, 2);
        NAVC_G2_GO(true, 128, 1);
    }
    if (tbn == 256) { if (cl == 2) NAVC_G2_GO(false, 256, 2); NAVC_G2_GO(false, 256, 1); }
    if (cl == 2) NAVC_G2_GO(false, 128, 2);
    NAVC_G2_GO(false, 128, 1);
#undef NAVC_G2_GO
}

// Y0 = epi0(X W0^T) then Y1 = epi1(Y0 W1^T) in one launch (gemm2_chain_kernel): N == K (square layers), 128-wide tiles.
int g2_chain(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w0_hi, const uint16_t* w0_lo, int ldw0,
             const EpiParams& e0, const uint16_t* w1_hi, const uint16_t* w1_lo, int ldw1, const EpiParams& e1, int M, int N, int K,
             cudaStream_t st) {
    NAVC_REQUIRE(tc_ready(), "navc_linear_chain_tc: navc_init() has not been called");
    if (g2_init()) return 2;
    const bool x3 = mode == NAVC_TC_BF16X3;
    NAVC_REQUIRE(N == K && N % 128 == 0 && M > 0, "navc_linear_chain_tc: needs N == K, N %% 128 == 0 (N=%d K=%d)", N, K);
    NAVC_REQUIRE(e0.out_hi && e1.out_hi && (!x3 || (e0.out_lo && e1.out_lo)) && !e0.out_f32 && !e1.out_f32 && !e0.residual && !e1.residual &&
                     e0.m_dev == e1.m_dev, "navc_linear_chain_tc: bf16 hi/lo outputs only, same device-side row count");
    NAVC_REQUIRE(e0.ld_out % 16 == 0 && ((((uintptr_t)e0.out_hi) | ((uintptr_t)e0.out_lo)) & 31) == 0,
                 "navc_linear_chain_tc: y0 needs 32-byte aligned rows (256-bit stores)");
    int sms = navc_sm_count();
    if (sms <= 0) sms = 148;
    const int m_blocks = (M + G2_BM - 1) / G2_BM, n_blocks = N / 128;
    NAVC_REQUIRE(m_blocks <= kChainCntInts, "navc_linear_chain_tc: too many rows");
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mp_hi, mp_lo, mo_hi, mo_lo;
    G2Chain ch;
    if (tc_make_map(&ma_hi, x_hi, M, K, ldx, G2_BM) || tc_make_map(&mb_hi, w0_hi, N, K, ldw0, 128) || tc_make_map(&mp_hi, w0_hi, N, K, ldw0, G2_PBOX) ||
        tc_make_map(&ch.a_hi, e0.out_hi, M, N, e0.ld_out, G2_BM) || tc_make_map(&ch.b_hi, w1_hi, N, K, ldw1, 128) ||
        tc_make_map(&ch.p_hi, w1_hi, N, K, ldw1, G2_PBOX)) return 1;
    ma_lo = ma_hi; mb_lo = mb_hi; mp_lo = mp_hi; ch.a_lo = ch.a_hi; ch.b_lo = ch.b_hi; ch.p_lo = ch.p_hi;
    if (x3) {
        if (tc_make_map(&ma_lo, x_lo, M, K, ldx, G2_BM) || tc_make_map(&mb_lo, w0_lo, N, K, ldw0, 128) || tc_make_map(&mp_lo, w0_lo, N, K, ldw0, G2_PBOX) ||
            tc_make_map(&ch.a_lo, e0.out_lo, M, N, e0.ld_out, G2_BM) || tc_make_map(&ch.b_lo, w1_lo, N, K, ldw1, 128) ||
            tc_make_map(&ch.p_lo, w1_lo, N, K, ldw1, G2_PBOX)) return 1;
    }
    if (tc_make_store_map16(&mo_hi, e0.out_hi, M, N, e0.ld_out) || tc_make_store_map16(&ch.o_hi, e1.out_hi, M, N, e1.ld_out)) return 1;
    mo_lo = mo_hi; ch.o_lo = ch.o_hi;
    if (e0.out_lo && tc_make_store_map16(&mo_lo, e0.out_lo, M, N, e0.ld_out)) return 1;
    if (e1.out_lo && tc_make_store_map16(&ch.o_lo, e1.out_lo, M, N, e1.ld_out)) return 1;
    ch.epi = e1;
    ch.cnt = g_chain_cnt;
    NAVC_CUDA(cudaMemsetAsync(g_chain_cnt, 0, (size_t)m_blocks * sizeof(int), st));
    const int units = 2 * m_blocks * n_blocks;
    int grid = sms;
    if (grid > units) grid = units;   // (all CTAs co-resident: one per SM)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(G2_THREADS + 32);   // + the publisher warp
    cfg.dynamicSmemBytes = x3 ? G2Cfg<true, 128, 1>::kSmemBytes : G2Cfg<false, 128, 1>::kSmemBytes;
    cfg.stream = st;
    if (x3)
        NAVC_CUDA(cudaLaunchKernelEx(&cfg, gemm2_chain_kernel<true>, ma_hi, ma_lo, mb_hi, mb_lo, mp_hi, mp_lo, mo_hi, mo_lo, M, N, K, e0, ch));
    else
        NAVC_CUDA(cudaLaunchKernelEx(&cfg, gemm2_chain_kernel<false>, ma_hi, ma_lo, mb_hi, mb_lo, mp_hi, mp_lo, mo_hi, mo_lo, M, N, K, e0, ch));
    return check_launch("navc_linear_chain_tc");
}

}  // namespace navc

// Two chained linear layers in one launch: y0 = epilogue0(x w0^T) (bias, activation, bf16 hi/lo residual, row mask), then
// y1 = epilogue1(y0 w1^T), bf16 hi (/ lo) outputs; N == K (square layers).  See G2Chain in this file.
extern "C" int navc_linear_chain_tc(int mode, const uint16_t* x_hi, const uint16_t* x_lo, int ldx, const uint16_t* w0_hi,
                                    const uint16_t* w0_lo, int ldw0, const navc_epilogue_t* e0, const uint16_t* w1_hi,
                                    const uint16_t* w1_lo, int ldw1, const navc_epilogue_t* e1, int M, int N, int K, void* stream) {
    using namespace navc;
    NAVC_REQUIRE(e0 && e1 && x_hi && w0_hi && w1_hi, "navc_linear_chain_tc: null pointer");
    NAVC_REQUIRE(mode == NAVC_TC_BF16 || (mode == NAVC_TC_BF16X3 && x_lo && w0_lo && w1_lo), "navc_linear_chain_tc: bad mode / missing lo operands");
    NAVC_REQUIRE(ldx % 8 == 0 && ldw0 % 8 == 0 && ldw1 % 8 == 0 && e0->ld_out % 8 == 0 && e1->ld_out % 8 == 0,
                 "navc_linear_chain_tc: leading dimensions must be multiples of 8");
    const EpiParams p0 = to_params(e0), p1 = to_params(e1);
    return g2_chain(mode, x_hi, x_lo, ldx, w0_hi, w0_lo, ldw0, p0, w1_hi, w1_lo, ldw1, p1, M, N, K, as_stream(stream));
}
