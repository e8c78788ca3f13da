// fp32 CUDA-core GEMM  Y = epilogue(X[M,K] * W[N,K]^T)  -- the reference-exact precision mode.
// 128x128x16 tiles, 256 threads, 8x8 micro-tile, double-buffered shared memory.
#include "common.cuh"

namespace navc {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct VocabEpi {
    const float* bias;
    float* part_max;
    float* part_sum;
    int32_t* part_idx;
    const int64_t* target;
    float* target_logit;
    int n_tiles;
};

template <bool kVocab>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(const float* __restrict__ A, int lda,
                                                      const float* __restrict__ W, int ldw, int M, int N,
                                                      int K, EpiParams epi, VocabEpi vep) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lrow = tid & 127;        // row of the tile this thread loads
    const int lk = (tid >> 7) * 4;     // k offset 0 or 4 (and +8 for the second load)
    const int ty = tid >> 4, tx = tid & 15;

    const float* a_ptr = A + (size_t)min(m0 + lrow, M - 1) * lda;
    const float* w_ptr = W + (size_t)min(n0 + lrow, N - 1) * ldw;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ra[i] = *reinterpret_cast<const float4*>(a_ptr + k0 + lk + 8 * i);
            rb[i] = *reinterpret_cast<const float4*>(w_ptr + k0 + lk + 8 * i);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int k = lk + 8 * i;
            As[buf][k + 0][lrow] = ra[i].x; As[buf][k + 1][lrow] = ra[i].y;
            As[buf][k + 2][lrow] = ra[i].z; As[buf][k + 3][lrow] = ra[i].w;
            Bs[buf][k + 0][lrow] = rb[i].x; Bs[buf][k + 1][lrow] = rb[i].y;
            Bs[buf][k + 2][lrow] = rb[i].z; Bs[buf][k + 3][lrow] = rb[i].w;
        }
    };

    const int nk = K / BK;  // K % 16 == 0 enforced by the host wrapper
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) gload((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    if constexpr (!kVocab) {
        const bool vec = (epi.ld_out % 4 == 0) && (N % 4 == 0) && (!epi.residual || epi.ld_res % 4 == 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (row >= M) continue;
            bool rz = epi.row_tokens ? (epi.row_tokens[row] == NAVC_PAD) : false;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int col = n0 + h * 64 + tx * 4;
                if (col >= N) continue;
                if (vec) {
                    epi_store4(epi, row, col, make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1],
                                                          acc[i][h * 4 + 2], acc[i][h * 4 + 3]), rz);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < N) epi_store1(epi, row, col + j, acc[i][h * 4 + j], rz);
                }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            int64_t tgt = (vep.target && row < M) ? vep.target[row] : -1;
            SoftPart p;
            p.m = -INFINITY; p.s = 0.f; p.i = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (col < N) {
                    float v = acc[i][j] + (vep.bias ? vep.bias[col] : 0.f);
                    acc[i][j] = v;
                    if (v > p.m) { p.m = v; p.i = col; }
                    if (col == tgt) vep.target_logit[row] = v;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (col < N) p.s += expf(acc[i][j] - p.m);
            }
            // reduce over the 16 threads (tx) that share this row: lanes differ in the low 4 bits
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                SoftPart q;
                q.m = __shfl_xor_sync(0xffffffffu, p.m, o);
                q.s = __shfl_xor_sync(0xffffffffu, p.s, o);
                q.i = __shfl_xor_sync(0xffffffffu, p.i, o);
                p = soft_combine(p, q);
            }
            if (tx == 0 && row < M) {
                size_t o = (size_t)row * vep.n_tiles + blockIdx.x;
                vep.part_max[o] = p.m; vep.part_sum[o] = p.s; vep.part_idx[o] = p.i;
            }
        }
    }
}

}  // namespace navc

using namespace navc;

extern "C" int navc_linear_f32(const float* x, int ldx, const float* w, int ldw, int M, int N, int K,
                               const navc_epilogue_t* e, void* stream) {
    NAVC_REQUIRE(x && w && e, "navc_linear_f32: null pointer");
    NAVC_REQUIRE(M > 0 && N > 0 && K > 0, "navc_linear_f32: bad shape M=%d N=%d K=%d", M, N, K);
    NAVC_REQUIRE(K % BK == 0 && ldx % 4 == 0 && ldw % 4 == 0,
                 "navc_linear_f32: need K%%16==0, ldx%%4==0, ldw%%4==0 (K=%d ldx=%d ldw=%d)", K, ldx, ldw);
    NAVC_REQUIRE(e->out_f32 || e->out_hi, "navc_linear_f32: no output");
    NAVC_REQUIRE(!(e->accumulate || e->split_k > 1) || e->out_f32, "navc_linear_f32: accumulate needs out_f32");
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    NAVC_REQUIRE(grid.y <= 65535, "navc_linear_f32: M too large");
    VocabEpi v = {};
    EpiParams p = to_params(e);
    p.split_k = 1;  // the CUDA-core path runs the whole K loop per tile (accumulate still honoured)
    gemm_f32_kernel<false><<<grid, NT, 0, as_stream(stream)>>>(x, ldx, w, ldw, M, N, K, p, v);
    return check_launch("navc_linear_f32");
}

extern "C" int navc_vocab_partials_f32(const float* h, int ldh, const float* w, int ldw, const float* bias,
                                       int M, int V, int K, float* part_max, float* part_sum,
                                       int32_t* part_idx, const int64_t* target, float* target_logit,
                                       void* stream) {
    NAVC_REQUIRE(h && w && part_max && part_sum && part_idx, "navc_vocab_partials_f32: null pointer");
    NAVC_REQUIRE(K % BK == 0 && ldh % 4 == 0 && ldw % 4 == 0, "navc_vocab_partials_f32: alignment");
    NAVC_REQUIRE(!target || target_logit, "navc_vocab_partials_f32: target without target_logit");
    dim3 grid((V + BN - 1) / BN, (M + BM - 1) / BM);
    NAVC_REQUIRE(grid.y <= 65535, "navc_vocab_partials_f32: M too large");
    VocabEpi v = {bias, part_max, part_sum, part_idx, target, target_logit, (int)grid.x};
    EpiParams e = {};
    gemm_f32_kernel<true><<<grid, NT, 0, as_stream(stream)>>>(h, ldh, w, ldw, M, V, K, e, v);
    return check_launch("navc_vocab_partials_f32");
}
