// Fixed cost of a launch on B200: empty kernels with the launch shapes of the hot kernels, 200 back-to-back launches captured
// in a CUDA graph, per-launch time from CUDA events.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launch_overhead launch_overhead.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }

template <int kTmem, int kBars>
__global__ void k_prologue(int* p) {
    extern __shared__ uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bars[32];
    if (kBars && threadIdx.x == 0) {
        for (int i = 0; i < kBars; ++i) {
            uint32_t a = (uint32_t)__cvta_generic_to_shared(&bars[i]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(1) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kTmem && threadIdx.x < 32) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(kTmem) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (kTmem && threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(kTmem) : "memory");
    }
    if (p && threadIdx.x == 9999) *p = smem[0];
}

struct Big { unsigned long long w[16]; };   // 128 bytes, like a CUtensorMap
__global__ void k_params(const __grid_constant__ Big a, const __grid_constant__ Big b, const __grid_constant__ Big c, const __grid_constant__ Big d,
                         const __grid_constant__ Big e, const __grid_constant__ Big f, const __grid_constant__ Big g, const __grid_constant__ Big h, int* p) {
    if (p && threadIdx.x == 9999) *p = (int)(a.w[0] + b.w[1] + c.w[2] + d.w[3] + e.w[4] + f.w[5] + g.w[6] + h.w[7]);
}

// every CTA allocates TMEM, spins for `cycles`, optionally writes `bytes` per thread to global memory, then leaves
template <int kWrite>
__global__ void k_busy(long long cycles, float* out, int* p) {
    extern __shared__ uint8_t smem[];
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (kWrite) {
        float4* o = reinterpret_cast<float4*>(out) + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kWrite;
        for (int i = 0; i < kWrite; ++i) o[i] = make_float4(1.f, 2.f, 3.f, (float)i);
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
    if (p && threadIdx.x == 9999) *p = smem[0];
}

template <typename F>
static int time_graph(const char* name, F launch) {
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    for (int i = 0; i < 3; ++i) launch(st);
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < 200; ++i) launch(st);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 5; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-64s %6.2f us per launch\n", name, ms * 1e3f / 1000.f);
    return 0;
}

static cudaError_t launch_ex(void (*k)(int*), int grid, int block, size_t smem, int cluster, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster > 1 ? 1 : 0;
    int* null = nullptr;
    return cudaLaunchKernelEx(&cfg, k, null);
}

int main() {
    const int big = 200 * 1024, max = 226 * 1024;
    CK(cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, max));
    CK(cudaFuncSetAttribute(k_prologue<512, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max));
    CK(cudaFuncSetAttribute(k_prologue<512, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, max));
    CK(cudaFuncSetAttribute(k_prologue<0, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, max));
    time_graph("empty, 148 x 128 threads, no smem", [&](cudaStream_t s) { launch_ex(k_empty, 148, 128, 0, 1, s); });
    time_graph("empty, 148 x 576 threads, no smem", [&](cudaStream_t s) { launch_ex(k_empty, 148, 576, 0, 1, s); });
    time_graph("empty, 148 x 576 threads, 200 KB smem", [&](cudaStream_t s) { launch_ex(k_empty, 148, 576, big, 1, s); });
    time_graph("empty, 148 x 576 threads, 200 KB smem, clusters of 2", [&](cudaStream_t s) { launch_ex(k_empty, 148, 576, big, 2, s); });
    time_graph("empty, 148 x 320 threads, 226 KB smem", [&](cudaStream_t s) { launch_ex(k_empty, 148, 320, max, 1, s); });
    time_graph("24 mbarriers + __syncthreads, 576 threads, 200 KB", [&](cudaStream_t s) { launch_ex(k_prologue<0, 24>, 148, 576, big, 1, s); });
    time_graph("+ TMEM alloc / dealloc 512 columns", [&](cudaStream_t s) { launch_ex(k_prologue<512, 24>, 148, 576, big, 1, s); });
    time_graph("+ TMEM alloc / dealloc 512 columns, clusters of 2", [&](cudaStream_t s) { launch_ex(k_prologue<512, 24>, 148, 576, big, 2, s); });
    time_graph("alternating 200 KB / no-smem kernels (carve-out switches)", [&](cudaStream_t s) {
        launch_ex(k_empty, 148, 576, big, 1, s); launch_ex(k_empty, 148, 256, 0, 1, s); });
    {
        Big z = {};
        int* null = nullptr;
        time_graph("empty, 1 KB of __grid_constant__ parameters, 576 threads, 200 KB", [&](cudaStream_t s) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = 0; cfg.stream = s;
            cudaLaunchKernelEx(&cfg, k_params, z, z, z, z, z, z, z, z, null); });
    }
    {
        float* out; int* null = nullptr;
        CK(cudaMalloc(&out, (size_t)148 * 576 * 16 * 16));
        CK(cudaFuncSetAttribute(k_busy<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max));
        CK(cudaFuncSetAttribute(k_busy<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max));
        const long long cyc = 19000;   // ~10 us at 1.9 GHz
        time_graph("148 CTAs busy 19000 clk (~10 us), TMEM, 200 KB smem", [&](cudaStream_t s) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = big; cfg.stream = s;
            cudaLaunchKernelEx(&cfg, k_busy<0>, cyc, out, null); });
        time_graph("  ... + 256 B written per thread before exit (21 MB per launch)", [&](cudaStream_t s) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = big; cfg.stream = s;
            cudaLaunchKernelEx(&cfg, k_busy<16>, cyc, out, null); });
        time_graph("  ... 8 CTAs only", [&](cudaStream_t s) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(8); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = big; cfg.stream = s;
            cudaLaunchKernelEx(&cfg, k_busy<0>, cyc, out, null); });
    }
    return 0;
}
