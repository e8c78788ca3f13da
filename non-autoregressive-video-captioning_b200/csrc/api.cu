// Library-level entry points: version, error reporting, one-time init.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace navc {

static thread_local char g_err[512] = "";
static int g_sm_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 3;
    }
    return 0;
}

int tc_init();  // gemm_tc.cu

}  // namespace navc

using namespace navc;

extern "C" int navc_version(void) { return NAVC_VERSION; }

extern "C" const char* navc_last_error(void) { return g_err; }

extern "C" int navc_init(int device) {
    NAVC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NAVC_CUDA(cudaGetDeviceProperties(&prop, device));
    NAVC_REQUIRE(prop.major == 10, "navc_init: libnavc is built for sm_100a only, device %d is sm_%d%d", device,
                 prop.major, prop.minor);
    g_sm_count = prop.multiProcessorCount;
    return tc_init();
}

extern "C" int navc_sm_count(void) {
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sm_count;
}
