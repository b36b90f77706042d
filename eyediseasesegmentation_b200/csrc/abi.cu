// Library-level entry points: version, error string, device probe, one-time init.
#include "common.cuh"
#include <cstring>

namespace eds {

static thread_local char g_err[512] = "";

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
        return EDS_ERR_CUDA;
    }
    return EDS_OK;
}

int igemm_init();  // conv_igemm_sm100.cu

}  // namespace eds

extern "C" int eds_version(void) { return 100; }  // 0.1.0

extern "C" const char* eds_last_error(void) { return eds::g_err; }

extern "C" int eds_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 0;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

extern "C" int eds_init(void) {
    if (!eds_device_ok()) {
        eds::set_error("eds_init: no sm_100 CUDA device visible; this library has no CPU fallback");
        return EDS_ERR_UNSUPPORTED;
    }
    return eds::igemm_init();
}
