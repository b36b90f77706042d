// Shared helpers for the eds_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include <mutex>

#include "../../include/eds_b200.h"

namespace eds {

// ---- error plumbing -------------------------------------------------------
// eds_last_error() returns the thread-local message of the last failing call.
void set_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaGetLastError -> EDS_ERR_CUDA

#define EDS_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::eds::set_error(__VA_ARGS__);                       \
            return EDS_ERR_INVALID;                              \
        }                                                        \
    } while (0)

// ---- one-time-per-DEVICE initialisation ---------------------------------------
// cudaFuncSetAttribute (dynamic shared memory opt-in) and __constant__ uploads belong to a device's
// context, not to the process: a process that drives two GPUs must repeat them on the second one.
// `PerDevice once; once.run([](int dev) { ...; return EDS_OK; })` runs the body the first time it is
// reached with each current device (a failed body is retried on the next call).
constexpr int kMaxDevices = 64;
struct PerDevice {
    std::atomic<bool> done[kMaxDevices];
    std::mutex mu;
    PerDevice() { for (auto& d : done) d.store(false); }
    template <typename F> int run(F&& body) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
            set_error("cannot identify the current CUDA device (index %d)", dev);
            return EDS_ERR_CUDA;
        }
        if (done[dev].load(std::memory_order_acquire)) return EDS_OK;
        std::lock_guard<std::mutex> lock(mu);
        if (done[dev].load(std::memory_order_relaxed)) return EDS_OK;
        const int rc = body(dev);
        if (rc == EDS_OK) done[dev].store(true, std::memory_order_release);
        return rc;
    }
};
// dynamic shared-memory opt-in of one kernel, once per device
template <typename Kern> static inline int smem_opt_in(PerDevice& once, Kern kern, int bytes, const char* what) {
    return once.run([&](int) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) {
            set_error("%s: cannot opt in to %d B of shared memory: %s", what, bytes, cudaGetErrorString(e));
            return (int)EDS_ERR_CUDA;
        }
        return (int)EDS_OK;
    });
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- activation element access (bf16 or fp32 storage, fp32 math) ----------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 8 consecutive channels as one vector access (16 B for bf16, 32 B for fp32).
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 raw = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 raw;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = raw;
    }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// The same 8 channels as four float2 for the packed fp32 pipe of sm_100 (FFMA2 / FADD2 / FMUL2:
// two IEEE fp32 operations per instruction, results identical to the scalar forms).
template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float2 (&v)[4]) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p);
        const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(u[i] << 16), __uint_as_float(u[i] & 0xffff0000u));
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float2 (&v)[4]) {
        uint4 raw;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
        *reinterpret_cast<uint4*>(p) = raw;
    }
};
template <> struct V8<float> {
    static __device__ __forceinline__ void ld(const float* p, float2 (&v)[4]) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        const float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
        v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
    }
    static __device__ __forceinline__ void st(float* p, const float2 (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
    }
};
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ void ld8f(const float* p, float2 (&v)[4]) { V8<float>::ld(p, v); }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Dispatch a templated launcher on the activation dtype code.
#define EDS_DISPATCH_DTYPE(dtype, T, ...)                              \
    do {                                                               \
        if ((dtype) == EDS_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else if ((dtype) == EDS_F32) { using T = float; __VA_ARGS__; } \
        else { ::eds::set_error("unknown dtype code %d", (int)(dtype)); return EDS_ERR_INVALID; } \
    } while (0)

}  // namespace eds
