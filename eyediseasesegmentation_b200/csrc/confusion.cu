// Per-image confusion counts of two binarised uint8 masks (the "next" row 8f-1 of the scope table).
//
// Reference: src/main/stat_result.py:30-57 (and stat_result_vessel.py): both masks are opened as 'L',
// thresholded with `x > 50` and reduced with numpy to true_p = sum(gt & pred), actual_p = sum(gt),
// pred_p = sum(pred); every metric of the CSVs (SN, PPV, SP, IoU, Dice) follows from those three integers
// and the pixel count.  One HBM-bound pass, 2 bytes per pixel; 16 pixels per 128-bit load, byte-wise
// compares with the SIMD-in-a-word video instructions, popcount, one 64-bit atomic per CTA and counter.
#include "common.cuh"
#include <algorithm>

namespace eds {

constexpr int kConfThreads = 256;

__device__ __forceinline__ uint32_t gt_mask4(uint32_t v, uint32_t thr4) { return __vcmpgtu4(v, thr4); }

// grid = (chunks, n_images); counts[img][3] = {true_p, actual_p, pred_p}
__global__ void __launch_bounds__(kConfThreads)
confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, int64_t n_px, int thr_pred,
                 int thr_gt, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_red[3][kConfThreads / 32];
    const int img = blockIdx.y;
    const uint8_t* p = pred + (int64_t)img * n_px;
    const uint8_t* g = gt + (int64_t)img * n_px;
    const uint32_t tp4 = (uint32_t)thr_pred * 0x01010101u, tg4 = (uint32_t)thr_gt * 0x01010101u;
    unsigned int c_tp = 0, c_ap = 0, c_pp = 0;
    const bool vec_ok = ((((uintptr_t)p) | ((uintptr_t)g)) & 15) == 0;
    const int64_t n_vec = vec_ok ? n_px / 16 : 0;
    const int64_t stride = (int64_t)gridDim.x * kConfThreads;
    for (int64_t i = (int64_t)blockIdx.x * kConfThreads + threadIdx.x; i < n_vec; i += stride) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p) + i);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(g) + i);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t mp = gt_mask4(av[q], tp4), mg = gt_mask4(bv[q], tg4);
            c_pp += __popc(mp);
            c_ap += __popc(mg);
            c_tp += __popc(mp & mg);
        }
    }
    // every mask byte is 0xFF or 0x00: 8 set bits per pixel
    c_tp >>= 3; c_ap >>= 3; c_pp >>= 3;
    for (int64_t i = n_vec * 16 + (int64_t)blockIdx.x * kConfThreads + threadIdx.x; i < n_px; i += stride) {
        const bool bp = p[i] > thr_pred, bg = g[i] > thr_gt;
        c_pp += bp;
        c_ap += bg;
        c_tp += bp && bg;
    }
    unsigned int v[3] = {c_tp, c_ap, c_pp};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int w = 0; w < kConfThreads / 32; ++w) t += s_red[threadIdx.x][w];
        if (t) atomicAdd(counts + (int64_t)img * 3 + threadIdx.x, t);
    }
}

}  // namespace eds

using namespace eds;

extern "C" int eds_confusion_u8(const uint8_t* pred, const uint8_t* gt, int64_t n_pixels, int n_images, int thr_pred,
                                int thr_gt, uint64_t* counts, void* stream) {
    EDS_REQUIRE(pred && gt && counts, "confusion: null pointer");
    EDS_REQUIRE(n_pixels > 0 && n_images > 0 && n_images <= 65535, "confusion: bad sizes (n_pixels=%lld n_images=%d)",
                (long long)n_pixels, n_images);
    EDS_REQUIRE(thr_pred >= 0 && thr_pred <= 255 && thr_gt >= 0 && thr_gt <= 255, "confusion: thresholds are bytes");
    // a thread's 32-bit partial counts 8 bits per pixel: keep its share below 2^29 pixels
    int chunks = (int)std::min<int64_t>(148 * 8, (n_pixels / 16 + kConfThreads - 1) / kConfThreads);
    if (chunks < 1) chunks = 1;
    EDS_REQUIRE(n_pixels / ((int64_t)chunks * kConfThreads) < (1ll << 28), "confusion: image too large for one launch");
    dim3 grid(chunks, n_images);
    confusion_kernel<<<grid, kConfThreads, 0, as_stream(stream)>>>(pred, gt, n_pixels, thr_pred, thr_gt,
                                                                 reinterpret_cast<unsigned long long*>(counts));
    return check_launch("confusion_kernel");
}
