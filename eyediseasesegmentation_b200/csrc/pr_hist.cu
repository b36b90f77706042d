// PR / ROC scoring: per-image score histograms in shared memory + threshold scan.
//
// Replaces, for src/main/aucpr.py of the reference:
//   - sklearn average_precision_score / roc_auc_score (aucpr.py:24,38): sort-based
//     on the CPU -> one HBM-bound pass that bins every pixel by a monotone key of its
//     fp32 score (symmetric about 1/2, 9 mantissa bits per binade of min(p, 1-p)), then a scan over the bins in descending score order;
//   - the 19 numpy passes per image of aucpr.py:60-81 / 136-170: the same histogram
//     gives every "score > threshold" count exactly, because the bin that shares a
//     key with threshold k is split by the `straddle` counters.
//
// Data layout: hist [n_images][2][EDS_PR_BINS] u32 (class-major so the positive and
// negative rows are contiguous for the scan), straddle [n_images][19][2] u32.
#include "common.cuh"
#include <cmath>
#include <cstring>

namespace eds {

constexpr int kBins = EDS_PR_BINS;
constexpr int kNT = EDS_PR_NTHRESH;
constexpr int kHistThreads = 1024;

// Bit pattern of the largest fp32 <= threshold k and the key of that pattern.
// "score(float32) > threshold(float64)" <=> bits(score) > thr_bits[k] for score >= 0.
struct ThreshTable {
    int bits[kNT];
    int key[kNT];
};
__constant__ ThreshTable c_thresh;
static ThreshTable h_thresh;
static bool h_thresh_ready = false;

__host__ __device__ __forceinline__ int float_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int b;
    memcpy(&b, &f, 4);
    return b;
#endif
}

// include/eds_b200.h "PR/ROC histogram geometry": monotone in p, symmetric about 1/2.
__host__ __device__ __forceinline__ int score_key(float p) {
    const bool hi = p >= 0.5f;
    const float q = hi ? 1.0f - p : p;
    int k = (float_bits(q) >> EDS_PR_KEY_SHIFT) - EDS_PR_KEY_BIAS;
    k = k < 0 ? 0 : k;
    k = k > EDS_PR_HALF - 1 ? EDS_PR_HALF - 1 : k;
    return hi ? kBins - 1 - k : k;
}

static const double kThresholds[kNT] = {0,   0.00001, 0.0001, 0.001, 0.01,  0.1,    0.2,
                                        0.3, 0.4,     0.5,    0.6,   0.7,   0.8,    0.9,
                                        0.99, 0.999,  0.9999, 0.99999, 1};

static int ensure_thresholds() {
    if (h_thresh_ready) return EDS_OK;
    for (int k = 0; k < kNT; ++k) {
        float f = (float)kThresholds[k];
        if ((double)f > kThresholds[k]) f = nextafterf(f, -1.0f);
        h_thresh.bits[k] = float_bits(f);
        h_thresh.key[k] = score_key(f);
    }
    cudaError_t e = cudaMemcpyToSymbol(c_thresh, &h_thresh, sizeof(h_thresh));
    if (e != cudaSuccess) {
        set_error("pr_hist: cudaMemcpyToSymbol failed: %s", cudaGetErrorString(e));
        return EDS_ERR_CUDA;
    }
    h_thresh_ready = true;
    return EDS_OK;
}

struct HistSmem {
    uint32_t hist[2 * kBins];
    uint32_t straddle[kNT * 2];
    uint8_t flag[kBins + 2];      // 1 where the bin shares its key with one of the 19 thresholds
};

__device__ __forceinline__ void straddle_update(HistSmem* s, int bits, int key, int cls) {
#pragma unroll 1
    for (int k = 0; k < kNT; ++k)
        if (key == c_thresh.key[k] && bits > c_thresh.bits[k]) atomicAdd(&s->straddle[k * 2 + cls], 1u);
}

__device__ __forceinline__ void hist_one(HistSmem* s, float p, uint8_t g) {
    const int bits = __float_as_int(p);
    const int key = score_key(p);
    const int cls = g != 0;
    atomicAdd(&s->hist[cls * kBins + key], 1u);
    if (s->flag[key]) straddle_update(s, bits, key, cls);
}

constexpr int kQuadUnroll = 2;    // float4 + uchar4 pairs in flight per thread

// grid = (splits, n_images); each CTA bins a contiguous slice of one image.
// The loop body is ~12 instructions per pixel (key: shift, subtract, clamp; class offset; one
// shared-memory atomic; one byte load for the threshold flag), which is what lets one SM retire
// several pixels per clock: the first version spent 60 instructions per pixel and was bound by
// instruction issue at ~1 pixel/clk/SM (profiles/r01_probe1_full.md), not by the atomics.
__global__ void __launch_bounds__(kHistThreads, 1)
pr_hist_kernel(const float* __restrict__ prob, const uint8_t* __restrict__ gt, int64_t n_pixels,
               uint32_t* __restrict__ g_hist, uint32_t* __restrict__ g_straddle, int vec_ok) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HistSmem* s = reinterpret_cast<HistSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int img = blockIdx.y;
    for (int i = tid; i < 2 * kBins; i += kHistThreads) s->hist[i] = 0;
    for (int i = tid; i < kBins + 2; i += kHistThreads) s->flag[i] = 0;
    if (tid < kNT * 2) s->straddle[tid] = 0;
    __syncthreads();
    if (tid < kNT) s->flag[c_thresh.key[tid]] = 1;
    __syncthreads();

    const float* p_img = prob + (int64_t)img * n_pixels;
    const uint8_t* g_img = gt + (int64_t)img * n_pixels;
    // slice boundaries in units of 4 pixels so the vector path stays aligned
    const int64_t n_quads = (n_pixels + 3) / 4;
    const int64_t q_per = (n_quads + gridDim.x - 1) / gridDim.x;
    const int64_t q_begin = (int64_t)blockIdx.x * q_per;
    int64_t q_end = q_begin + q_per;
    if (q_end > n_quads) q_end = n_quads;

    if (vec_ok && q_end > q_begin) {
        const float4* p4 = reinterpret_cast<const float4*>(p_img) + q_begin;
        const uchar4* g4 = reinterpret_cast<const uchar4*>(g_img) + q_begin;
        const int nq = (int)(q_end - q_begin);           // < 2^31: a CTA slice is part of one image
        uint32_t* hist = s->hist;
        const uint8_t* flag = s->flag;
        for (int q0 = 0; q0 < nq; q0 += kHistThreads * kQuadUnroll) {
            float4 p[kQuadUnroll];
            uchar4 g[kQuadUnroll];
            // unconditional loads (index clamped into the slice) keep both pairs in flight
#pragma unroll
            for (int u = 0; u < kQuadUnroll; ++u) {
                const int q = min(q0 + u * kHistThreads + tid, nq - 1);
                p[u] = __ldg(p4 + q);
                g[u] = __ldg(g4 + q);
            }
#pragma unroll
            for (int u = 0; u < kQuadUnroll; ++u) {
                const bool live = q0 + u * kHistThreads + tid < nq;
                const int b0 = __float_as_int(p[u].x), b1 = __float_as_int(p[u].y);
                const int b2 = __float_as_int(p[u].z), b3 = __float_as_int(p[u].w);
                const int k0 = score_key(p[u].x), k1 = score_key(p[u].y), k2 = score_key(p[u].z), k3 = score_key(p[u].w);
                const int a0 = k0 + (g[u].x ? kBins : 0), a1 = k1 + (g[u].y ? kBins : 0);
                const int a2 = k2 + (g[u].z ? kBins : 0), a3 = k3 + (g[u].w ? kBins : 0);
                const bool same4 = (a0 == a1) & (a1 == a2) & (a2 == a3);
                // flat regions (background of a fundus image): one atomic per thread, or per warp
                // when the whole warp sees one bin -- same-address atomics would serialise otherwise
                if (__all_sync(0xffffffffu, same4 && live)) {
                    const int lead = __shfl_sync(0xffffffffu, a0, 0);
                    if (__all_sync(0xffffffffu, a0 == lead)) {
                        if ((tid & 31) == 0) atomicAdd(hist + lead, 128u);
                    } else {
                        atomicAdd(hist + a0, 4u);
                    }
                } else if (live) {
                    atomicAdd(hist + a0, 1u);
                    atomicAdd(hist + a1, 1u);
                    atomicAdd(hist + a2, 1u);
                    atomicAdd(hist + a3, 1u);
                }
                if (live) {
                    const uint32_t hit = (uint32_t)flag[k0] | flag[k1] | flag[k2] | flag[k3];
                    if (hit) {
                        if (flag[k0]) straddle_update(s, b0, k0, g[u].x != 0);
                        if (flag[k1]) straddle_update(s, b1, k1, g[u].y != 0);
                        if (flag[k2]) straddle_update(s, b2, k2, g[u].z != 0);
                        if (flag[k3]) straddle_update(s, b3, k3, g[u].w != 0);
                    }
                }
            }
        }
    } else if (!vec_ok) {
        int64_t i_end = q_end * 4;
        if (i_end > n_pixels) i_end = n_pixels;
        for (int64_t i = q_begin * 4 + tid; i < i_end; i += kHistThreads) hist_one(s, p_img[i], g_img[i]);
    }
    __syncthreads();

    uint32_t* gh = g_hist + (int64_t)img * 2 * kBins;
    for (int i = tid; i < 2 * kBins; i += kHistThreads) {
        const uint32_t v = s->hist[i];
        if (v) atomicAdd(gh + i, v);
    }
    if (tid < kNT * 2) {
        const uint32_t v = s->straddle[tid];
        if (v) atomicAdd(g_straddle + (int64_t)img * kNT * 2 + tid, v);
    }
}

// ---- scan ------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int lo = __double2loint(v), hi = __double2hiint(v);
        lo = __shfl_xor_sync(0xffffffffu, lo, o);
        hi = __shfl_xor_sync(0xffffffffu, hi, o);
        v += __hiloint2double(hi, lo);
    }
    return v;
}

constexpr int kScanThreads = 1024;
constexpr int kPerThread = (kBins + kScanThreads - 1) / kScanThreads;

// One CTA per image.  Bins are walked in DESCENDING key order (highest score first),
// matching sklearn's _binary_clf_curve ordering.
__global__ void __launch_bounds__(kScanThreads, 1)
pr_scan_kernel(const uint32_t* __restrict__ g_hist, const uint32_t* __restrict__ g_straddle,
               double* __restrict__ ap, double* __restrict__ roc, uint64_t* __restrict__ counts,
               uint64_t* __restrict__ totals) {
    __shared__ uint32_t s_pos[kScanThreads];
    __shared__ uint32_t s_neg[kScanThreads];
    __shared__ double s_red[2][32];
    __shared__ uint32_t s_above[kNT][2];  // {tp, pp} over bins strictly above key[k]
    const int tid = threadIdx.x;
    const int img = blockIdx.x;
    const uint32_t* neg = g_hist + (int64_t)img * 2 * kBins;
    const uint32_t* pos = neg + kBins;

    // descending position d = 0 .. kBins-1 maps to bin kBins-1-d
    const int d_begin = tid * kPerThread;
    uint32_t lp = 0, ln = 0;
    for (int j = 0; j < kPerThread; ++j) {
        const int d = d_begin + j;
        if (d < kBins) {
            lp += pos[kBins - 1 - d];
            ln += neg[kBins - 1 - d];
        }
    }
    s_pos[tid] = lp;
    s_neg[tid] = ln;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 per-thread totals
    for (int off = 1; off < kScanThreads; off <<= 1) {
        uint32_t ap_ = 0, an_ = 0;
        if (tid >= off) {
            ap_ = s_pos[tid - off];
            an_ = s_neg[tid - off];
        }
        __syncthreads();
        s_pos[tid] += ap_;
        s_neg[tid] += an_;
        __syncthreads();
    }
    const uint32_t n_pos = s_pos[kScanThreads - 1];
    const uint32_t n_neg = s_neg[kScanThreads - 1];
    uint32_t tp = s_pos[tid] - lp;  // exclusive prefix = counts strictly above this chunk
    uint32_t fp = s_neg[tid] - ln;

    double ap_acc = 0.0, roc_acc = 0.0;
    const double inv_pos = n_pos ? 1.0 / (double)n_pos : 0.0;
    const double inv_neg = n_neg ? 1.0 / (double)n_neg : 0.0;
    for (int j = 0; j < kPerThread; ++j) {
        const int d = d_begin + j;
        if (d >= kBins) break;
        const int bin = kBins - 1 - d;
#pragma unroll 1
        for (int k = 0; k < kNT; ++k)
            if (bin == c_thresh.key[k]) {
                s_above[k][0] = tp;
                s_above[k][1] = tp + fp;
            }
        const uint32_t cp = pos[bin], cn = neg[bin];
        const uint32_t tp_new = tp + cp, fp_new = fp + cn;
        if (cp) ap_acc += ((double)cp * inv_pos) * ((double)tp_new / (double)(tp_new + fp_new));
        if (cn) roc_acc += ((double)cn * inv_neg) * (0.5 * ((double)tp + (double)tp_new) * inv_pos);
        tp = tp_new;
        fp = fp_new;
    }
    ap_acc = warp_sum_f64(ap_acc);
    roc_acc = warp_sum_f64(roc_acc);
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = ap_acc;
        s_red[1][tid >> 5] = roc_acc;
    }
    __syncthreads();
    if (tid < 32) {
        double a = s_red[0][tid], r = s_red[1][tid];
        a = warp_sum_f64(a);
        r = warp_sum_f64(r);
        if (tid == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            ap[img] = n_pos ? a : nan;
            roc[img] = (n_pos && n_neg) ? r : nan;
            totals[img * 2 + 0] = n_pos;
            totals[img * 2 + 1] = n_neg;
        }
    }
    if (tid < kNT) {
        const uint32_t* st = g_straddle + (int64_t)img * kNT * 2 + tid * 2;
        const uint64_t tpk = (uint64_t)s_above[tid][0] + st[1];
        const uint64_t ppk = (uint64_t)s_above[tid][1] + st[0] + st[1];
        counts[((int64_t)img * kNT + tid) * 2 + 0] = tpk;
        counts[((int64_t)img * kNT + tid) * 2 + 1] = ppk;
    }
}

}  // namespace eds

using namespace eds;

extern "C" int eds_pr_hist_f32(const float* prob, const uint8_t* gt, int64_t n_pixels, int n_images,
                               uint32_t* hist, uint32_t* straddle, int splits, void* stream) {
    EDS_REQUIRE(prob && gt && hist && straddle, "pr_hist: null pointer");
    EDS_REQUIRE(n_pixels > 0 && n_images > 0, "pr_hist: empty input (n_pixels=%lld n_images=%d)",
                (long long)n_pixels, n_images);
    EDS_REQUIRE(n_images <= 65535, "pr_hist: n_images %d > 65535", n_images);
    if (int rc = ensure_thresholds()) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (splits <= 0) {
        // enough CTAs for >= 2 waves when several images are batched, one wave for one image
        splits = ceil_div(2 * sms, n_images);
        if (splits > sms) splits = sms;
        if (n_images == 1) splits = sms;
    }
    const int64_t min_quads = 4096;  // do not split below 16 Ki pixels per CTA
    const int64_t n_quads = (n_pixels + 3) / 4;
    if ((int64_t)splits * min_quads > n_quads) splits = (int)((n_quads + min_quads - 1) / min_quads);
    if (splits < 1) splits = 1;
    const int vec_ok = (n_pixels % 4 == 0) && (((uintptr_t)prob & 15) == 0) && (((uintptr_t)gt & 3) == 0);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pr_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(HistSmem));
        if (e != cudaSuccess) {
            set_error("pr_hist: cannot opt in to %zu B shared memory: %s", sizeof(HistSmem),
                      cudaGetErrorString(e));
            return EDS_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(splits, n_images);
    pr_hist_kernel<<<grid, kHistThreads, sizeof(HistSmem), as_stream(stream)>>>(prob, gt, n_pixels, hist,
                                                                             straddle, vec_ok);
    return check_launch("pr_hist_kernel");
}

extern "C" int eds_pr_scan(const uint32_t* hist, const uint32_t* straddle, int n_images, double* ap,
                           double* roc, uint64_t* counts, uint64_t* totals, void* stream) {
    EDS_REQUIRE(hist && straddle && ap && roc && counts && totals, "pr_scan: null pointer");
    EDS_REQUIRE(n_images > 0, "pr_scan: n_images=%d", n_images);
    if (int rc = ensure_thresholds()) return rc;
    pr_scan_kernel<<<n_images, kScanThreads, 0, as_stream(stream)>>>(hist, straddle, ap, roc, counts, totals);
    return check_launch("pr_scan_kernel");
}
