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
// Kernels: pr_hist_kernel (whole images, one persistent wave, 0.79 of the HBM peak on B200),
// pr_hist_rects_kernel (the rectangles one rank owns under the (image, tile) partition), pr_scan_kernel.
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

static PerDevice g_thresh_once;      // __constant__ memory is per device
static int ensure_thresholds() {
    return g_thresh_once.run([](int) {
        ThreshTable h;
        for (int k = 0; k < kNT; ++k) {
            float f = (float)kThresholds[k];
            if ((double)f > kThresholds[k]) f = nextafterf(f, -1.0f);
            h.bits[k] = float_bits(f);
            h.key[k] = score_key(f);
            if (k > 0 && h.key[k] <= h.key[k - 1]) {      // the counter tags name ONE threshold per bin
                set_error("pr_hist: thresholds %d and %d share a histogram key", k - 1, k);
                return (int)EDS_ERR_INVALID;
            }
        }
        cudaError_t e = cudaMemcpyToSymbol(c_thresh, &h, sizeof(h));
        if (e != cudaSuccess) {
            set_error("pr_hist: cudaMemcpyToSymbol failed: %s", cudaGetErrorString(e));
            return (int)EDS_ERR_CUDA;
        }
        return (int)EDS_OK;
    });
}

// Shared-memory image of one CTA: interleaved counters [key][class] (class 1 = positive), the straddle
// counters and the bit patterns of the thresholds.
// The counters of the 19 bins that share a key with a threshold carry a TAG in their top six bits:
// bit 31 set, bits 26-30 = index of that threshold.  The atomic's return value therefore tells a pixel
// that it landed in such a bin and which threshold to compare with, so the common path needs no second
// shared-memory access per pixel (round 1 paid one LDS.U8 per pixel for a flag table -- as many
// shared-memory wavefronts as the atomics themselves -- and a 19-step search per tagged pixel).
// The count field is 26 bits: the launcher keeps a CTA's share of one image below 2^26 pixels.
constexpr uint32_t kTag = 0x80000000u;
constexpr int kTagShift = 26;
constexpr uint32_t kCountMask = (1u << kTagShift) - 1;
struct HistSmem {
    uint32_t hist[2 * kBins];
    uint32_t straddle[kNT * 2];
    int thr_bits[32];
};

// Byte offset of the counter of (score, label byte j of the 4-label word) inside HistSmem::hist.
// q = min(p, 1-p) equals "p >= 1/2 ? 1-p : p" for every input (1-p is exact for p in [1/2, 1]); the two
// integer clamps keep the documented behaviour for p < 2^-24, p > 1, negative values and NaN.
__device__ __forceinline__ uint32_t counter_offset(float p, uint32_t nz, int j) {
    const float q = fminf(p, 1.0f - p);
    int k = (__float_as_int(q) >> EDS_PR_KEY_SHIFT) - EDS_PR_KEY_BIAS;
    k = max(k, 0);
    k = min(k, EDS_PR_HALF - 1);
    const int key = p >= 0.5f ? kBins - 1 - k : k;
    return ((uint32_t)key << 3) | ((nz >> (8 * j + 5)) & 4u);
}

// 0x80 in every byte of g that is non-zero
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t g) {
    return (((g & 0x7f7f7f7fu) + 0x7f7f7f7fu) | g) & 0x80808080u;
}

// A pixel whose atomic returned a tagged counter: strictly above the threshold of that bin?
__device__ __forceinline__ void straddle_one(HistSmem* s, uint32_t old, float p, uint32_t off) {
    if (old & kTag) {
        const int k = (old >> kTagShift) & 31;
        if (__float_as_int(p) > s->thr_bits[k]) atomicAdd(&s->straddle[k * 2 + ((off >> 2) & 1)], 1u);
    }
}

__device__ __forceinline__ void hist_one(HistSmem* s, float p, uint8_t g) {
    const uint32_t off = counter_offset(p, g ? 0x80u : 0u, 0);
    const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(s->hist) + off), 1u);
    straddle_one(s, old, p, off);
}

__device__ __forceinline__ void hist_clear(HistSmem* s, int tid) {
    for (int i = tid; i < 2 * kBins; i += kHistThreads) s->hist[i] = 0;
    if (tid < kNT * 2) s->straddle[tid] = 0;
    if (tid < 32) s->thr_bits[tid] = tid < kNT ? c_thresh.bits[tid] : 0x7fffffff;
    __syncthreads();
    if (tid < kNT) {                       // the 19 keys are distinct (checked on the host)
        const uint32_t tag = kTag | ((uint32_t)tid << kTagShift);
        s->hist[2 * c_thresh.key[tid]] = tag;
        s->hist[2 * c_thresh.key[tid] + 1] = tag;
    }
    __syncthreads();
}

__device__ __forceinline__ void hist_flush(HistSmem* s, int tid, uint32_t* __restrict__ gh, uint32_t* __restrict__ gs) {
    __syncthreads();
    for (int i = tid; i < 2 * kBins; i += kHistThreads) {
        const uint32_t v = s->hist[i] & kCountMask;       // i = key * 2 + class
        if (v) atomicAdd(gh + (i & 1) * kBins + (i >> 1), v);
    }
    if (tid < kNT * 2) {
        const uint32_t v = s->straddle[tid];
        if (v) atomicAdd(gs + tid, v);
    }
    __syncthreads();
}

constexpr int kQuadUnroll = 4;    // float4 + uchar4 pairs in flight per thread (80 KB per SM)
constexpr int kCheckEvery = 8;    // steps between two crowded-bin tests outside crowded regions (power of two)
constexpr int kAggLanes = 8;      // lanes sharing one counter from which the warp-aggregated path pays

// Persistent grid of <= one CTA per SM.  The test set is ONE linear range of 4-pixel quads (image-major);
// CTA c takes the contiguous share [c*T/G, (c+1)*T/G) and, when its share crosses an image boundary,
// flushes and clears its shared-memory histogram there.  (Round 1 launched ceil(2*SMs/n_images) CTAs per
// image: 297 CTAs for 27 images = two full waves plus ONE straggler CTA.)
// Common path per pixel, ~13 instructions: key (fmin, shift, clamps, mirror), label bit, one returning
// shared-memory atomic, one OR of the returned tag; one warp-uniform branch per 16 pixels of a thread.
__global__ void __launch_bounds__(kHistThreads, 1)
pr_hist_kernel(const float* __restrict__ prob, const uint8_t* __restrict__ gt, int64_t n_pixels, int n_images,
               uint32_t* __restrict__ g_hist, uint32_t* __restrict__ g_straddle, int vec_ok, int per_image_splits) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HistSmem* s = reinterpret_cast<HistSmem*>(smem_raw);
    const int tid = threadIdx.x;
    char* hbase = reinterpret_cast<char*>(s->hist);

    const int64_t n_quads = (n_pixels + 3) / 4;           // per image; vec_ok implies n_pixels % 4 == 0
    int64_t w_begin, w_end;                               // this CTA's share of [0, n_images * n_quads)
    if (per_image_splits > 0) {                           // caller-fixed split: CTA = (image, split)
        const int img = blockIdx.x / per_image_splits, sp = blockIdx.x % per_image_splits;
        const int64_t q_per = (n_quads + per_image_splits - 1) / per_image_splits;
        w_begin = (int64_t)img * n_quads + min((int64_t)sp * q_per, n_quads);
        w_end = (int64_t)img * n_quads + min((int64_t)(sp + 1) * q_per, n_quads);
    } else {
        const int64_t total = (int64_t)n_images * n_quads;
        w_begin = total / gridDim.x * blockIdx.x + min((int64_t)blockIdx.x, total % gridDim.x);
        w_end = w_begin + total / gridDim.x + (blockIdx.x < total % gridDim.x ? 1 : 0);
    }

    while (w_begin < w_end) {
        const int img = (int)(w_begin / n_quads);
        const int64_t q_begin = w_begin - (int64_t)img * n_quads;
        const int64_t q_end = min(q_begin + (w_end - w_begin), n_quads);
        w_begin += q_end - q_begin;
        const float* p_img = prob + (int64_t)img * n_pixels;
        const uint8_t* g_img = gt + (int64_t)img * n_pixels;
        hist_clear(s, tid);

        if (vec_ok) {
            const float4* p4 = reinterpret_cast<const float4*>(p_img) + q_begin;
            const uint32_t* g4 = reinterpret_cast<const uint32_t*>(g_img) + q_begin;
            const int nq = (int)(q_end - q_begin);           // < 2^24: the launcher bounds a share
            constexpr int kStep = kHistThreads * kQuadUnroll;
            const int n_full = nq / kStep * kStep;
            bool crowded = false;                            // warp-uniform state of the crowded-bin test
            int since_check = 0;
            for (int q0 = tid; q0 < n_full; q0 += kStep) {
                float4 p[kQuadUnroll];
                uint32_t g[kQuadUnroll];
#pragma unroll
                for (int u = 0; u < kQuadUnroll; ++u) {
                    p[u] = __ldcs(p4 + q0 + u * kHistThreads);
                    g[u] = __ldcs(g4 + q0 + u * kHistThreads);
                }
                uint32_t off[kQuadUnroll][4];
#pragma unroll
                for (int u = 0; u < kQuadUnroll; ++u) {
                    const uint32_t nz = nonzero_bytes(g[u]);
                    off[u][0] = counter_offset(p[u].x, nz, 0);
                    off[u][1] = counter_offset(p[u].y, nz, 1);
                    off[u][2] = counter_offset(p[u].z, nz, 2);
                    off[u][3] = counter_offset(p[u].w, nz, 3);
                }
                // Crowded bins (saturated background: p < 2^-24 or p == 1 over large areas; flat regions).
                // Same-address shared-memory atomics serialise lane by lane (measured: 1.4 TB/s on a flat map),
                // so when >= kAggLanes lanes of the warp open this step in ONE counter -- candidates: the lowest
                // and the highest offset in the warp, which is where saturation lands -- every pixel of that
                // counter is tallied in registers and added by ONE atomic per warp; the rest go one by one.
                // The test itself (two warp reductions, two votes) stalls the warp's 16 atomics behind it --
                // measured 0.80 -> 0.59 of the HBM peak on uniform scores when done at every step -- so it runs
                // on every kCheckEvery-th step and on every step while the warp is inside a crowded region
                // (crowding is spatially coherent: a region is entered at most kCheckEvery - 1 steps late).
                uint32_t lead = 0;
                int n_lead = 0;
                if (crowded || since_check == 0) {
                    const uint32_t o00 = off[0][0];
                    lead = __reduce_min_sync(0xffffffffu, o00);
                    n_lead = __popc(__ballot_sync(0xffffffffu, o00 == lead));
                    if (n_lead < kAggLanes) {
                        lead = __reduce_max_sync(0xffffffffu, o00);
                        n_lead = __popc(__ballot_sync(0xffffffffu, o00 == lead));
                    }
                    crowded = n_lead >= kAggLanes;
                }
                since_check = (since_check + 1) & (kCheckEvery - 1);
                if (n_lead >= kAggLanes) {
                    // crowded step: the lead counter gets ONE atomic per warp; the other pixels go one by one and
                    // remember the top byte (tag bit + threshold index) of the counter they hit
                    uint32_t pk[kQuadUnroll], cnt = 0;
#pragma unroll
                    for (int u = 0; u < kQuadUnroll; ++u) {
                        pk[u] = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool is_lead = off[u][j] == lead;
                            cnt += is_lead;
                            if (!is_lead)
                                pk[u] = __byte_perm(pk[u], atomicAdd(reinterpret_cast<uint32_t*>(hbase + off[u][j]), 1u),
                                                    0x3210 & ~(0xf << (4 * j)) | (7 << (4 * j)));
                        }
                    }
                    const uint32_t total = __reduce_add_sync(0xffffffffu, cnt);
                    uint32_t old_lead = 0;
                    if ((tid & 31) == 0) old_lead = atomicAdd(reinterpret_cast<uint32_t*>(hbase + lead), total);
                    old_lead = __shfl_sync(0xffffffffu, old_lead, 0);
                    if (old_lead & kTag) {                   // the crowded bin holds a threshold (0 and 1 do)
                        const int k = (old_lead >> kTagShift) & 31;
                        const int thr = s->thr_bits[k];
                        uint32_t above = 0;
#pragma unroll
                        for (int u = 0; u < kQuadUnroll; ++u) {
                            above += (off[u][0] == lead) & (__float_as_int(p[u].x) > thr);
                            above += (off[u][1] == lead) & (__float_as_int(p[u].y) > thr);
                            above += (off[u][2] == lead) & (__float_as_int(p[u].z) > thr);
                            above += (off[u][3] == lead) & (__float_as_int(p[u].w) > thr);
                        }
                        above = __reduce_add_sync(0xffffffffu, above);
                        if ((tid & 31) == 0 && above) atomicAdd(&s->straddle[k * 2 + ((lead >> 2) & 1)], above);
                    }
#pragma unroll
                    for (int u = 0; u < kQuadUnroll; ++u) {  // tagged pixels outside the lead counter (rare here)
                        const float pv[4] = {p[u].x, p[u].y, p[u].z, p[u].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            straddle_one(s, (pk[u] >> (8 * j)) << 24, pv[j], off[u][j]);
                    }
                } else {
                    // common step: 16 returning atomics; on uniform scores three steps in four see SOME tagged
                    // pixel in the warp, so the follow-up must cost (almost) nothing for lanes without one:
                    // returned values stay in registers, one OR + one branch per thread, one test per pixel
                    uint32_t old[kQuadUnroll][4], tag = 0;
#pragma unroll
                    for (int u = 0; u < kQuadUnroll; ++u)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            old[u][j] = atomicAdd(reinterpret_cast<uint32_t*>(hbase + off[u][j]), 1u);
                            tag |= old[u][j];
                        }
                    if (tag & kTag) {
#pragma unroll
                        for (int u = 0; u < kQuadUnroll; ++u) {
                            straddle_one(s, old[u][0], p[u].x, off[u][0]);
                            straddle_one(s, old[u][1], p[u].y, off[u][1]);
                            straddle_one(s, old[u][2], p[u].z, off[u][2]);
                            straddle_one(s, old[u][3], p[u].w, off[u][3]);
                        }
                    }
                }
            }
            for (int q = n_full + tid; q < nq; q += kHistThreads) {      // tail: fewer than kStep quads
                const float4 pq = __ldcs(p4 + q);
                const uint32_t gq = __ldcs(g4 + q);
                hist_one(s, pq.x, (uint8_t)(gq & 0xff));
                hist_one(s, pq.y, (uint8_t)((gq >> 8) & 0xff));
                hist_one(s, pq.z, (uint8_t)((gq >> 16) & 0xff));
                hist_one(s, pq.w, (uint8_t)(gq >> 24));
            }
        } else {
            const int64_t i_end = min(q_end * 4, n_pixels);
            for (int64_t i = q_begin * 4 + tid; i < i_end; i += kHistThreads) hist_one(s, p_img[i], g_img[i]);
        }
        hist_flush(s, tid, g_hist + (int64_t)img * 2 * kBins, g_straddle + (int64_t)img * kNT * 2);
    }
}

// Histogram of a list of RECTANGLES of one image (row stride = image width): the pixels a rank OWNS under
// the (image, tile) partition of SURVEY.md 8e -- a tile's window minus every later tile's window, which for
// make_grid's layouts is a handful of rectangles whose origins follow the tile origins (not 16-byte aligned).
// One warp per rectangle row, lanes along the row (coalesced 4-byte loads); same shared-memory histogram,
// tags and flush as the streaming kernel, so the two agree bin for bin.
constexpr int kMaxRects = 32;
struct HistRects {
    int y[kMaxRects], x[kMaxRects], h[kMaxRects], w[kMaxRects];
    int n;
};

__global__ void __launch_bounds__(kHistThreads, 1)
pr_hist_rects_kernel(const float* __restrict__ prob, const uint8_t* __restrict__ gt, int row_stride, HistRects r,
                     uint32_t* __restrict__ g_hist, uint32_t* __restrict__ g_straddle) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HistSmem* s = reinterpret_cast<HistSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    hist_clear(s, tid);
    int total_rows = 0;
    for (int k = 0; k < r.n; ++k) total_rows += r.h[k];
    const int per = (total_rows + gridDim.x - 1) / gridDim.x;
    const int row_begin = blockIdx.x * per, row_end = min(total_rows, row_begin + per);
    for (int row = row_begin + warp; row < row_end; row += kHistThreads / 32) {
        int k = 0, rr = row;
        while (rr >= r.h[k]) { rr -= r.h[k]; ++k; }
        const int64_t base = (int64_t)(r.y[k] + rr) * row_stride + r.x[k];
        for (int c = lane; c < r.w[k]; c += 32) hist_one(s, __ldcs(prob + base + c), __ldcs(gt + base + c));
    }
    hist_flush(s, tid, g_hist, g_straddle);
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

static PerDevice g_hist_attr_once;
static int hist_smem_opt_in(int) {
    return smem_opt_in(g_hist_attr_once, pr_hist_kernel, (int)sizeof(HistSmem), "pr_hist");
}

extern "C" int eds_pr_hist_f32(const float* prob, const uint8_t* gt, int64_t n_pixels, int n_images,
                               uint32_t* hist, uint32_t* straddle, int splits, void* stream) {
    EDS_REQUIRE(prob && gt && hist && straddle, "pr_hist: null pointer");
    EDS_REQUIRE(n_pixels > 0 && n_images > 0, "pr_hist: empty input (n_pixels=%lld n_images=%d)",
                (long long)n_pixels, n_images);
    EDS_REQUIRE(n_images <= 65535, "pr_hist: n_images %d > 65535", n_images);
    EDS_REQUIRE(n_pixels < ((int64_t)1 << 33), "pr_hist: n_pixels %lld per image exceeds 2^33", (long long)n_pixels);
    if (int rc = ensure_thresholds()) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t min_quads = 4096;  // do not split below 16 Ki pixels per CTA
    const int64_t n_quads = (n_pixels + 3) / 4;
    int grid;
    if (splits > 0) {                // caller-fixed CTAs per image
        if ((int64_t)splits * min_quads > n_quads) splits = (int)((n_quads + min_quads - 1) / min_quads);
        grid = splits * n_images;
    } else {                         // one persistent wave over the linear range of all images
        const int64_t total = n_quads * n_images;
        grid = (int)(total / min_quads < sms ? (total + min_quads - 1) / min_quads : sms);
        if (grid < 1) grid = 1;
    }
    // the count field of a shared-memory counter is 26 bits: bound a CTA's share of one image
    const int64_t max_quads = ((int64_t)1 << 24) - 1;
    if (splits > 0) {
        if ((n_quads + splits - 1) / splits > max_quads) {
            splits = (int)((n_quads + max_quads - 1) / max_quads);
            grid = splits * n_images;
        }
    } else if ((n_quads * n_images + grid - 1) / grid > max_quads) {
        grid = (int)((n_quads * n_images + max_quads - 1) / max_quads);
    }
    const int vec_ok = (n_pixels % 4 == 0) && (((uintptr_t)prob & 15) == 0) && (((uintptr_t)gt & 3) == 0);
    if (int rc = hist_smem_opt_in(dev)) return rc;
    pr_hist_kernel<<<grid, kHistThreads, sizeof(HistSmem), as_stream(stream)>>>(prob, gt, n_pixels, n_images, hist,
                                                                             straddle, vec_ok, splits > 0 ? splits : 0);
    return check_launch("pr_hist_kernel");
}

extern "C" int eds_pr_hist_rects_f32(const float* prob, const uint8_t* gt, int img_h, int img_w, int n_rects,
                                     const int* rects_host, uint32_t* hist, uint32_t* straddle, void* stream) {
    EDS_REQUIRE(prob && gt && hist && straddle && rects_host, "pr_hist_rects: null pointer");
    EDS_REQUIRE(img_h > 0 && img_w > 0, "pr_hist_rects: bad image size %dx%d", img_h, img_w);
    EDS_REQUIRE(n_rects >= 0 && n_rects <= kMaxRects, "pr_hist_rects: n_rects=%d (0..%d)", n_rects, kMaxRects);
    if (int rc = ensure_thresholds()) return rc;
    HistRects r;
    memset(&r, 0, sizeof(r));
    int64_t rows = 0, pixels = 0;
    for (int k = 0; k < n_rects; ++k) {
        const int y = rects_host[4 * k], x = rects_host[4 * k + 1], h = rects_host[4 * k + 2], w = rects_host[4 * k + 3];
        EDS_REQUIRE(y >= 0 && x >= 0 && h >= 0 && w >= 0 && y + h <= img_h && x + w <= img_w,
                    "pr_hist_rects: rectangle %d (y=%d x=%d h=%d w=%d) leaves the %dx%d image", k, y, x, h, w, img_h, img_w);
        if (h == 0 || w == 0) continue;
        r.y[r.n] = y; r.x[r.n] = x; r.h[r.n] = h; r.w[r.n] = w;
        ++r.n;
        rows += h;
        pixels += (int64_t)h * w;
    }
    if (r.n == 0) return EDS_OK;                       // nothing owned: nothing to add
    EDS_REQUIRE(pixels < ((int64_t)1 << 26), "pr_hist_rects: %lld pixels exceed the 2^26 counter field",
                (long long)pixels);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = (int)((pixels + (1 << 16) - 1) >> 16);   // >= 64 Ki pixels per CTA: the 188 KB clear + flush amortise
    if (grid > sms) grid = sms;
    if (grid > rows) grid = (int)rows;
    if (grid < 1) grid = 1;
    static PerDevice once;
    if (int rc = smem_opt_in(once, pr_hist_rects_kernel, (int)sizeof(HistSmem), "pr_hist_rects")) return rc;
    pr_hist_rects_kernel<<<grid, kHistThreads, sizeof(HistSmem), as_stream(stream)>>>(prob, gt, img_w, r, hist, straddle);
    return check_launch("pr_hist_rects_kernel");
}

extern "C" int eds_pr_scan(const uint32_t* hist, const uint32_t* straddle, int n_images, double* ap,
                           double* roc, uint64_t* counts, uint64_t* totals, void* stream) {
    EDS_REQUIRE(hist && straddle && ap && roc && counts && totals, "pr_scan: null pointer");
    EDS_REQUIRE(n_images > 0, "pr_scan: n_images=%d", n_images);
    if (int rc = ensure_thresholds()) return rc;
    pr_scan_kernel<<<n_images, kScanThreads, 0, as_stream(stream)>>>(hist, straddle, ap, roc, counts, totals);
    return check_launch("pr_scan_kernel");
}
