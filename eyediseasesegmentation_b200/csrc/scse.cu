// Decoder concat + SCSE in two streaming passes -- the FIRST formulation, kept for already materialised maps
// and as an independent cross-check in the tests.  The networks run the deferred-gate formulation of
// scse_gated.cu (per-source statistics at native resolution + one gated write of the concat), which moves
// 40 % fewer bytes; see DESIGN.md 3.3.
//
// Reference: DecoderBlock.forward (src/main/archs/unetplusplusstar.py:127-161):
//     x_up = interpolate(x, 2, bilinear); x = cat([x_up, skip]); x = attention1(x)   # smp SCSEModule
// and attention2 on the block output.  SCSE needs two global quantities of its input before
// it can scale a single element: the channel means (cSE) and, per pixel, w_sse . x (sSE).
//
//   pass 1  eds_concat_stats : reads the sources ONCE, writes the concatenated map (optional),
//                              accumulates channel means and the per-pixel sSE logit;
//   (tiny)  eds_se_gate      : cSE MLP on the means;
//   pass 2  eds_scse_scale   : y = x * (cgate[n][c] + sigmoid(logit[n][p])) in place.
//
// Pass 1 layout: grid = (pixel chunks, N, 256-channel chunks).  A warp walks pixels; lane L owns
// ONE 8-channel vector (chunk*32 + L) of every pixel it visits, so its channel sums and sSE
// weights live in 16 registers, every global access is a contiguous 512 B per warp instruction,
// and the low register count keeps ~40 warps per SM in flight (the first version held 4
// vectors per lane, ran 16 warps/SM and reached 31 % of HBM).
#include "common.cuh"

namespace eds {

constexpr int kStatsThreads = 256;
constexpr int kStatsUnroll = 2;     // pixels in flight per lane group

struct StatSrc {
    const void* ptr[6];   // [0] may be upsampled; the rest are same-resolution maps
    int ch[6];
    int n;
};

template <typename T>
__device__ __forceinline__ void load_up2x(const T* __restrict__ base, int h, int w, int C0, int oy, int ox, int mode,
                                          float (&v)[8]) {
    if (mode == EDS_UP_NONE) {
        Vec8<T>::ld(base + ((int64_t)oy * w + ox) * C0, v);
        return;
    }
    const int iy = oy >> 1, ix = ox >> 1;
    if (mode == EDS_UP_NEAREST) {
        Vec8<T>::ld(base + ((int64_t)iy * w + ix) * C0, v);
        return;
    }
    int y0, y1, x0, x1;
    float ly, lx;
    if (oy & 1) { y0 = iy; y1 = min(iy + 1, h - 1); ly = 0.25f; }
    else if (iy == 0) { y0 = 0; y1 = min(1, h - 1); ly = 0.f; }
    else { y0 = iy - 1; y1 = iy; ly = 0.75f; }
    if (ox & 1) { x0 = ix; x1 = min(ix + 1, w - 1); lx = 0.25f; }
    else if (ix == 0) { x0 = 0; x1 = min(1, w - 1); lx = 0.f; }
    else { x0 = ix - 1; x1 = ix; lx = 0.75f; }
    float a[8], b[8], c[8], d[8];
    Vec8<T>::ld(base + ((int64_t)y0 * w + x0) * C0, a);
    Vec8<T>::ld(base + ((int64_t)y0 * w + x1) * C0, b);
    Vec8<T>::ld(base + ((int64_t)y1 * w + x0) * C0, c);
    Vec8<T>::ld(base + ((int64_t)y1 * w + x1) * C0, d);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = hy * (hx * a[i] + lx * b[i]) + ly * (hx * c[i] + lx * d[i]);
}

// h, w: resolution of source 0; output resolution H x W = up*h x up*w.  `lpp` lanes (a power of
// two) cooperate on one pixel, so a warp covers 32/lpp pixels per step (16 channels -> 2 lanes
// per pixel, 16 pixels per warp step).  With more than 256 channels gridDim.z > 1 and the sSE
// logit is accumulated with one atomicAdd per (pixel, channel chunk) into a zeroed buffer.
template <typename T>
__global__ void __launch_bounds__(kStatsThreads)
concat_stats_kernel(StatSrc src, int h, int w, int mode, int Ctot, int lpp, const float* __restrict__ w_sse,
                    float b_sse, float inv_hw, T* __restrict__ y, float* __restrict__ chan_mean,
                    float* __restrict__ sse_logit) {
    __shared__ float s_sum[kStatsThreads / 32][32][8 + 1];
    const int up = mode == EDS_UP_NONE ? 1 : 2;
    const int H = up * h, W = up * w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % lpp, sub = lane / lpp, ppw = 32 / lpp;
    const int n = blockIdx.y;
    const int v8 = blockIdx.z * 32 + l;              // this lane's 8-channel vector
    const bool live = v8 < Ctot / 8;
    const bool multi = gridDim.z > 1;

    int c = v8 * 8, k = 0;
    if (live)
        while (k < src.n - 1 && c >= src.ch[k]) { c -= src.ch[k]; ++k; }
    const int sc = live ? src.ch[k] : 0;             // pixel stride of the owning source
    const bool is_up = live && k == 0;
    const int n_pix = H * W;
    const T* sp = live ? reinterpret_cast<const T*>(src.ptr[k]) + (int64_t)n * (is_up ? h * w : n_pix) * sc + c
                       : nullptr;
    float wv[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        wv[i] = (live && w_sse) ? w_sse[v8 * 8 + i] : 0.f;
        acc[i] = 0.f;
    }
    const float bias = (blockIdx.z == 0) ? b_sse : 0.f;

    const int per = (n_pix + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * per;
    const int p_end = min(n_pix, p_begin + per);
    const int step = (kStatsThreads / 32) * ppw * kStatsUnroll;
    // the loop bound is warp-uniform (shuffles below); lanes past the end are masked by `ok`
    for (int pb = p_begin + warp * ppw * kStatsUnroll; pb < p_end; pb += step) {
        float v[kStatsUnroll][8];
        int p[kStatsUnroll];
        bool ok[kStatsUnroll];
#pragma unroll
        for (int u = 0; u < kStatsUnroll; ++u) {        // all loads first: 2 pixels in flight
            p[u] = pb + u * ppw + sub;
            ok[u] = live && p[u] < p_end;
            if (ok[u]) {
                if (is_up) {
                    const int oy = p[u] / W, ox = p[u] - oy * W;
                    load_up2x<T>(sp, h, w, sc, oy, ox, mode, v[u]);
                } else {
                    Vec8<T>::ld(sp + (int64_t)p[u] * sc, v[u]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < kStatsUnroll; ++u) {
            if (ok[u] && y) Vec8<T>::st(y + ((int64_t)n * n_pix + p[u]) * Ctot + v8 * 8, v[u]);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dot = fmaf(v[u][i], wv[i], dot);
                acc[i] += v[u][i];
            }
            if (sse_logit) {
                for (int o = lpp >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (l == 0 && p[u] < p_end) {
                    float* dst = sse_logit + (int64_t)n * n_pix + p[u];
                    if (multi) atomicAdd(dst, dot + bias);
                    else *dst = dot + bias;
                }
            }
        }
    }

    // channel sums: 8 warps x 32 lanes -> smem -> one atomicAdd per channel per CTA
#pragma unroll
    for (int i = 0; i < 8; ++i) s_sum[warp][lane][i] = acc[i];
    __syncthreads();
    // thread t < lpp*8 owns channel (t/8 -> vector, t%8 -> element) of this chunk
    if (threadIdx.x < lpp * 8) {
        const int vl = threadIdx.x >> 3, e = threadIdx.x & 7;
        const int ch = (blockIdx.z * 32 + vl) * 8 + e;
        if (ch < Ctot) {
            float t = 0.f;
            for (int wi = 0; wi < kStatsThreads / 32; ++wi)
                for (int s = 0; s < ppw; ++s) t += s_sum[wi][s * lpp + vl][e];
            atomicAdd(chan_mean + (int64_t)n * Ctot + ch, t * inv_hw);
        }
    }
}

// Bilinear x2 fast path (>= 256 channels, so a warp = one pixel's 256-channel chunk).  A warp walks
// one OUTPUT ROW at a time, two output pixels (2i, 2i+1) per step.  For the upsampled source the
// vertical blend of each low-resolution column is computed once and slides through registers
// (prev, cur, next), so a step costs 2 vector loads instead of 8 and no index division:
//     out[2i]   = 0.25 * col(i-1) + 0.75 * col(i)        col(j) = (1-ly) * row_y0[j] + ly * row_y1[j]
//     out[2i+1] = 0.75 * col(i)   + 0.25 * col(i+1)      (columns clamped at the borders)
// which is F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) with the two lerps
// in the other order (fp32 rounding differs by ~1 ulp).  Lanes that own a same-resolution source
// load pixels 2i, 2i+1 directly; both kinds run the same instruction stream (selects, no branches).
// grid = (row groups, N, 256-channel chunks).
template <typename T>
__global__ void __launch_bounds__(kStatsThreads)
concat_stats_bilinear_kernel(StatSrc src, int h, int w, int Ctot, const float* __restrict__ w_sse, float b_sse,
                             float inv_hw, T* __restrict__ y, float* __restrict__ chan_mean,
                             float* __restrict__ sse_logit) {
    __shared__ float s_sum[kStatsThreads / 32][32][8 + 1];
    const int H = 2 * h, W = 2 * w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const int v8 = blockIdx.z * 32 + lane;
    const bool live = v8 < Ctot / 8;
    const bool multi = gridDim.z > 1;
    int c = v8 * 8, k = 0;
    if (live)
        while (k < src.n - 1 && c >= src.ch[k]) { c -= src.ch[k]; ++k; }
    const int sc = live ? src.ch[k] : 8;
    const bool is_up = live && k == 0;
    // dead lanes read lane 0's data of source 0 (valid memory) and never store
    const T* sp = reinterpret_cast<const T*>(src.ptr[live ? k : 0]) +
                  (int64_t)n * (is_up || !live ? h * w : H * W) * (live ? sc : src.ch[0]) + (live ? c : 0);
    const int stride = live ? sc : src.ch[0];
    float wv[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        wv[i] = (live && w_sse) ? w_sse[v8 * 8 + i] : 0.f;
        acc[i] = 0.f;
    }
    const float bias = (blockIdx.z == 0) ? b_sse : 0.f;
    const bool up_like = is_up || !live;

    for (int oy = blockIdx.x * (kStatsThreads / 32) + warp; oy < H; oy += gridDim.x * (kStatsThreads / 32)) {
        const int iy = oy >> 1;
        int y0, y1;
        float ly;
        if (oy & 1) { y0 = iy; y1 = min(iy + 1, h - 1); ly = 0.25f; }
        else if (iy == 0) { y0 = 0; y1 = min(1, h - 1); ly = 0.f; }
        else { y0 = iy - 1; y1 = iy; ly = 0.75f; }
        const float hy = 1.f - ly;
        // up lanes: two low-res rows; skip lanes: the output row itself (ra = even pixels, rb = odd pixels)
        const T* ra = up_like ? sp + (int64_t)y0 * w * stride : sp + (int64_t)oy * W * stride;
        const T* rb = up_like ? sp + (int64_t)y1 * w * stride : ra + stride;
        const int64_t step = up_like ? stride : 2 * stride;   // element step of ra/rb per iteration
        float prev[8], cur[8], nxt[8], a[8], b[8];
        // column 0 (up) or pixels 0,1 (skip)
        Vec8<T>::ld(ra, a);
        Vec8<T>::ld(rb, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) { cur[i] = hy * a[i] + ly * b[i]; prev[i] = cur[i]; }
        float s0[8], s1[8];                                   // skip lanes: pixel pair of this step
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] = a[i]; s1[i] = b[i]; }
        // prefetch column 1 (up) / pixels 2,3 (skip)
        int jn = min(1, w - 1);
        Vec8<T>::ld(ra + (int64_t)jn * step, a);
        Vec8<T>::ld(rb + (int64_t)jn * step, b);
        T* yrow = y ? y + ((int64_t)(n * H + oy) * W) * Ctot + v8 * 8 : nullptr;
        float* lrow = sse_logit ? sse_logit + (int64_t)(n * H + oy) * W : nullptr;
        for (int i = 0; i < w; ++i) {
#pragma unroll
            for (int q = 0; q < 8; ++q) nxt[q] = hy * a[q] + ly * b[q];
            float n0[8], n1[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { n0[q] = a[q]; n1[q] = b[q]; }
            // issue the loads of the step after next before consuming this one
            const int j2 = min(i + 2, w - 1);
            Vec8<T>::ld(ra + (int64_t)j2 * step, a);
            Vec8<T>::ld(rb + (int64_t)j2 * step, b);
            float o0[8], o1[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float u0 = 0.25f * prev[q] + 0.75f * cur[q];
                const float u1 = 0.75f * cur[q] + 0.25f * nxt[q];
                o0[q] = is_up ? u0 : s0[q];
                o1[q] = is_up ? u1 : s1[q];
                prev[q] = cur[q];
                cur[q] = nxt[q];
                s0[q] = n0[q];
                s1[q] = n1[q];
            }
            if (live && yrow) {
                Vec8<T>::st(yrow + (int64_t)(2 * i) * Ctot, o0);
                Vec8<T>::st(yrow + (int64_t)(2 * i + 1) * Ctot, o1);
            }
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                d0 = fmaf(o0[q], wv[q], d0);
                d1 = fmaf(o1[q], wv[q], d1);
                acc[q] += live ? o0[q] + o1[q] : 0.f;
            }
            if (lrow) {
                d0 = warp_sum(d0);
                d1 = warp_sum(d1);
                if (lane == 0) {
                    if (multi) { atomicAdd(lrow + 2 * i, d0 + bias); atomicAdd(lrow + 2 * i + 1, d1 + bias); }
                    else { lrow[2 * i] = d0 + bias; lrow[2 * i + 1] = d1 + bias; }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_sum[warp][lane][i] = acc[i];
    __syncthreads();
    {
        const int vl = threadIdx.x >> 3, e = threadIdx.x & 7;      // 256 threads = 32 vectors x 8 elements
        const int ch = (blockIdx.z * 32 + vl) * 8 + e;
        if (ch < Ctot) {
            float t = 0.f;
            for (int wi = 0; wi < kStatsThreads / 32; ++wi) t += s_sum[wi][vl][e];
            atomicAdd(chan_mean + (int64_t)n * Ctot + ch, t * inv_hw);
        }
    }
}

// y = x * (cgate[n][c] + sigmoid(logit[n][p])), in place allowed.  grid-stride over (pixel, vec8).
template <typename T>
__global__ void __launch_bounds__(256)
scse_scale_kernel(const T* __restrict__ x, const float* __restrict__ cgate, const float* __restrict__ logit, int HW,
                  int C8, int64_t total, T* __restrict__ y) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = idx / C8;               // n * HW + p
        const int c8 = (int)(idx - pix * C8);
        const int n = (int)(pix / HW);
        const float s = sigmoidf_acc(__ldg(logit + pix));
        const float4* g4 = reinterpret_cast<const float4*>(cgate + (int64_t)n * C8 * 8 + c8 * 8);
        const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float v[8];
        Vec8<T>::ld(x + idx * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] * g[i] + v[i] * s;
        Vec8<T>::st(y + idx * 8, v);
    }
}

}  // namespace eds

using namespace eds;

extern "C" int eds_concat_stats(const void* x0, int N, int h, int w, int C0, int mode,
                                const void* const* skips_host, const int* skip_channels_host, int n_skips,
                                const float* w_sse, float b_sse, void* y, float* chan_mean, float* sse_logit,
                                int dtype, void* stream) {
    EDS_REQUIRE(x0 && chan_mean, "concat_stats: null pointer");
    EDS_REQUIRE(n_skips >= 0 && n_skips <= 5, "concat_stats: n_skips=%d not in 0..5", n_skips);
    EDS_REQUIRE(mode == EDS_UP_NEAREST || mode == EDS_UP_BILINEAR || mode == EDS_UP_NONE, "concat_stats: bad mode %d",
                mode);
    EDS_REQUIRE(N > 0 && N <= 65535 && h > 0 && w > 0 && C0 > 0 && C0 % 8 == 0, "concat_stats: bad shape");
    EDS_REQUIRE(!sse_logit || w_sse, "concat_stats: sse_logit requested without w_sse");
    StatSrc src;
    src.n = n_skips + 1;
    src.ptr[0] = x0;
    src.ch[0] = C0;
    int Ctot = C0;
    for (int k = 0; k < 5; ++k) {
        src.ptr[k + 1] = k < n_skips ? skips_host[k] : nullptr;
        src.ch[k + 1] = k < n_skips ? skip_channels_host[k] : 0;
        if (k < n_skips) {
            EDS_REQUIRE(src.ptr[k + 1] && src.ch[k + 1] > 0 && src.ch[k + 1] % 8 == 0, "concat_stats: bad skip %d", k);
            Ctot += src.ch[k + 1];
        }
    }
    const int up = mode == EDS_UP_NONE ? 1 : 2;
    const int n_pix = up * h * up * w;
    const int c8 = Ctot / 8;
    int lpp = 32;                                  // lanes per pixel: pow2 >= Ctot/8, capped at a warp
    while (lpp > 1 && lpp / 2 >= c8) lpp /= 2;
    const int zchunks = ceil_div(c8, 32);
    EDS_REQUIRE(zchunks <= 65535, "concat_stats: too many channels");
    cudaError_t e = cudaMemsetAsync(chan_mean, 0, sizeof(float) * (size_t)N * Ctot, as_stream(stream));
    if (e == cudaSuccess && sse_logit && zchunks > 1)
        e = cudaMemsetAsync(sse_logit, 0, sizeof(float) * (size_t)N * n_pix, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("concat_stats: memset failed: %s", cudaGetErrorString(e));
        return EDS_ERR_CUDA;
    }
    if (mode == EDS_UP_BILINEAR && c8 >= 32) {
        dim3 grid(ceil_div(2 * h, kStatsThreads / 32), N, zchunks);
        EDS_DISPATCH_DTYPE(dtype, T, (concat_stats_bilinear_kernel<T><<<grid, kStatsThreads, 0, as_stream(stream)>>>(
                                         src, h, w, Ctot, w_sse, b_sse, 1.0f / (float)n_pix, (T*)y, chan_mean,
                                         sse_logit)));
        return check_launch("concat_stats_bilinear_kernel");
    }
    int chunks = ceil_div(148 * 16, N * zchunks);
    const int max_chunks = ceil_div(n_pix, 8 * (32 / lpp) * kStatsUnroll * 4);   // >= 4 steps per warp
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, N, zchunks);
    EDS_DISPATCH_DTYPE(dtype, T, (concat_stats_kernel<T><<<grid, kStatsThreads, 0, as_stream(stream)>>>(
                                     src, h, w, mode, Ctot, lpp, w_sse, b_sse, 1.0f / (float)n_pix, (T*)y, chan_mean,
                                     sse_logit)));
    return check_launch("concat_stats_kernel");
}

extern "C" int eds_scse_scale(const void* x, const float* cgate, const float* sse_logit, int N, int HW, int C,
                              void* y, int dtype, void* stream) {
    EDS_REQUIRE(x && cgate && sse_logit && y, "scse_scale: null pointer");
    EDS_REQUIRE(C % 8 == 0 && C > 0 && N > 0 && HW > 0, "scse_scale: bad shape");
    const int64_t total = (int64_t)N * HW * (C / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    EDS_DISPATCH_DTYPE(dtype, T, (scse_scale_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
                                     (const T*)x, cgate, sse_logit, HW, C / 8, total, (T*)y)));
    return check_launch("scse_scale_kernel");
}
