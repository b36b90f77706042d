// Bandwidth-bound NHWC kernels between the convolutions: pools, squeeze-excite,
// SCSE, decoder upsample+concat, MHCA gate, segmentation head, casts.
// Reference call sites are cited per kernel.  All kernels move 8-channel vectors
// (16 B bf16 / 32 B fp32) and do their arithmetic in fp32.
#include "common.cuh"
#include <cfloat>

namespace eds {

// ---------------------------------------------------------------- max pool
// SENet layer0.pool = MaxPool2d(3, 2, ceil_mode=True) applied at unetplusplusstar.py:347-348,
// torchvision ResNet maxpool (3,2,pad 1), MHCA init_conv MaxPool2d(2) unetplusplusstar.py:106.
template <typename T, int K>
__global__ void __launch_bounds__(256)
maxpool_kernel(const T* __restrict__ x, int N, int H, int W, int C8, int stride, int pad, int Ho, int Wo,
               T* __restrict__ y) {
    // Window coordinates are clamped into the map instead of skipped: a clamped tap repeats a pixel of
    // the same window (pad < K and the ceil_mode rule keep one row / column of every window inside), so
    // the maximum is unchanged and all K*K vector loads are issued back to back.
    const int64_t total = (int64_t)N * Ho * Wo * C8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(idx % C8);
        int64_t r = idx / C8;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int n = (int)(r / Ho);
        float v[K * K][8];
#pragma unroll
        for (int dy = 0; dy < K; ++dy) {
            const int iy = min(max(oh * stride - pad + dy, 0), H - 1);
#pragma unroll
            for (int dx = 0; dx < K; ++dx) {
                const int ix = min(max(ow * stride - pad + dx, 0), W - 1);
                Vec8<T>::ld(x + (((int64_t)n * H + iy) * W + ix) * C8 * 8 + c8 * 8, v[dy * K + dx]);
            }
        }
        float best[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) best[i] = v[0][i];
#pragma unroll
        for (int q = 1; q < K * K; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) best[i] = fmaxf(best[i], v[q][i]);
        Vec8<T>::st(y + idx * 8, best);
    }
}

// ------------------------------------------------ avgpool(2) + affine (+relu)
// axial_attention_v2.py:256-259 (att_down = AvgPool2d(2) + BN) followed by the ReLU of :279.
template <typename T>
__global__ void avgpool2_affine_kernel(const T* __restrict__ x, int N, int H, int W, int C8,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       int relu, T* __restrict__ y) {
    const int Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)N * Ho * Wo * C8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(idx % C8);
        int64_t r = idx / C8;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int n = (int)(r / Ho);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float v[8];
                Vec8<T>::ld(x + (((int64_t)n * H + 2 * oh + dy) * W + 2 * ow + dx) * C8 * 8 + c8 * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += v[i];
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float o = acc[i] * 0.25f * scale[c8 * 8 + i] + shift[c8 * 8 + i];
            acc[i] = relu ? fmaxf(o, 0.f) : o;
        }
        Vec8<T>::st(y + idx * 8, acc);
    }
}

// -------------------------------------------------------- global average pool
// AdaptiveAvgPool2d(1) inside the SENet SE module and smp SCSEModule.cSE.
// grid = (pixel slices, N, channel chunks); CTA = 256 threads = 8 warps; inside a warp
// `lp` lanes cover one pixel's channel chunk (lp*8 channels), 32/lp pixels per warp.
template <typename T>
__global__ void __launch_bounds__(256)
channel_mean_kernel(const T* __restrict__ x, int HW, int C, int lp, float inv_hw, float* __restrict__ mean) {
    __shared__ float red[256][9];
    const int tid = threadIdx.x;
    const int lane_c = tid % lp;           // vec8 slot inside the chunk
    const int prow = tid / lp;             // pixel row handled by this thread
    const int rows = 256 / lp;
    const int n = blockIdx.y;
    const int c0 = (blockIdx.z * lp + lane_c) * 8;
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * per;
    const int p_end = min(HW, p_begin + per);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < C) {
        const T* base = x + (int64_t)n * HW * C + c0;
        for (int p = p_begin + prow; p < p_end; p += rows) {
            float v[8];
            Vec8<T>::ld(base + (int64_t)p * C, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[tid][i] = acc[i];
    __syncthreads();
    if (tid < lp && c0 < C) {
        float tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < rows; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) tot[i] += red[r * lp + tid][i];
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(mean + (int64_t)n * C + c0 + i, tot[i] * inv_hw);
    }
}

// ------------------------------------------------------------------ SE gate
// SENet SEModule fc1/relu/fc2/sigmoid and smp SCSEModule.cSE (both 1x1 convs on a 1x1 map).
__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ mean, int C, int Cr, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
               float* __restrict__ gate) {
    extern __shared__ float sm[];
    float* s_in = sm;        // [C]
    float* s_hid = sm + C;   // [Cr]
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < C; c += 256) s_in[c] = mean[(int64_t)n * C + c];
    __syncthreads();
    for (int j = warp; j < Cr; j += 8) {
        const float* wr = w1 + (int64_t)j * C;
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a += wr[c] * s_in[c];
        a = warp_sum(a);
        if (lane == 0) s_hid[j] = fmaxf(a + b1[j], 0.f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        const float* wr = w2 + (int64_t)c * Cr;
        float a = b2[c];
        for (int j = 0; j < Cr; ++j) a += wr[j] * s_hid[j];
        gate[(int64_t)n * C + c] = sigmoidf_acc(a);
    }
}

// ------------------------------------------- y = relu(x * gate + residual)
// Tail of SEBottleneck.forward: out = se_module(out) + residual; relu.
template <typename T>
__global__ void se_scale_add_relu_kernel(const T* __restrict__ x, const float* __restrict__ gate,
                                         const T* __restrict__ res, int64_t HW, int C8, int64_t total,
                                         T* __restrict__ y) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(idx % C8);
        const int64_t n = idx / (HW * C8);
        float v[8], r[8];
        Vec8<T>::ld(x + idx * 8, v);
        Vec8<T>::ld(res + idx * 8, r);
        const float* g = gate + n * C8 * 8 + c8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] * g[i] + r[i], 0.f);
        Vec8<T>::st(y + idx * 8, v);
    }
}

// --------------------------------------------------- upsample x2 + concat
// unetplusplusstar.py:128,153 (bilinear) / deep_supunetplusplus.py:49-51 (nearest) + torch.cat.
struct ConcatSrc {
    const void* ptr[5];
    int ch[5];
    int n;
};

template <typename T>
__device__ __forceinline__ void up2x_bilinear_vec(const T* __restrict__ x0, int n, int h, int w, int C0, int oy,
                                                  int ox, int c, float (&out)[8]) {
    // align_corners=False, scale 2: src = o/2 - 0.25, clamped at 0; neighbour clamped at size-1
    const int iy = oy >> 1, ix = ox >> 1;
    int y0, y1, x0i, x1i;
    float ly, lx;
    if (oy & 1) { y0 = iy; y1 = min(iy + 1, h - 1); ly = 0.25f; }
    else if (iy == 0) { y0 = 0; y1 = min(1, h - 1); ly = 0.f; }
    else { y0 = iy - 1; y1 = iy; ly = 0.75f; }
    if (ox & 1) { x0i = ix; x1i = min(ix + 1, w - 1); lx = 0.25f; }
    else if (ix == 0) { x0i = 0; x1i = min(1, w - 1); lx = 0.f; }
    else { x0i = ix - 1; x1i = ix; lx = 0.75f; }
    const T* base = x0 + (int64_t)n * h * w * C0 + c;
    float v00[8], v01[8], v10[8], v11[8];
    Vec8<T>::ld(base + ((int64_t)y0 * w + x0i) * C0, v00);
    Vec8<T>::ld(base + ((int64_t)y0 * w + x1i) * C0, v01);
    Vec8<T>::ld(base + ((int64_t)y1 * w + x0i) * C0, v10);
    Vec8<T>::ld(base + ((int64_t)y1 * w + x1i) * C0, v11);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = hy * (hx * v00[i] + lx * v01[i]) + ly * (hx * v10[i] + lx * v11[i]);
}

template <typename T>
__global__ void upsample2x_concat_kernel(const T* __restrict__ x0, int N, int h, int w, int C0, int mode,
                                         ConcatSrc skips, int Ctot, T* __restrict__ y) {
    const int up = mode == EDS_UP_NONE ? 1 : 2;
    const int H = up * h, W = up * w, Ct8 = Ctot / 8;
    const int64_t total = (int64_t)N * H * W * Ct8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % Ct8) * 8;
        int64_t r = idx / Ct8;
        const int ox = (int)(r % W); r /= W;
        const int oy = (int)(r % H);
        const int n = (int)(r / H);
        float v[8];
        if (c < C0) {
            if (mode == EDS_UP_BILINEAR) up2x_bilinear_vec<T>(x0, n, h, w, C0, oy, ox, c, v);
            else if (mode == EDS_UP_NEAREST) Vec8<T>::ld(x0 + (((int64_t)n * h + (oy >> 1)) * w + (ox >> 1)) * C0 + c, v);
            else Vec8<T>::ld(x0 + (((int64_t)n * h + oy) * w + ox) * C0 + c, v);
        } else {
            c -= C0;
            int k = 0;
            while (k < skips.n - 1 && c >= skips.ch[k]) { c -= skips.ch[k]; ++k; }
            const T* sp = reinterpret_cast<const T*>(skips.ptr[k]);
            Vec8<T>::ld(sp + (((int64_t)n * H + oy) * W + ox) * skips.ch[k] + c, v);
        }
        Vec8<T>::st(y + idx * 8, v);
    }
}

// ----------------------------------------------------------------- MHCA gate
// unetplusplusstar.py:146-147: skip = up_scale(skip) = Upsample(x2, bilinear)(Sigmoid(skip)); ori_skip * skip.
template <typename T>
__global__ void mhca_gate_kernel(const T* __restrict__ ori, const T* __restrict__ att, int N, int h, int w, int C,
                                 T* __restrict__ y) {
    const int H = 2 * h, W = 2 * w, C8 = C / 8;
    const int64_t total = (int64_t)N * H * W * C8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C8) * 8;
        int64_t r = idx / C8;
        const int ox = (int)(r % W); r /= W;
        const int oy = (int)(r % H);
        const int n = (int)(r / H);
        const int iy = oy >> 1, ix = ox >> 1;
        int y0, y1, x0i, x1i;
        float ly, lx;
        if (oy & 1) { y0 = iy; y1 = min(iy + 1, h - 1); ly = 0.25f; }
        else if (iy == 0) { y0 = 0; y1 = min(1, h - 1); ly = 0.f; }
        else { y0 = iy - 1; y1 = iy; ly = 0.75f; }
        if (ox & 1) { x0i = ix; x1i = min(ix + 1, w - 1); lx = 0.25f; }
        else if (ix == 0) { x0i = 0; x1i = min(1, w - 1); lx = 0.f; }
        else { x0i = ix - 1; x1i = ix; lx = 0.75f; }
        const T* base = att + (int64_t)n * h * w * C + c;
        float v00[8], v01[8], v10[8], v11[8], o[8];
        Vec8<T>::ld(base + ((int64_t)y0 * w + x0i) * C, v00);
        Vec8<T>::ld(base + ((int64_t)y0 * w + x1i) * C, v01);
        Vec8<T>::ld(base + ((int64_t)y1 * w + x0i) * C, v10);
        Vec8<T>::ld(base + ((int64_t)y1 * w + x1i) * C, v11);
        Vec8<T>::ld(ori + idx * 8, o);
        const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g = hy * (hx * sigmoidf_acc(v00[i]) + lx * sigmoidf_acc(v01[i])) +
                            ly * (hx * sigmoidf_acc(v10[i]) + lx * sigmoidf_acc(v11[i]));
            o[i] *= g;
        }
        Vec8<T>::st(y + idx * 8, o);
    }
}

// ----------------------------------------------------------- segmentation head
// SegmentationHead conv2d 3x3 pad 1 with bias (unetplusplusstar.py:163-168, :484); C <= 64.
template <typename T>
__global__ void __launch_bounds__(256)
head_conv3x3_kernel(const T* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ w,
                    const float* __restrict__ bias, int classes, float* __restrict__ logits) {
    extern __shared__ __align__(16) float s_w[];  // [classes][9][C]
    for (int i = threadIdx.x; i < classes * 9 * C; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
    const int64_t total = (int64_t)N * H * W;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int ox = (int)(p % W);
        const int oy = (int)((p / W) % H);
        const int n = (int)(p / ((int64_t)W * H));
        // taps outside the map read a clamped (valid) pixel and are weighted by 0: every load is
        // unconditional, so the nine taps of a pixel are in flight together
        const T* xp[9];
        float valid[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int iy = oy + dy - 1, ix = ox + dx - 1;
                valid[dy * 3 + dx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? 1.f : 0.f;
                xp[dy * 3 + dx] = x + (((int64_t)n * H + min(max(iy, 0), H - 1)) * W + min(max(ix, 0), W - 1)) * C;
            }
        for (int k = 0; k < classes; ++k) {
            float acc = bias ? bias[k] : 0.f;
            for (int c = 0; c < C; c += 8) {
                float2 v[9][4];
#pragma unroll
                for (int q = 0; q < 9; ++q) V8<T>::ld(xp[q] + c, v[q]);
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    // 8 weights per two 128-bit shared loads (one scalar load per multiply made this kernel
                    // LDS-bound); the products run on the packed fp32 pipe, two per instruction
                    const float4* wp = reinterpret_cast<const float4*>(s_w + (k * 9 + q) * C + c);
                    const float4 w0 = wp[0], w1 = wp[1];
                    float2 d = __fmul2_rn(v[q][0], make_float2(w0.x, w0.y));
                    d = __ffma2_rn(v[q][1], make_float2(w0.z, w0.w), d);
                    d = __ffma2_rn(v[q][2], make_float2(w1.x, w1.y), d);
                    d = __ffma2_rn(v[q][3], make_float2(w1.z, w1.w), d);
                    acc = fmaf(valid[q], d.x + d.y, acc);
                }
            }
            logits[(((int64_t)n * classes + k) * H + oy) * W + ox] = acc;
        }
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = __bfloat162float(x[i]);
}

static inline int grid_for(int64_t total, int block) {
    int64_t g = (total + block - 1) / block;
    const int64_t cap = 148 * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static inline int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// out[n][co] = w[co][:] . m[n][:] + b[co]: one warp per output, lanes along Cin (the SE squeeze of conv3's output
// from the channel means of its input; N * Cout <= 48 * 1024 dots of length <= 512).
__global__ void __launch_bounds__(256)
affine_rows_kernel(const float* __restrict__ m, int Cin, const float* __restrict__ w, const float* __restrict__ b,
                   int Cout, int total, float* __restrict__ out) {
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (idx >= total) return;
    const int n = idx / Cout, co = idx - n * Cout;
    float acc = 0.f;
    for (int ci = lane; ci < Cin; ci += 32) acc = fmaf(__ldg(w + (int64_t)co * Cin + ci), __ldg(m + (int64_t)n * Cin + ci), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[idx] = acc + (b ? b[co] : 0.f);
}

}  // namespace eds

using namespace eds;

extern "C" int eds_maxpool2d(const void* x, int N, int H, int W, int C, int k, int stride, int pad, int ceil_mode,
                             void* y, int dtype, void* stream) {
    EDS_REQUIRE(x && y, "maxpool2d: null pointer");
    EDS_REQUIRE(C % 8 == 0 && C > 0, "maxpool2d: C=%d must be a multiple of 8", C);
    EDS_REQUIRE(k >= 1 && stride >= 1 && pad >= 0 && pad <= k / 2, "maxpool2d: bad k/stride/pad");
    auto out_size = [&](int in) {
        int num = in + 2 * pad - k;
        int o = (ceil_mode ? (num + stride - 1) / stride : num / stride) + 1;
        if (ceil_mode && (o - 1) * stride >= in + pad) --o;
        return o;
    };
    const int Ho = out_size(H), Wo = out_size(W);
    EDS_REQUIRE(Ho > 0 && Wo > 0, "maxpool2d: empty output");
    const int64_t total = (int64_t)N * Ho * Wo * (C / 8);
    EDS_REQUIRE(k == 2 || k == 3, "maxpool2d: k=%d (2 and 3 are on the path)", k);
    EDS_DISPATCH_DTYPE(dtype, T, {
        if (k == 2)
            maxpool_kernel<T, 2><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const T*)x, N, H, W, C / 8, stride,
                                                                                     pad, Ho, Wo, (T*)y);
        else
            maxpool_kernel<T, 3><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const T*)x, N, H, W, C / 8, stride,
                                                                                     pad, Ho, Wo, (T*)y);
    });
    return check_launch("maxpool_kernel");
}

extern "C" int eds_avgpool2_affine(const void* x, int N, int H, int W, int C, const float* scale,
                                   const float* shift, int relu, void* y, int dtype, void* stream) {
    EDS_REQUIRE(x && y && scale && shift, "avgpool2_affine: null pointer");
    EDS_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool2_affine: C%%8, H%%2, W%%2 must be 0");
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
    EDS_DISPATCH_DTYPE(dtype, T, (avgpool2_affine_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
                                     (const T*)x, N, H, W, C / 8, scale, shift, relu, (T*)y)));
    return check_launch("avgpool2_affine_kernel");
}

extern "C" int eds_channel_mean(const void* x, int N, int HW, int C, float* mean, int dtype, void* stream) {
    EDS_REQUIRE(x && mean, "channel_mean: null pointer");
    EDS_REQUIRE(C % 8 == 0 && C > 0 && HW > 0 && N > 0 && N <= 65535, "channel_mean: bad shape N=%d HW=%d C=%d", N,
                HW, C);
    const int c8 = C / 8;
    const int lp = c8 >= 32 ? 32 : pow2_floor(c8);  // lanes per pixel
    const int chunks = ceil_div(c8, lp);
    int slices = ceil_div(HW, (256 / lp) * 16);     // >= 16 pixels per thread row
    const int cap = ceil_div(148 * 8, N * chunks);
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    cudaError_t e = cudaMemsetAsync(mean, 0, sizeof(float) * (size_t)N * C, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("channel_mean: memset failed: %s", cudaGetErrorString(e));
        return EDS_ERR_CUDA;
    }
    dim3 grid(slices, N, chunks);
    EDS_DISPATCH_DTYPE(dtype, T, (channel_mean_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(
                                     (const T*)x, HW, C, lp, 1.0f / (float)HW, mean)));
    return check_launch("channel_mean_kernel");
}

extern "C" int eds_se_gate(const float* mean, int N, int C, int Cr, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* gate, void* stream) {
    EDS_REQUIRE(mean && w1 && b1 && w2 && b2 && gate, "se_gate: null pointer");
    EDS_REQUIRE(N > 0 && C > 0 && Cr > 0 && (size_t)(C + Cr) * 4 <= 48 * 1024, "se_gate: bad sizes C=%d Cr=%d", C, Cr);
    se_gate_kernel<<<N, 256, (C + Cr) * sizeof(float), as_stream(stream)>>>(mean, C, Cr, w1, b1, w2, b2, gate);
    return check_launch("se_gate_kernel");
}

extern "C" int eds_se_scale_add_relu(const void* x, const float* gate, const void* residual, int N, int HW, int C,
                                     void* y, int dtype, void* stream) {
    EDS_REQUIRE(x && gate && residual && y, "se_scale_add_relu: null pointer");
    EDS_REQUIRE(C % 8 == 0, "se_scale_add_relu: C=%d must be a multiple of 8", C);
    const int64_t total = (int64_t)N * HW * (C / 8);
    EDS_DISPATCH_DTYPE(dtype, T, (se_scale_add_relu_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
                                     (const T*)x, gate, (const T*)residual, HW, C / 8, total, (T*)y)));
    return check_launch("se_scale_add_relu_kernel");
}

extern "C" int eds_affine_rows(const float* m, int N, int Cin, const float* w, const float* b, int Cout, float* out,
                               void* stream) {
    EDS_REQUIRE(m && w && out && N > 0 && Cin > 0 && Cout > 0, "affine_rows: bad arguments");
    const int total = N * Cout;
    affine_rows_kernel<<<ceil_div(total, 8), 256, 0, as_stream(stream)>>>(m, Cin, w, b, Cout, total, out);
    return check_launch("affine_rows_kernel");
}

extern "C" int eds_upsample2x_concat(const void* x0, int N, int h, int w, int C0, int mode,
                                     const void* const* skips_host, const int* skip_channels_host, int n_skips,
                                     void* y, int dtype, void* stream) {
    EDS_REQUIRE(x0 && y, "upsample2x_concat: null pointer");
    EDS_REQUIRE(n_skips >= 0 && n_skips <= 5, "upsample2x_concat: n_skips=%d not in 0..5", n_skips);
    EDS_REQUIRE(C0 % 8 == 0 && C0 > 0, "upsample2x_concat: C0=%d must be a multiple of 8", C0);
    EDS_REQUIRE(mode == EDS_UP_NEAREST || mode == EDS_UP_BILINEAR || mode == EDS_UP_NONE,
                "upsample2x_concat: bad mode %d", mode);
    ConcatSrc src;
    src.n = n_skips;
    int Ctot = C0;
    for (int k = 0; k < 5; ++k) {
        src.ptr[k] = k < n_skips ? skips_host[k] : nullptr;
        src.ch[k] = k < n_skips ? skip_channels_host[k] : 0;
        if (k < n_skips) {
            EDS_REQUIRE(src.ptr[k] && src.ch[k] > 0 && src.ch[k] % 8 == 0, "upsample2x_concat: bad skip %d", k);
            Ctot += src.ch[k];
        }
    }
    const int up = mode == EDS_UP_NONE ? 1 : 2;
    const int64_t total = (int64_t)N * up * h * up * w * (Ctot / 8);
    EDS_DISPATCH_DTYPE(dtype, T, (upsample2x_concat_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
                                     (const T*)x0, N, h, w, C0, mode, src, Ctot, (T*)y)));
    return check_launch("upsample2x_concat_kernel");
}

extern "C" int eds_mhca_gate(const void* ori, const void* att, int N, int h, int w, int C, void* y, int dtype,
                             void* stream) {
    EDS_REQUIRE(ori && att && y, "mhca_gate: null pointer");
    EDS_REQUIRE(C % 8 == 0 && C > 0, "mhca_gate: C=%d must be a multiple of 8", C);
    const int64_t total = (int64_t)N * 4 * h * w * (C / 8);
    EDS_DISPATCH_DTYPE(dtype, T, (mhca_gate_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
                                     (const T*)ori, (const T*)att, N, h, w, C, (T*)y)));
    return check_launch("mhca_gate_kernel");
}

extern "C" int eds_head_conv3x3(const void* x, int N, int H, int W, int C, const float* w, const float* bias,
                                int classes, float* logits, int dtype, void* stream) {
    EDS_REQUIRE(x && w && logits, "head_conv3x3: null pointer");
    EDS_REQUIRE(C % 8 == 0 && C > 0 && C <= 64 && classes >= 1 && classes <= 16, "head_conv3x3: C=%d classes=%d", C,
                classes);
    const int64_t total = (int64_t)N * H * W;
    const size_t smem = sizeof(float) * classes * 9 * C;
    EDS_DISPATCH_DTYPE(dtype, T, (head_conv3x3_kernel<T><<<grid_for(total, 256), 256, smem, as_stream(stream)>>>(
                                     (const T*)x, N, H, W, C, w, bias, classes, logits)));
    return check_launch("head_conv3x3_kernel");
}

extern "C" int eds_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
    EDS_REQUIRE(x && y && n >= 0, "cast: bad args");
    if (n == 0) return EDS_OK;
    cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)y, n);
    return check_launch("cast_f32_bf16_kernel");
}
extern "C" int eds_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
    EDS_REQUIRE(x && y && n >= 0, "cast: bad args");
    if (n == 0) return EDS_OK;
    cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, y, n);
    return check_launch("cast_bf16_f32_kernel");
}
