// Encoder stem: 7x7 stride-2 pad-3 convolution 3 -> 64 with folded BatchNorm + ReLU,
// reading the reference's fp32 NCHW input directly and folding the TTA view
// (ttach HorizontalFlip / VerticalFlip / Rotate90, src/main/tta.py:92-99) into the
// loader's coordinates so augmented copies of the input are never materialised.
//
// Reference: SENet layer0 (conv1/bn1/relu1) as used by BoTSER50.forward
// (src/main/archs/unetplusplusstar.py:341-352) and the smp ResNet/SENet encoders.
// Cin = 3 makes this a CUDA-core kernel (K = 147 is too ragged for a UMMA tile);
// it is ~0.3 % of the network's FLOPs.
#include "common.cuh"
#include <cstdlib>

namespace eds {

struct StemViews {
    int m[8][6];
};

constexpr int kStemTile = 16;                    // output pixels per tile edge
constexpr int kStemPatch = kStemTile * 2 + 5;    // 37 input pixels per edge
constexpr int kStemK = 7 * 7 * 3;                // 147

// smem: s_w[147][64] fp32 (cout innermost) + s_in[3][37][37] fp32
template <typename T>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ x, int B, int H, int W, StemViews views, const float* __restrict__ w,
                 const float* __restrict__ bias, T* __restrict__ y) {
    extern __shared__ float sm[];
    float* s_w = sm;
    float* s_in = sm + kStemK * 64;
    const int tid = threadIdx.x;
    const int Ho = H / 2, Wo = W / 2;
    const int img = blockIdx.z;         // v * B + b
    const int v = img / B, b = img % B;
    const int oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;

    // w is [7][7][3][64] (cout innermost) = s_w[(r*7+s)*3+c][cout]
    for (int i = tid; i < 64 * kStemK; i += 256) s_w[i] = __ldg(w + i);
    const int* m = views.m[v];
    const float* xb = x + (int64_t)b * 3 * H * W;
    for (int i = tid; i < 3 * kStemPatch * kStemPatch; i += 256) {
        const int c = i / (kStemPatch * kStemPatch);
        const int rem = i % (kStemPatch * kStemPatch);
        const int py = rem / kStemPatch, px = rem % kStemPatch;
        const int iy = 2 * oy0 - 3 + py, ix = 2 * ox0 - 3 + px;  // coordinates in the augmented view
        float val = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            const int sy = m[0] * iy + m[1] * ix + m[2];
            const int sx = m[3] * iy + m[4] * ix + m[5];
            val = __ldg(xb + ((int64_t)c * H + sy) * W + sx);
        }
        s_in[i] = val;
    }
    __syncthreads();

    const int cg = tid >> 6;            // 16-channel group
    const int pg = tid & 63;
    const int row = pg >> 2;            // output row inside the tile
    const int col0 = (pg & 3) * 4;      // first of 4 output columns
    // 4 pixels x 16 couts per thread as 4 x 8 float2: the packed fp32 pipe (FFMA2) retires two of the
    // 118 G multiply-adds of a 48-map batch per instruction.
    float2 acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[p][q] = make_float2(0.f, 0.f);

    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 7; ++r) {
            const float* in_row = s_in + (c * kStemPatch + 2 * row + r) * kStemPatch + 2 * col0;
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const float4* wv = reinterpret_cast<const float4*>(s_w + ((r * 7 + s) * 3 + c) * 64 + cg * 16);
                const float4 w0 = wv[0], w1 = wv[1], w2 = wv[2], w3 = wv[3];
                const float2 wk[8] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                      make_float2(w1.z, w1.w), make_float2(w2.x, w2.y), make_float2(w2.z, w2.w),
                                      make_float2(w3.x, w3.y), make_float2(w3.z, w3.w)};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float a = in_row[2 * p + s];
                    const float2 a2 = make_float2(a, a);
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[p][q] = __ffma2_rn(a2, wk[q], acc[p][q]);
                }
            }
        }

    const int oy = oy0 + row;
    if (oy < Ho) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ox = ox0 + col0 + p;
            if (ox >= Wo) continue;
            T* yp = y + (((int64_t)img * Ho + oy) * Wo + ox) * 64 + cg * 16;
            float o0[8], o1[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o0[2 * q] = fmaxf(acc[p][q].x + bias[cg * 16 + 2 * q], 0.f);
                o0[2 * q + 1] = fmaxf(acc[p][q].y + bias[cg * 16 + 2 * q + 1], 0.f);
                o1[2 * q] = fmaxf(acc[p][q + 4].x + bias[cg * 16 + 8 + 2 * q], 0.f);
                o1[2 * q + 1] = fmaxf(acc[p][q + 4].y + bias[cg * 16 + 8 + 2 * q + 1], 0.f);
            }
            Vec8<T>::st(yp, o0);
            Vec8<T>::st(yp + 8, o1);
        }
    }
}

int stem_conv_mma_launch(const float* x, int B, int H, int W, int V, const int* maps, const void* w_packed,
                         const float* bias, void* y, cudaStream_t stream);   // stem_mma.cu
int stem_pack_launch(const float* w, void* out, cudaStream_t stream);

}  // namespace eds

using namespace eds;

static int stem_check_views(int B, int V, int H, int W, const int* aug_maps_host, StemViews* views) {
    EDS_REQUIRE(B >= 1 && V >= 1 && V <= 8 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0,
                "stem_conv: bad shape B=%d V=%d H=%d W=%d", B, V, H, W);
    EDS_REQUIRE((int64_t)B * V <= 65535, "stem_conv: B*V too large");
    for (int v = 0; v < V; ++v) {
        const int* m = aug_maps_host + v * 6;
        const bool straight = m[1] == 0 && m[3] == 0 && (m[0] == 1 || m[0] == -1) && (m[4] == 1 || m[4] == -1);
        const bool swapped = m[0] == 0 && m[4] == 0 && (m[1] == 1 || m[1] == -1) && (m[3] == 1 || m[3] == -1);
        EDS_REQUIRE(straight || swapped, "stem_conv: view %d map is not a flip/rot90", v);
        EDS_REQUIRE(!swapped || H == W, "stem_conv: rot90 views need a square input (H=%d W=%d)", H, W);
        for (int corner = 0; corner < 4; ++corner) {
            const int i = (corner & 1) ? H - 1 : 0, j = (corner & 2) ? W - 1 : 0;
            const int r = m[0] * i + m[1] * j + m[2], c = m[3] * i + m[4] * j + m[5];
            EDS_REQUIRE(r >= 0 && r < H && c >= 0 && c < W, "stem_conv: view %d map leaves the image", v);
        }
        for (int q = 0; q < 6; ++q) views->m[v][q] = m[q];
    }
    return EDS_OK;
}

extern "C" int eds_stem_conv7x7s2(const float* x, int B, int H, int W, int V, const int* aug_maps_host,
                                  const float* w, const float* bias, void* y, int dtype, void* stream) {
    EDS_REQUIRE(x && w && bias && y && aug_maps_host, "stem_conv: null pointer");
    StemViews views;
    if (int rc = stem_check_views(B, V, H, W, aug_maps_host, &views)) return rc;
    const size_t smem = sizeof(float) * (kStemK * 64 + 3 * kStemPatch * kStemPatch);
    dim3 grid(ceil_div(W / 2, kStemTile), ceil_div(H / 2, kStemTile), B * V);
    cudaError_t e = cudaSuccess;
    EDS_DISPATCH_DTYPE(dtype, T, {
        e = cudaFuncSetAttribute(stem_conv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            stem_conv_kernel<T><<<grid, 256, smem, as_stream(stream)>>>(x, B, H, W, views, w, bias, (T*)y);
    });
    if (e != cudaSuccess) {
        set_error("stem_conv: shared-memory opt-in failed: %s", cudaGetErrorString(e));
        return EDS_ERR_CUDA;
    }
    return check_launch("stem_conv_kernel");
}

extern "C" int eds_stem_pack_weights(const float* w, void* w_packed, void* stream) {
    EDS_REQUIRE(w && w_packed, "stem_pack_weights: null pointer");
    EDS_REQUIRE(((uintptr_t)w_packed & 15) == 0, "stem_pack_weights: output must be 16-byte aligned");
    return stem_pack_launch(w, w_packed, as_stream(stream));
}

extern "C" int eds_stem_conv7x7s2_mma(const float* x, int B, int H, int W, int V, const int* aug_maps_host,
                                      const void* w_packed, const float* bias, void* y, void* stream) {
    EDS_REQUIRE(x && w_packed && bias && y && aug_maps_host, "stem_conv_mma: null pointer");
    EDS_REQUIRE((((uintptr_t)w_packed | (uintptr_t)y) & 15) == 0, "stem_conv_mma: pointers must be 16-byte aligned");
    StemViews views;
    if (int rc = stem_check_views(B, V, H, W, aug_maps_host, &views)) return rc;
    return stem_conv_mma_launch(x, B, H, W, V, aug_maps_host, w_packed, bias, y, as_stream(stream));
}
