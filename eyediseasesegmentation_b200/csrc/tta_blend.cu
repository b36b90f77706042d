// TTA de-augment + mean + sigmoid, bilinear paste, and sliding-window tile fetch.
//
// Replaces (reference file:line):
//   - ttach.SegmentationTTAWrapper de-augmentation and Merger('mean') as used at
//     src/main/tta.py:92-99,173-180, followed by the sigmoid of tta.py:114,210;
//   - cv2.resize(..., INTER_LINEAR) + `preds[x1:x2, y1:y2] = tile` (tta.py:211-213)
//     and center_crop + GF.resize (tta.py:117-119);
//   - dataset.read(window) + A.Resize + preprocessing_fn + ToTensorV2 (tta.py:201-204).
// All three are HBM-bound streaming kernels; the flip / rot90 views are never
// materialised, they are index maps applied while reading.
#include "common.cuh"

namespace eds {

struct ViewMaps {
    int m[8][6];
};

// One CTA (32 x 8 threads) = one 32x32 output tile of one image; a thread owns 4 rows of its column.
// All 4 x V loads of a thread are issued before the first use (addresses clamped into the tile, so no
// load is conditional) -- 32 x 4 B in flight per thread; views whose map transposes the axes are read
// along THEIR rows (coalesced) and turned through one shared-memory tile per view.  The sum runs in
// view order in fp32, like ttach's Merger.
__global__ void __launch_bounds__(256)
tta_merge_kernel(const float* __restrict__ logits, int V, int B, int S, ViewMaps maps, int apply_sigmoid,
                 float* __restrict__ prob) {
    __shared__ float tile[8][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 8
    const int b = blockIdx.z;
    const int j = blockIdx.x * 32 + tx;                 // output col
    const int jc = min(j, S - 1);
    float val[4][8];
    bool any_t = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int rr = ty + 8 * r;                       // row inside the tile
        const int ic = min(blockIdx.y * 32 + rr, S - 1);
        // transposing views: thread (rr,tx) fetches the value of output (row tx, col rr) of the tile
        const int io = min(blockIdx.y * 32 + tx, S - 1), jo = min(blockIdx.x * 32 + rr, S - 1);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            val[r][v] = 0.f;
            if (v < V) {
                const int* m = maps.m[v];
                const float* src = logits + ((int64_t)v * B + b) * S * S;
                if (m[1] == 0) {
                    val[r][v] = __ldg(src + (int64_t)(m[0] * ic + m[2]) * S + (m[4] * jc + m[5]));
                } else {
                    val[r][v] = __ldg(src + (int64_t)(m[1] * jo + m[2]) * S + (m[3] * io + m[5]));
                    any_t = true;
                }
            }
        }
    }
    if (any_t) {                       // CTA-uniform (the maps are kernel parameters)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int v = 0; v < 8; ++v)
                if (v < V && maps.m[v][1] != 0) tile[v][tx][ty + 8 * r] = val[r][v];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int v = 0; v < 8; ++v)
                if (v < V && maps.m[v][1] != 0) val[r][v] = tile[v][ty + 8 * r][tx];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = blockIdx.y * 32 + ty + 8 * r;
        float acc = val[r][0];
#pragma unroll
        for (int v = 1; v < 8; ++v)
            if (v < V) acc += val[r][v];
        if (i < S && j < S) {
            const float mean = acc / (float)V;
            prob[((int64_t)b * S + i) * S + j] = apply_sigmoid ? sigmoidf_acc(mean) : mean;
        }
    }
}

// Bilinear resize + overwrite paste.  Coordinate arithmetic follows cv2's resizeLinear for CV_32F
// (double coordinates, float weights, horizontal pass then vertical pass).  A CTA covers a 64 x 32
// block of the output; its 64 column and 32 row coordinates are computed ONCE in double precision
// (fp64 is a trickle pipe on this chip) into shared memory, then every thread produces 8 pixels.
constexpr int kPasteBX = 64, kPasteBY = 32;
__global__ void __launch_bounds__(256)
resize_paste_kernel(const float* __restrict__ src, int src_w, int crop_y, int crop_x, int crop_h,
                    int crop_w, float* __restrict__ dst, int dst_h, int dst_w, int dst_y,
                    int dst_x, int out_h, int out_w, double scale_y, double scale_x) {
    __shared__ int s_i0[kPasteBX + kPasteBY], s_i1[kPasteBX + kPasteBY];
    __shared__ float s_f[kPasteBX + kPasteBY];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int bx0 = blockIdx.x * kPasteBX, by0 = blockIdx.y * kPasteBY;
    if (tid < kPasteBX + kPasteBY) {
        // source coordinate in double, fraction rounded to float once (matches cv2 4.x to 2e-7;
        // rounding the coordinate itself to float would cost 5e-5 at x ~ 1000)
        const bool is_x = tid < kPasteBX;
        const int o = is_x ? bx0 + tid : by0 + (tid - kPasteBX);
        const int lim = is_x ? crop_w : crop_h;
        const double d = (o + 0.5) * (is_x ? scale_x : scale_y) - 0.5;
        int s0 = (int)floor(d);
        float f = (float)(d - (double)s0);
        if (s0 < 0) { s0 = 0; f = 0.f; }
        if (s0 >= lim - 1) { s0 = lim - 1; f = 0.f; }
        s_i0[tid] = s0;
        s_i1[tid] = s0 + 1 < lim ? s0 + 1 : lim - 1;
        s_f[tid] = f;
    }
    __syncthreads();
    const int lx = tid & 63;                 // column inside the block
    const int ox = bx0 + lx;
    const int gx = dst_x + ox;
    if (ox >= out_w || gx < 0 || gx >= dst_w) return;
    const int sx = s_i0[lx], sx1 = s_i1[lx];
    const float a1 = s_f[lx], a0 = 1.f - a1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ly = (tid >> 6) + 4 * k;   // 4 rows per pass, 8 passes
        const int oy = by0 + ly;
        const int gy = dst_y + oy;
        if (oy >= out_h || gy < 0 || gy >= dst_h) continue;
        const int sy = s_i0[kPasteBX + ly], sy1 = s_i1[kPasteBX + ly];
        const float b1 = s_f[kPasteBX + ly], b0 = 1.f - b1;
        const float* r0 = src + (int64_t)(crop_y + sy) * src_w + crop_x;
        const float* r1 = src + (int64_t)(crop_y + sy1) * src_w + crop_x;
        const float h0 = __fadd_rn(__fmul_rn(__ldg(r0 + sx), a0), __fmul_rn(__ldg(r0 + sx1), a1));
        const float h1 = __fadd_rn(__fmul_rn(__ldg(r1 + sx), a0), __fmul_rn(__ldg(r1 + sx1), a1));
        dst[(int64_t)gy * dst_w + gx] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
    }
}

struct PreLut {
    float v[3][256];
};

// out[c][y][x] = lut[c][ (p00+p01+p10+p11+2) >> 2 ], window pixels outside the image read as 0.
__global__ void preprocess_tile_kernel(const uint8_t* __restrict__ img, int img_h, int img_w, int y0, int x0, int S,
                                       PreLut lut, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= S || y >= S) return;
    int sum[3] = {2, 2, 2};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int iy = y0 + 2 * y + dy, ix = x0 + 2 * x + dx;
            if (iy >= 0 && iy < img_h && ix >= 0 && ix < img_w) {
                const uint8_t* p = img + ((int64_t)iy * img_w + ix) * 3;
                sum[0] += p[0];
                sum[1] += p[1];
                sum[2] += p[2];
            }
        }
    const int64_t plane = (int64_t)S * S;
    const int64_t o = (int64_t)y * S + x;
    out[o] = lut.v[0][sum[0] >> 2];
    out[plane + o] = lut.v[1][sum[1] >> 2];
    out[2 * plane + o] = lut.v[2][sum[2] >> 2];
}

}  // namespace eds

using namespace eds;

extern "C" int eds_tta_merge(const float* logits, int V, int B, int S, const int* view_maps_host,
                             int apply_sigmoid, float* prob, void* stream) {
    EDS_REQUIRE(logits && prob && view_maps_host, "tta_merge: null pointer");
    EDS_REQUIRE(V >= 1 && V <= 8, "tta_merge: V=%d not in 1..8", V);
    EDS_REQUIRE(B >= 1 && B <= 65535 && S >= 1, "tta_merge: bad B=%d S=%d", B, S);
    ViewMaps maps;
    for (int v = 0; v < V; ++v) {
        const int* m = view_maps_host + v * 6;
        // a signed permutation with offsets that keeps [0,S)^2 inside [0,S)^2
        const bool straight = m[1] == 0 && m[3] == 0 && (m[0] == 1 || m[0] == -1) && (m[4] == 1 || m[4] == -1);
        const bool swapped = m[0] == 0 && m[4] == 0 && (m[1] == 1 || m[1] == -1) && (m[3] == 1 || m[3] == -1);
        EDS_REQUIRE(straight || swapped, "tta_merge: view %d map is not a flip/rot90", v);
        for (int corner = 0; corner < 4; ++corner) {
            const int i = (corner & 1) ? S - 1 : 0, j = (corner & 2) ? S - 1 : 0;
            const int r = m[0] * i + m[1] * j + m[2], c = m[3] * i + m[4] * j + m[5];
            EDS_REQUIRE(r >= 0 && r < S && c >= 0 && c < S, "tta_merge: view %d map leaves the tile", v);
        }
        for (int q = 0; q < 6; ++q) maps.m[v][q] = m[q];
    }
    dim3 block(32, 8), grid(ceil_div(S, 32), ceil_div(S, 32), B);
    tta_merge_kernel<<<grid, block, 0, as_stream(stream)>>>(logits, V, B, S, maps, apply_sigmoid, prob);
    return check_launch("tta_merge_kernel");
}

extern "C" int eds_resize_paste_f32(const float* src, int src_h, int src_w, int crop_y, int crop_x, int crop_h,
                                    int crop_w, float* dst, int dst_h, int dst_w, int dst_y, int dst_x,
                                    int out_h, int out_w, void* stream) {
    EDS_REQUIRE(src && dst, "resize_paste: null pointer");
    EDS_REQUIRE(crop_h > 0 && crop_w > 0 && out_h > 0 && out_w > 0, "resize_paste: empty crop or output");
    EDS_REQUIRE(crop_y >= 0 && crop_x >= 0 && crop_y + crop_h <= src_h && crop_x + crop_w <= src_w,
                "resize_paste: crop [%d:%d,%d:%d] outside %dx%d source", crop_y, crop_y + crop_h, crop_x,
                crop_x + crop_w, src_h, src_w);
    dim3 block(64, 4), grid(ceil_div(out_w, kPasteBX), ceil_div(out_h, kPasteBY));
    resize_paste_kernel<<<grid, block, 0, as_stream(stream)>>>(src, src_w, crop_y, crop_x, crop_h, crop_w, dst,
                                                             dst_h, dst_w, dst_y, dst_x, out_h, out_w,
                                                             (double)crop_h / out_h, (double)crop_w / out_w);
    return check_launch("resize_paste_kernel");
}

extern "C" int eds_preprocess_tile_u8(const uint8_t* img, int img_h, int img_w, int y0, int x0, int S,
                                      const double* mean3_host, const double* std3_host, float* out,
                                      void* stream) {
    EDS_REQUIRE(img && out && mean3_host && std3_host, "preprocess_tile: null pointer");
    EDS_REQUIRE(S > 0 && img_h > 0 && img_w > 0, "preprocess_tile: bad sizes");
    PreLut lut;
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v)
            // float64 arithmetic then one rounding, as numpy does for archs/__init__.py:88-97 + .float()
            lut.v[c][v] = (float)((((double)v / 255.0) - mean3_host[c]) / std3_host[c]);
    dim3 block(32, 8), grid(ceil_div(S, 32), ceil_div(S, 8));
    preprocess_tile_kernel<<<grid, block, 0, as_stream(stream)>>>(img, img_h, img_w, y0, x0, S, lut, out);
    return check_launch("preprocess_tile_kernel");
}
