// Axial attention (MHSA encoder blocks, MHCA skip gates) on the tensor cores, bf16 activations.
//
// Same contract and arithmetic as attention.cu (reference: src/main/archs/axial_attention_v2.py:178-213
// AxialAttention.forward, :100-135 CrossAxialAttention.forward; q/k/v projections already applied):
//   sim[d][j] = s_qr * sum_i q[d][i] rq[i][d-j+L-1] + s_kr * sum_i k[d][i] rk[i][d-j+L-1] + s_dots * q[d].k[j]
//   attn      = softmax_j(sim)
//   y[d][i]   = a_kv[i] * sum_j attn[d][j] rv[i][d-j+L-1] + c_kv[i] + a_out[i] * sum_j attn[d][j] v[j][i] + c_out[i]
// Every term is a small matrix product per (sequence, head); the relative-position terms index their
// table at r = d - j + L - 1, i.e. they are products against the table followed by a per-row skew:
//   QR[d][r] = q[d] . rq[:, r],  KR[d][r] = k[d] . rk[:, r]        (m16n8k8, K = 8)
//   S1[d][j] = q[d] . k[j]                                          (m16n8k8)
//   sim[d][j] = s_dots S1[d][j] + (s_qr QR + s_kr KR)[d][d-j+L-1]   (skew read through shared memory)
//   out = attn . V                                                  (m16n8k16, K = L)
//   kv  = skew(attn) . RV^T,  skew(attn)[d][r] = attn[d][d+L-1-r]   (m16n8k16, K = 2L-1)
// The tiles are 16 rows x L <= 64 columns per head -- far below a tcgen05 tile (128 x N with the
// accumulator in TMEM), so these run as warp-level mma.sync with fp32 accumulation: 2.0 of the 1872
// GFLOP of a forward, and the kernel is bound by staging its operands, not by the MMAs.
//
// One CTA = one sequence x 4 heads, 8 warps: two warps per head alternate over 16-row blocks.
#include "common.cuh"

namespace eds {

constexpr int kMmaAttnThreads = 256;
constexpr int kHeadsPerCta = 4;

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

template <int DV, int LP> struct AttnSmem {
    static constexpr int RP = 2 * LP;                 // padded table length (>= 2L-1)
    static constexpr int VT_STRIDE = LP + 8;          // bf16; keeps the B-fragment loads conflict free
    static constexpr int RV_STRIDE = RP + 8;
    static constexpr int E_STRIDE = RP + 4;           // fp32
    static constexpr int P_STRIDE = LP + 8;
    static constexpr size_t relq = 0;                                            // [RP][8] bf16
    static constexpr size_t relk = relq + (size_t)RP * 8 * 2;                    // [RP][8]
    static constexpr size_t rv = relk + (size_t)RP * 8 * 2;                      // [DV][RV_STRIDE]
    static constexpr size_t q = rv + (size_t)DV * RV_STRIDE * 2;                 // [4][LP][8]
    static constexpr size_t k = q + (size_t)kHeadsPerCta * LP * 8 * 2;           // [4][LP][8]
    static constexpr size_t vT = k + (size_t)kHeadsPerCta * LP * 8 * 2;          // [4][DV][VT_STRIDE]
    static constexpr size_t P = vT + (size_t)kHeadsPerCta * DV * VT_STRIDE * 2;  // [8 warps][16][P_STRIDE]
    static constexpr size_t E = (P + (size_t)8 * 16 * P_STRIDE * 2 + 15) & ~(size_t)15;   // [8][16][E_STRIDE] fp32
    static constexpr size_t total = E + (size_t)8 * 16 * E_STRIDE * 4;
};

template <int DV, int LP>
__global__ void __launch_bounds__(kMmaAttnThreads)
axial_attention_mma_kernel(const __nv_bfloat16* __restrict__ qk, int qk_cstride, const __nv_bfloat16* __restrict__ vsrc,
                           int v_cstride, int H, int W, int axis, int heads, const float* __restrict__ rel,
                           const float* __restrict__ sim_scale, const float* __restrict__ out_scale,
                           const float* __restrict__ out_shift, int relu, __nv_bfloat16* __restrict__ y) {
    using S = AttnSmem<DV, LP>;
    constexpr int RP = S::RP;
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* s_relq = reinterpret_cast<__nv_bfloat16*>(smem + S::relq);
    __nv_bfloat16* s_relk = reinterpret_cast<__nv_bfloat16*>(smem + S::relk);
    __nv_bfloat16* s_rv = reinterpret_cast<__nv_bfloat16*>(smem + S::rv);
    __nv_bfloat16* s_q = reinterpret_cast<__nv_bfloat16*>(smem + S::q);
    __nv_bfloat16* s_k = reinterpret_cast<__nv_bfloat16*>(smem + S::k);
    __nv_bfloat16* s_vT = reinterpret_cast<__nv_bfloat16*>(smem + S::vT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = axis == 0 ? H : W;
    const int R = 2 * L - 1;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * kHeadsPerCta;
    int64_t pix0, pstride;
    if (axis == 0) {
        const int n = seq / W, wq = seq % W;
        pix0 = (int64_t)n * H * W + wq;
        pstride = W;
    } else {
        pix0 = (int64_t)seq * W;
        pstride = 1;
    }
    const int G = 16 + (vsrc ? 0 : DV);
    const int CO = heads * DV;

    // ---- stage operands (padding stays zero: 0 * garbage must not become NaN) -------------------------
    for (int i = tid; i < (int)(S::P / 16); i += kMmaAttnThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int idx = tid; idx < (16 + DV) * R; idx += kMmaAttnThreads) {
        const int c = idx / R, r = idx - c * R;
        const __nv_bfloat16 v = __float2bfloat16_rn(__ldg(rel + idx));
        if (c < 8) s_relq[r * 8 + c] = v;
        else if (c < 16) s_relk[r * 8 + (c - 8)] = v;
        else s_rv[(c - 16) * S::RV_STRIDE + r] = v;
    }
    {
        const int vecs = G / 8;                               // 16-byte vectors per (pixel, head)
        for (int idx = tid; idx < L * kHeadsPerCta * vecs; idx += kMmaAttnThreads) {
            const int vec = idx % vecs;
            const int t = idx / vecs;
            const int hl = t % kHeadsPerCta, d = t / kHeadsPerCta;
            const uint4 raw = *reinterpret_cast<const uint4*>(qk + (pix0 + (int64_t)d * pstride) * qk_cstride +
                                                              (h0 + hl) * G + vec * 8);
            if (vec == 0) *reinterpret_cast<uint4*>(s_q + (hl * LP + d) * 8) = raw;
            else if (vec == 1) *reinterpret_cast<uint4*>(s_k + (hl * LP + d) * 8) = raw;
            else {
                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
                for (int i = 0; i < 8; ++i) s_vT[(hl * DV + (vec - 2) * 8 + i) * S::VT_STRIDE + d] = e[i];
            }
        }
        if (vsrc) {
            constexpr int vv = DV / 8;
            for (int idx = tid; idx < L * kHeadsPerCta * vv; idx += kMmaAttnThreads) {
                const int vec = idx % vv;
                const int t = idx / vv;
                const int hl = t % kHeadsPerCta, d = t / kHeadsPerCta;
                const uint4 raw = *reinterpret_cast<const uint4*>(vsrc + (pix0 + (int64_t)d * pstride) * v_cstride +
                                                                  (h0 + hl) * DV + vec * 8);
                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
                for (int i = 0; i < 8; ++i) s_vT[(hl * DV + vec * 8 + i) * S::VT_STRIDE + d] = e[i];
            }
        }
    }
    __syncthreads();

    // ---- per warp: head hl, 16-row blocks rb = sub, sub + 2, ... -------------------------------------------
    const int hl = warp >> 1, sub = warp & 1;
    const int head = h0 + hl;
    const int g = lane >> 2, t = lane & 3;
    const float a_qr = sim_scale[head * 3 + 0], a_kr = sim_scale[head * 3 + 1], a_dots = sim_scale[head * 3 + 2];
    const __nv_bfloat16* hq = s_q + hl * LP * 8;
    const __nv_bfloat16* hk = s_k + hl * LP * 8;
    const __nv_bfloat16* hv = s_vT + hl * DV * S::VT_STRIDE;
    __nv_bfloat16* wP = reinterpret_cast<__nv_bfloat16*>(smem + S::P) + warp * 16 * S::P_STRIDE;
    float* wE = reinterpret_cast<float*>(smem + S::E) + warp * 16 * S::E_STRIDE;
    constexpr int NT_L = LP / 8, NT_R = RP / 8, NT_V = DV / 8;

    for (int rb = sub; rb * 16 < L; rb += 2) {
        const int d0 = rb * 16;
        // (1) E = s_qr * Q.RQ + s_kr * K.RK for this row block -> shared memory (fp32)
        const uint32_t qa0 = lds32(hq + (d0 + g) * 8 + 2 * t), qa1 = lds32(hq + (d0 + g + 8) * 8 + 2 * t);
        const uint32_t ka0 = lds32(hk + (d0 + g) * 8 + 2 * t), ka1 = lds32(hk + (d0 + g + 8) * 8 + 2 * t);
#pragma unroll
        for (int nt = 0; nt < NT_R; ++nt) {
            float cq[4] = {0.f, 0.f, 0.f, 0.f}, ck[4] = {0.f, 0.f, 0.f, 0.f};
            mma_1688(cq, qa0, qa1, lds32(s_relq + (nt * 8 + g) * 8 + 2 * t));
            mma_1688(ck, ka0, ka1, lds32(s_relk + (nt * 8 + g) * 8 + 2 * t));
            float* e0 = wE + g * S::E_STRIDE + nt * 8 + 2 * t;
            float* e1 = wE + (g + 8) * S::E_STRIDE + nt * 8 + 2 * t;
            e0[0] = a_qr * cq[0] + a_kr * ck[0];
            e0[1] = a_qr * cq[1] + a_kr * ck[1];
            e1[0] = a_qr * cq[2] + a_kr * ck[2];
            e1[1] = a_qr * cq[3] + a_kr * ck[3];
        }
        __syncwarp();
        // (2) S1 = Q.K^T, sim = s_dots * S1 + E[d][d-j+L-1], softmax over j
        float p[NT_L][4];
        float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
        for (int nt = 0; nt < NT_L; ++nt) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            mma_1688(c, qa0, qa1, lds32(hk + (nt * 8 + g) * 8 + 2 * t));
            const int j = nt * 8 + 2 * t;
            // rows g / g+8 of the block = sequence positions d0+g / d0+g+8; column index r = d - j + L - 1
            const int r0 = d0 + g - j + L - 1, r1 = r0 + 8;
            // (masked columns j >= L still form an address: clamp it into the row)
            p[nt][0] = j < L ? a_dots * c[0] + wE[g * S::E_STRIDE + min(max(r0, 0), RP - 1)] : -3.0e38f;
            p[nt][1] = j + 1 < L ? a_dots * c[1] + wE[g * S::E_STRIDE + min(max(r0 - 1, 0), RP - 1)] : -3.0e38f;
            p[nt][2] = j < L ? a_dots * c[2] + wE[(g + 8) * S::E_STRIDE + min(max(r1, 0), RP - 1)] : -3.0e38f;
            p[nt][3] = j + 1 < L ? a_dots * c[3] + wE[(g + 8) * S::E_STRIDE + min(max(r1 - 1, 0), RP - 1)] : -3.0e38f;
            m0 = fmaxf(m0, fmaxf(p[nt][0], p[nt][1]));
            m1 = fmaxf(m1, fmaxf(p[nt][2], p[nt][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT_L; ++nt) {
            p[nt][0] = __expf(p[nt][0] - m0);
            p[nt][1] = __expf(p[nt][1] - m0);
            p[nt][2] = __expf(p[nt][2] - m1);
            p[nt][3] = __expf(p[nt][3] - m1);
            s0 += p[nt][0] + p[nt][1];
            s1 += p[nt][2] + p[nt][3];
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float i0 = 1.f / s0, i1 = 1.f / s1;
        uint32_t pa[NT_L][2];                 // attn as bf16 pairs: [nt][0] row g, [nt][1] row g+8
#pragma unroll
        for (int nt = 0; nt < NT_L; ++nt) {
            pa[nt][0] = pack_bf16(p[nt][0] * i0, p[nt][1] * i0);
            pa[nt][1] = pack_bf16(p[nt][2] * i1, p[nt][3] * i1);
            *reinterpret_cast<uint32_t*>(wP + g * S::P_STRIDE + nt * 8 + 2 * t) = pa[nt][0];
            *reinterpret_cast<uint32_t*>(wP + (g + 8) * S::P_STRIDE + nt * 8 + 2 * t) = pa[nt][1];
        }
        __syncwarp();
        // (3) out = attn . V (A fragments straight from the softmax registers)
        float acc_o[NT_V][4], acc_k[NT_V][4];
#pragma unroll
        for (int nt = 0; nt < NT_V; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc_o[nt][i] = acc_k[nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < LP / 16; ++ks) {
            const uint32_t a[4] = {pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1]};
#pragma unroll
            for (int nt = 0; nt < NT_V; ++nt) {
                const __nv_bfloat16* bp = hv + (nt * 8 + g) * S::VT_STRIDE + ks * 16 + 2 * t;
                mma_16816(acc_o[nt], a, lds32(bp), lds32(bp + 8));
            }
        }
        // (4) kv = skew(attn) . RV^T; skew(attn)[d][r] = attn[d][d+L-1-r], non-zero for r in [d, d+L-1]
        const int ks_lo = d0 / 16, ks_hi = min(RP / 16 - 1, (d0 + 15 + L - 1) / 16);
        for (int ks = ks_lo; ks <= ks_hi; ++ks) {
            uint32_t a[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int row = g + (f & 1) * 8;
                const int r = ks * 16 + 2 * t + (f >> 1) * 8;
                const int j = d0 + row + L - 1 - r;                   // column of attn for r; r+1 -> j-1
                const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
                const __nv_bfloat16 lo = (j >= 0 && j < L) ? wP[row * S::P_STRIDE + j] : zero;
                const __nv_bfloat16 hi = (j - 1 >= 0 && j - 1 < L) ? wP[row * S::P_STRIDE + j - 1] : zero;
                a[f] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
            }
#pragma unroll
            for (int nt = 0; nt < NT_V; ++nt) {
                const __nv_bfloat16* bp = s_rv + (nt * 8 + g) * S::RV_STRIDE + ks * 16 + 2 * t;
                mma_16816(acc_k[nt], a, lds32(bp), lds32(bp + 8));
            }
        }
        // (5) out_norm affine (+ ReLU), bf16 store
#pragma unroll
        for (int nt = 0; nt < NT_V; ++nt) {
            const int ch = head * DV + nt * 8 + 2 * t;
            const float2 sk = *reinterpret_cast<const float2*>(out_scale + ch);
            const float2 hk2 = *reinterpret_cast<const float2*>(out_shift + ch);
            const float2 so = *reinterpret_cast<const float2*>(out_scale + CO + ch);
            const float2 ho = *reinterpret_cast<const float2*>(out_shift + CO + ch);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int d = d0 + g + half * 8;
                if (d < L) {
                    float v0 = (sk.x * acc_k[nt][2 * half] + hk2.x) + (so.x * acc_o[nt][2 * half] + ho.x);
                    float v1 = (sk.y * acc_k[nt][2 * half + 1] + hk2.y) + (so.y * acc_o[nt][2 * half + 1] + ho.y);
                    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    *reinterpret_cast<uint32_t*>(y + (pix0 + (int64_t)d * pstride) * CO + ch) = pack_bf16(v0, v1);
                }
            }
        }
        __syncwarp();      // wE / wP are rewritten by the next row block
    }
}

template <int DV, int LP>
static int launch_attn_mma(const void* qk, int qk_cstride, const void* v, int v_cstride, int n_seq, int H, int W, int axis,
                           int heads, const float* rel, const float* sim_scale, const float* out_scale,
                           const float* out_shift, int relu, void* y, cudaStream_t stream) {
    constexpr size_t smem = AttnSmem<DV, LP>::total;
    static PerDevice once;                    // one flag per template instance and per device
    if (smem > 48 * 1024)
        if (int rc = smem_opt_in(once, axial_attention_mma_kernel<DV, LP>, (int)smem, "axial_attention_mma")) return rc;
    dim3 grid(n_seq, heads / kHeadsPerCta);
    axial_attention_mma_kernel<DV, LP><<<grid, kMmaAttnThreads, smem, stream>>>(
        (const __nv_bfloat16*)qk, qk_cstride, (const __nv_bfloat16*)v, v_cstride, H, W, axis, heads, rel, sim_scale,
        out_scale, out_shift, relu, (__nv_bfloat16*)y);
    return check_launch("axial_attention_mma_kernel");
}

// Returns EDS_OK after launching, or 1 when the shape is outside this kernel (caller falls back to the
// CUDA-core kernel of attention.cu, which handles any shape and the fp32 parity mode).
int axial_attention_mma_try(const void* qk, int qk_cstride, const void* v, int v_cstride, int N, int H, int W, int axis,
                            int heads, int dqk, int dv, const float* rel, const float* sim_scale,
                            const float* out_scale, const float* out_shift, int relu, void* y, cudaStream_t stream) {
    const int L = axis == 0 ? H : W;
    if (dqk != 8 || heads % kHeadsPerCta != 0 || L > 64 || L < 1) return 1;
    if (((uintptr_t)qk | (uintptr_t)v | (uintptr_t)y) & 15) return 1;
    if (qk_cstride % 8 != 0 || (v && v_cstride % 8 != 0)) return 1;
    const int LP = (L + 15) / 16 * 16;
    const int n_seq = axis == 0 ? N * W : N * H;
#define EDS_ATTN_CASE(DV_, LP_)                                                                                    \
    if (dv == DV_ && LP == LP_)                                                                                    \
        return launch_attn_mma<DV_, LP_>(qk, qk_cstride, v, v_cstride, n_seq, H, W, axis, heads, rel, sim_scale,   \
                                         out_scale, out_shift, relu, y, stream);
    EDS_ATTN_CASE(64, 16) EDS_ATTN_CASE(64, 32) EDS_ATTN_CASE(64, 48) EDS_ATTN_CASE(64, 64)
    EDS_ATTN_CASE(16, 16) EDS_ATTN_CASE(16, 32) EDS_ATTN_CASE(16, 48) EDS_ATTN_CASE(16, 64)
    EDS_ATTN_CASE(8, 16) EDS_ATTN_CASE(8, 32) EDS_ATTN_CASE(8, 48) EDS_ATTN_CASE(8, 64)
#undef EDS_ATTN_CASE
    return 1;
}

}  // namespace eds
