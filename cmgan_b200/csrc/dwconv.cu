// GLU + depthwise Conv1d(k = 31, zero pad 15/15, bias) of the conformer convolution module.
// Reference: conformer.py:30-48,164-168.  Input g (M, 256) = pointwise-conv output, u = g[:, :128] * sigmoid(g[:, 128:]),
// out[tok, c] = bias[c] + sum_k w[c, k] * u[tok + k - 15, c] along the sequence axis (strided rows, SeqGeom).
// HBM-bound: one pass over g, one write of out; the 31-tap window lives in shared memory (staged with 128-bit loads, all of a
// tile's loads in flight before the first use).
#include "common.cuh"
#include "../../include/cmgan_b200.h"

namespace {

constexpr int CH = 128, KS = 31, PADL = 15, TT = 32, ROWS = TT + KS - 1;   // 62 staged rows per tile
constexpr int C4 = CH / 4;

// U[r][c] = a * sigmoid(b) for tokens t0 - 15 + r;  SG (optional, TT rows) = sigmoid(b) of the centre rows
__device__ __forceinline__ void stage_glu(float* U, float* SG, const float* __restrict__ g, long base, long tok_stride, int t0, int L) {
#pragma unroll 4
    for (int idx = threadIdx.x; idx < ROWS * C4; idx += blockDim.x) {
        const int r = idx / C4, c4 = idx % C4;
        const int tok = t0 - PADL + r;
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f), s = u;
        if (tok >= 0 && tok < L) {
            const float4* p = reinterpret_cast<const float4*>(g + (base + (long)tok * tok_stride) * (2 * CH));
            const float4 a = __ldg(p + c4), b = __ldg(p + C4 + c4);
            s = make_float4(sigmoidf_(b.x), sigmoidf_(b.y), sigmoidf_(b.z), sigmoidf_(b.w));
            u = make_float4(a.x * s.x, a.y * s.y, a.z * s.z, a.w * s.w);
        }
        reinterpret_cast<float4*>(U)[idx] = u;
        if (SG && r >= PADL && r < PADL + TT) reinterpret_cast<float4*>(SG)[(r - PADL) * C4 + c4] = s;
    }
}

__global__ void __launch_bounds__(256) glu_dwconv_fwd_kernel(const float* __restrict__ g, SeqGeom sg, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out) {
    __shared__ __align__(16) float U[ROWS * CH];
    const int s = blockIdx.x, t0 = blockIdx.y * TT;
    const long base = seq_base(sg, s);
    stage_glu(U, nullptr, g, base, sg.tok_stride, t0, sg.L);
    const int c = threadIdx.x % CH, half = threadIdx.x / CH;     // 2 halves x 16 tokens
    float wr[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) wr[k] = __ldg(w + c * KS + k);
    const float b = __ldg(bias + c);
    __syncthreads();
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
        const int tl = half * 16 + grp * 4;
        float a0 = b, a1 = b, a2 = b, a3 = b;
#pragma unroll
        for (int m = 0; m < KS + 3; ++m) {
            float u = U[(tl + m) * CH + c];
            if (m < KS) a0 = fmaf(wr[m], u, a0);
            if (m >= 1 && m - 1 < KS) a1 = fmaf(wr[m - 1], u, a1);
            if (m >= 2 && m - 2 < KS) a2 = fmaf(wr[m - 2], u, a2);
            if (m >= 3) a3 = fmaf(wr[m - 3], u, a3);
        }
        float a[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int tok = t0 + tl + q;
            if (tok < sg.L) out[(base + (long)tok * sg.tok_stride) * CH + c] = a[q];
        }
    }
}

// backward.  dz = grad wrt out (M, 128).  dg (M, 256) overwritten; dw (128, 31), dbias (128) accumulated.
__global__ void __launch_bounds__(256) glu_dwconv_bwd_kernel(const float* __restrict__ g, const float* __restrict__ dz, SeqGeom sg,
                                                             const float* __restrict__ w, int seqs_per_block, float* __restrict__ dg,
                                                             float* __restrict__ dw, float* __restrict__ dbias, int rnd) {
    extern __shared__ __align__(16) float smem[];
    float* U = smem;                    // [ROWS][CH]
    float* DZ = smem + ROWS * CH;       // [ROWS][CH]
    float* SG = DZ + ROWS * CH;         // [TT][CH]
    const int c = threadIdx.x % CH, half = threadIdx.x / CH;
    float wr[KS], dwr[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) { wr[k] = __ldg(w + c * KS + k); dwr[k] = 0.f; }
    float db = 0.f;
    const int s_beg = blockIdx.x * seqs_per_block, s_end = min(s_beg + seqs_per_block, sg.n_seq);
    for (int s = s_beg; s < s_end; ++s) {
        const long base = seq_base(sg, s);
        for (int t0 = 0; t0 < sg.L; t0 += TT) {
            __syncthreads();
            stage_glu(U, SG, g, base, sg.tok_stride, t0, sg.L);
#pragma unroll 4
            for (int idx = threadIdx.x; idx < ROWS * C4; idx += blockDim.x) {
                const int r = idx / C4, c4 = idx % C4;
                const int tok = t0 - PADL + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tok >= 0 && tok < sg.L) v = __ldg(reinterpret_cast<const float4*>(dz + (base + (long)tok * sg.tok_stride) * CH) + c4);
                reinterpret_cast<float4*>(DZ)[idx] = v;
            }
            __syncthreads();
            // 4 consecutive tokens per pass share one sweep over the 34 staged rows they touch (68 shared-memory reads for 248 FMAs)
#pragma unroll 1
            for (int grp = 0; grp < 4; ++grp) {
                const int tl = half * 16 + grp * 4;
                if (t0 + tl >= sg.L) break;
                float du[4] = {0.f, 0.f, 0.f, 0.f}, dzc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { dzc[i] = DZ[(tl + i + PADL) * CH + c]; db += dzc[i]; }
                // du[tok] = sum_k w[k] * dz[tok - k + 15]: staged row tl + m feeds token i with k = 30 - m + i
#pragma unroll
                for (int m = 0; m < KS + 3; ++m) {
                    const float z = DZ[(tl + m) * CH + c];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (m - i >= 0 && m - i < KS) du[i] = fmaf(wr[KS - 1 - m + i], z, du[i]);
                }
                // dw[k] += dz[tok] * u[tok + k - 15]: staged row tl + m feeds (token i, k = m - i)
#pragma unroll
                for (int m = 0; m < KS + 3; ++m) {
                    const float u = U[(tl + m) * CH + c];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (m - i >= 0 && m - i < KS) dwr[m - i] = fmaf(dzc[i], u, dwr[m - i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int tok = t0 + tl + i;
                    if (tok >= sg.L) break;
                    const long row = base + (long)tok * sg.tok_stride;
                    const float sg_ = SG[(tl + i) * CH + c];
                    const float u = U[(tl + i + PADL) * CH + c];            // a * sigmoid(b)
                    dg[row * (2 * CH) + c] = cmgan_maybe_rna(du[i] * sg_, rnd);                    // dg feeds two tensor-core contractions
                    dg[row * (2 * CH) + CH + c] = cmgan_maybe_rna(du[i] * u * (1.f - sg_), rnd);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) atomicAdd(dw + c * KS + k, dwr[k]);
    atomicAdd(dbias + c, db);
}

}  // namespace

CMGAN_API int cmgan_glu_dwconv_fwd(const float* g, const float* w, const float* bias, int B, int T, int F, int axis, float* out, void* stream) {
    CMGAN_REQUIRE(g && w && bias && out && (((uintptr_t)g) & 15) == 0, "cmgan_glu_dwconv_fwd: bad pointer");
    CMGAN_REQUIRE(axis == 0 || axis == 1, "cmgan_glu_dwconv_fwd: bad axis");
    SeqGeom sg = make_seq_geom(B, T, F, axis);
    if (sg.n_seq == 0) return 0;
    dim3 grid(sg.n_seq, cdiv(sg.L, TT));
    glu_dwconv_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, sg, w, bias, out);
    return cmgan_check_launch("glu_dwconv_fwd_kernel");
}

CMGAN_API int cmgan_glu_dwconv_bwd(const float* g, const float* dz, const float* w, int B, int T, int F, int axis, float* dg, float* dw,
                                   float* dbias, void* stream) {
    CMGAN_REQUIRE(g && dz && w && dg && dw && dbias && ((((uintptr_t)g) | ((uintptr_t)dz)) & 15) == 0, "cmgan_glu_dwconv_bwd: bad pointer");
    CMGAN_REQUIRE(axis == 0 || axis == 1, "cmgan_glu_dwconv_bwd: bad axis");
    SeqGeom sg = make_seq_geom(B, T, F, axis);
    if (sg.n_seq == 0) return 0;
    static bool attr_set = false;
    const int smem = (2 * ROWS + TT) * CH * (int)sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(glu_dwconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        CMGAN_REQUIRE(e == cudaSuccess, "cmgan_glu_dwconv_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    // enough blocks for ~4 waves of 148 SMs x 2 resident blocks, but at least one sequence per block
    int spb = sg.n_seq / (148 * 2 * 4);
    if (spb < 1) spb = 1;
    glu_dwconv_bwd_kernel<<<cdiv(sg.n_seq, spb), 256, smem, (cudaStream_t)stream>>>(g, dz, sg, w, spb, dg, dw, dbias, g_cmgan_round_tf32);
    return cmgan_check_launch("glu_dwconv_bwd_kernel");
}
