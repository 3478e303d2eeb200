// Tensor-core (tf32 mma.sync) forward of the relative-position attention, flash style.  Same contract as attn_fwd_kernel in
// attention.cu (which stays the exact-fp32 path): qkv (M, 192) -> ctx (M, 64), lse (M, 4).
//
// One warp owns 16 queries of one (sequence, head); a block = 4 warps = 64 queries; keys are visited in tiles of 64.
// Per key tile and warp:   S  = Q K^T                       16 x 64   (16 mma.m16n8k8, K = head dim 16)
//                          R  = Q E_win^T                   16 x 80   relative-position logits for every distance the tile
//                                                                    can see; written to shared memory and read back skewed:
//                          S[i, j] += R[i, i - j]
//                          online softmax in the accumulator layout (row max / sum across the 4 lanes of a quad)
//                          O += P V                         P re-used in place as the A operand: the accumulator layout
//                                                           (cols 2t, 2t+1) equals the A layout (k = t, t+4) under a fixed
//                                                           permutation of the 8 keys of a k-step, applied to V's rows instead.
// Operands are rounded to tf32 when staged; accumulation, softmax and the running statistics are fp32.
#include "common.cuh"
#include "../../include/cmgan_b200.h"

namespace {

constexpr int D = 16, H = 4, CQ = 64, LDQ = 192;
constexpr int QB = 64;            // queries per block (4 warps x 16)
constexpr int KT = 64;            // keys per tile
constexpr int LDS_ = 20;          // smem row stride of the 16-float operand rows (conflict-free fragment loads)
constexpr int EW = QB + KT - 1;   // 127 relative distances visible to a block per key tile
constexpr int RW = 80;            // relative distances visible to one warp (16 + 64 - 1 = 79, padded to 10 n-tiles)
constexpr int LDR = 84;
constexpr int MAXPOS = 512;
constexpr float SCALE_LOG2E = 0.25f * 1.4426950408889634f;

__device__ __forceinline__ float tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float c[4], const float a[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(const float* __restrict__ qkv, SeqGeom g, const float* __restrict__ E,
                                                           float* __restrict__ ctx, float* __restrict__ lse) {
    __shared__ __align__(16) float Ks[KT * LDS_], Vs[KT * LDS_], Es[EW * LDS_ + 4 * LDS_], Rs[4][16 * LDR];
    const int s = blockIdx.x / H, h = blockIdx.x % H;
    const int i0 = blockIdx.y * QB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t = lane & 3;
    const long base = seq_base(g, s);
    const int iw = i0 + warp * 16;                    // first query of this warp
    const bool warp_active = iw < g.L;

    // ---- Q fragments (rows gq, gq+8), scaled and rounded once
    float qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = iw + gq + (r & 1) * 8, col = t + (r >> 1) * 4 + ks * 8;
            float v = 0.f;
            if (row < g.L) v = __ldg(qkv + (base + (long)row * g.tok_stride) * LDQ + h * D + col) * SCALE_LOG2E;
            qa[ks][r] = tf32r(v);
        }
    float o[2][4];
#pragma unroll
    for (int nd = 0; nd < 2; ++nd)
#pragma unroll
        for (int r = 0; r < 4; ++r) o[nd][r] = 0.f;
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};     // rows gq and gq+8 (lrun: this lane's partial sum)

    for (int j0 = 0; j0 < g.L; j0 += KT) {
        const int nk = min(KT, g.L - j0);
        __syncthreads();
        // ---- stage K, V (64 rows) and the E window (127 rows), rounded to tf32
        for (int idx = tid; idx < KT * 4; idx += 128) {
            const int r = idx >> 2, q4 = idx & 3;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (r < nk) {
                const float* p = qkv + (base + (long)(j0 + r) * g.tok_stride) * LDQ + h * D;
                kv = __ldg(reinterpret_cast<const float4*>(p + CQ) + q4);
                vv = __ldg(reinterpret_cast<const float4*>(p + 2 * CQ) + q4);
            }
            *reinterpret_cast<float4*>(Ks + r * LDS_ + q4 * 4) = make_float4(tf32r(kv.x), tf32r(kv.y), tf32r(kv.z), tf32r(kv.w));
            *reinterpret_cast<float4*>(Vs + r * LDS_ + q4 * 4) = make_float4(tf32r(vv.x), tf32r(vv.y), tf32r(vv.z), tf32r(vv.w));
        }
        // window row w <-> relative distance r = (i0 - j0 - (KT - 1)) + w;  rows beyond EW - 1 (padding of the last n-tile) repeat the edge
        const int rfirst = i0 - j0 - (KT - 1);
        for (int idx = tid; idx < (EW + 4) * 4; idx += 128) {
            const int w = idx >> 2, q4 = idx & 3;
            const int e = clampi(rfirst + w, -MAXPOS, MAXPOS) + MAXPOS;
            const float4 ev = __ldg(reinterpret_cast<const float4*>(E + e * D) + q4);
            *reinterpret_cast<float4*>(Es + w * LDS_ + q4 * 4) = make_float4(tf32r(ev.x), tf32r(ev.y), tf32r(ev.z), tf32r(ev.w));
        }
        __syncthreads();
        if (!warp_active) continue;

        // ---- S = Q K^T
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const float* kp = Ks + (nt * 8 + gq) * LDS_ + t + ks * 8;
                mma_tf32(sc[nt], qa[ks], kp[0], kp[4]);
            }
        }
        // ---- R = Q E^T over this warp's 80 distances: window rows [16 w, 16 w + 80)   (r = iw - j0 - 63 + c)
        float* R = Rs[warp];
#pragma unroll
        for (int nt = 0; nt < RW / 8; ++nt) {
            float rc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const float* ep = Es + (warp * 16 + nt * 8 + gq) * LDS_ + t + ks * 8;
                mma_tf32(rc, qa[ks], ep[0], ep[4]);
            }
            *reinterpret_cast<float2*>(R + gq * LDR + nt * 8 + 2 * t) = make_float2(rc[0], rc[1]);
            *reinterpret_cast<float2*>(R + (gq + 8) * LDR + nt * 8 + 2 * t) = make_float2(rc[2], rc[3]);
        }
        __syncwarp();
        // ---- S[i, j] += R[i, i - j]:  column of (query row q, key column c) is q - c + 63
        float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int qrow = gq + (r >> 1) * 8, kcol = nt * 8 + 2 * t + (r & 1);
                float v = sc[nt][r] + R[qrow * LDR + qrow - kcol + (KT - 1)];
                if (kcol >= nk) v = -INFINITY;
                sc[nt][r] = v;
                tmax[r >> 1] = fmaxf(tmax[r >> 1], v);
            }
        }
        __syncwarp();                 // R is rewritten in the next key tile
        // ---- online softmax (rows gq, gq+8): reduce over the quad
        float corr[2];
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            float m = tmax[hrow];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            const float mnew = fmaxf(mrun[hrow], m);
            corr[hrow] = exp2f(mrun[hrow] - mnew);
            mrun[hrow] = mnew;
            lrun[hrow] *= corr[hrow];
        }
#pragma unroll
        for (int nd = 0; nd < 2; ++nd) { o[nd][0] *= corr[0]; o[nd][1] *= corr[0]; o[nd][2] *= corr[1]; o[nd][3] *= corr[1]; }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float p = exp2f(sc[nt][r] - mrun[r >> 1]);
                lrun[r >> 1] += p;
                sc[nt][r] = tf32r(p);
            }
        // ---- O += P V.  k-step kk covers keys 8 kk .. 8 kk + 7; A-operand column t <-> key 2t, column t+4 <-> key 2t+1
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const float pa[4] = {sc[kk][0], sc[kk][2], sc[kk][1], sc[kk][3]};
#pragma unroll
            for (int nd = 0; nd < 2; ++nd) {
                const float* vp = Vs + (kk * 8 + 2 * t) * LDS_ + nd * 8 + gq;
                mma_tf32(o[nd], pa, vp[0], vp[LDS_]);
            }
        }
    }
    if (!warp_active) return;
    // ---- finish: row sums across the quad, normalise, store
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
        float l = lrun[hrow];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        lrun[hrow] = l;
    }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
        const int i = iw + gq + hrow * 8;
        if (i >= g.L) continue;
        const float inv = 1.f / lrun[hrow];
        const long row = base + (long)i * g.tok_stride;
#pragma unroll
        for (int nd = 0; nd < 2; ++nd)
            *reinterpret_cast<float2*>(ctx + row * CQ + h * D + nd * 8 + 2 * t) = make_float2(o[nd][hrow * 2] * inv, o[nd][hrow * 2 + 1] * inv);
        if (lse && t == 0) lse[row * H + h] = mrun[hrow] + log2f(lrun[hrow]);
    }
}

}  // namespace

// tf32 tensor-core forward (same outputs as cmgan_attention_fwd; logits carry tf32 operand rounding)
CMGAN_API int cmgan_attention_fwd_tf32(const float* qkv, const float* E, int B, int T, int F, int axis, float* ctx, float* lse, void* stream) {
    CMGAN_REQUIRE(qkv && E && ctx, "cmgan_attention_fwd_tf32: null pointer");
    CMGAN_REQUIRE(axis == 0 || axis == 1, "cmgan_attention_fwd_tf32: axis must be 0 (time) or 1 (freq)");
    SeqGeom g = make_seq_geom(B, T, F, axis);
    if (g.n_seq == 0 || g.L == 0) return 0;
    dim3 grid(g.n_seq * H, cdiv(g.L, QB));
    attn_fwd_mma_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(qkv, g, E, ctx, lse);
    return cmgan_check_launch("attn_fwd_mma_kernel");
}
