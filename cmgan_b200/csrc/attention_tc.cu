// tcgen05 forward of the relative-position attention (reference conformer.py:100-131): same contract as cmgan_attention_fwd_tf32
// (qkv (M, 192) -> ctx (M, 64), lse (M, 4); q scaled by 0.25 log2 e, logits in the log2 domain), but every contraction runs on the 5th-gen
// tensor cores with accumulators in TMEM, and the softmax works in a thread-per-query-row layout: no shuffles, no block barriers.
//
// Work item = (sequence, head, tile of 128 queries); a CTA (2 per SM: 256 TMEM columns and ~107 KB of shared memory each, so that one CTA's
// softmax hides the other's MMA / load latency) walks over items, and per item over key tiles of 64:
//   issue warp   S[128 x 64]  = Q K^T            (UMMA M 128, N 64,  K 16)      -> TMEM columns [0, 64)
//                R[128 x 192] = Q E_win^T        (UMMA M 128, N 192, K 16)      -> TMEM columns [64, 256): logits against every relative distance
//                                                                                  the tile can see, d = (i0 - j0 - 63) + w
//   4 softmax    tcgen05.ld S and the R window of the warp's 32 rows; S[i, j] += R[i, i - j + 63] through a 4.6 KB per-warp skew buffer
//   warps        (written position-skewed with conflict-free scalar stores, read back with conflict-free 128-bit loads); running max / sum per
//                thread; P = 2^(s - m) rounded to tf32 into a K-major SWIZZLE_128B tile
//   issue warp   O_tile[128 x 16] = P V         (UMMA M 128, N 16,  K 64)      -> TMEM columns [0, 16) (S is dead by then)
//   4 softmax    O = O * 2^(m_old - m_new) + O_tile in registers; at the end of the item O / l -> ctx, m + log2 l -> lse
//   loader warp  K tile, transposed V tile, E window by cp.async / st.shared, one tile ahead wherever the buffer is already free
// Operand rows are 16 floats; they sit in the first half of 128-byte SWIZZLE_128B rows whose second half stays zero (the K = 16 contraction
// issues two K = 8 steps), which keeps every descriptor in the one layout the rest of this library uses.
#include "common.cuh"
#include "../../include/cmgan_b200.h"
#include "tc_ptx.cuh"

namespace {
using namespace cmgan_tc;

constexpr int D = 16, H = 4, CQ = 64, LDQ = 192;
constexpr int QT = 128, KT = 64, WIN = 192;            // queries / keys per tile, relative distances per tile (191 used)
constexpr int MAXPOS = 512;
constexpr float SCALE_LOG2E = 0.25f * 1.4426950408889634f;
constexpr int PITCH = 36;                             // skew buffer row pitch (floats): PITCH - 1 odd (scalar stores), PITCH % 32 == 4 (128-bit loads)
constexpr int Q_BYTES = QT * 128, K_BYTES = KT * 128, V_BYTES = 2 * 16 * 128, E_BYTES = WIN * 128, P_BYTES = 2 * QT * 128;
constexpr int SKEW_BYTES = 4 * 32 * PITCH * 4;
constexpr int SMEM_ATT = 1024 + Q_BYTES + K_BYTES + V_BYTES + E_BYTES + P_BYTES + SKEW_BYTES + 256;
constexpr int NT = 192;                               // 4 softmax warps + issue warp + loader warp
constexpr int TMEM_ATT = 256;

__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ float tf32q(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(NT, 2) attn_fwd_tc_kernel(const float* __restrict__ qkv, SeqGeom g, const float* __restrict__ E,
                                                            float* __restrict__ ctx, float* __restrict__ lse, int n_items, int nqt, int nkt) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = sQ + Q_BYTES, sV = sK + K_BYTES, sE = sV + V_BYTES, sP = sE + E_BYTES, sSk = sP + P_BYTES;
    const uint32_t bars = sSk + SKEW_BYTES;
    const uint32_t q_full = bars, k_full = bars + 8, v_full = bars + 16, e_full = bars + 24, k_free = bars + 32, v_free = bars + 40, e_free = bars + 48;
    const uint32_t s_full = bars + 56, p_full = bars + 64, o_full = bars + 72, o_read = bars + 80, tmem_ptr_addr = bars + 96;
    float* skew = reinterpret_cast<float*>(base_ptr + (sSk - base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        mbar_init(q_full, 4); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(e_full, 1);
        mbar_init(k_free, 1); mbar_init(v_free, 1); mbar_init(e_free, 1);
        mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_read, 4);
        fence_barrier_init();
    }
    // the operand tiles' second halves (K columns 16 .. 31 of every 128-byte row) are zero for the kernel's lifetime
    for (uint32_t off = tid * 16u; off < (uint32_t)(Q_BYTES + K_BYTES + V_BYTES + E_BYTES); off += NT * 16u)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(sQ + off), "f"(0.f) : "memory");
    if (warp == 4) tmem_alloc(tmem_ptr_addr, TMEM_ATT);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp < 4) {
        // ================================ softmax warps: thread = query row ================================
        const int rloc = warp * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* sk = skew + warp * 32 * PITCH;
        long t = 0;                                              // tile counter of this CTA (barrier phases)
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int qt = item % nqt, h = (item / nqt) % H, s = item / (nqt * H);
            const long sbase = seq_base(g, s);
            const int i = qt * QT + rloc;
            const bool rok = i < g.L;
            const long grow = sbase + (long)(rok ? i : 0) * g.tok_stride;
            // ---- Q row: scaled, rounded, into the A tile (all MMAs of the previous item have completed: its last o_full was waited for)
            {
                float4 qv[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) qv[c] = rok ? __ldg(reinterpret_cast<const float4*>(qkv + grow * LDQ + h * D) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sQ + sw_off(rloc, c)), "f"(to_tf32(qv[c].x * SCALE_LOG2E)),
                                 "f"(to_tf32(qv[c].y * SCALE_LOG2E)), "f"(to_tf32(qv[c].z * SCALE_LOG2E)), "f"(to_tf32(qv[c].w * SCALE_LOG2E)) : "memory");
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(q_full);
            }
            float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = 0.f;
            for (int kt = 0; kt < nkt; ++kt, ++t) {
                const int j0 = kt * KT;
                const int nk = min(KT, g.L - j0);
                mbar_wait(s_full, (uint32_t)(t & 1));
                tc_fence_after();
                // ---- pass A: s[j] = S[i, j] + R[i, i - j + 63], running max
                float sv[64];
                float mx = -INFINITY;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    float rr[64];
                    const uint32_t c0 = (uint32_t)(64 + 32 * (warp - h2 + 1));           // first of the 64 window columns this warp's rows can need
                    tmem_ld16f_nowait(trow + c0, rr); tmem_ld16f_nowait(trow + c0 + 16, rr + 16);
                    tmem_ld16f_nowait(trow + c0 + 32, rr + 32); tmem_ld16f_nowait(trow + c0 + 48, rr + 48);
                    tmem_ld16f_nowait(trow + (uint32_t)(32 * h2), sv + 32 * h2); tmem_ld16f_nowait(trow + (uint32_t)(32 * h2 + 16), sv + 32 * h2 + 16);
                    tmem_wait_ld();
                    __syncwarp();                                   // the previous half's reads of the skew buffer are done
#pragma unroll
                    for (int k = 0; k < 64; ++k) {
                        const int p = k - lane;                     // position-skewed store: every lane's 32 values land in columns 0 .. 31
                        if (p >= 0 && p < 32) sk[lane * PITCH + p] = rr[k];
                    }
                    __syncwarp();
#pragma unroll
                    for (int mq = 0; mq < 8; ++mq) {                // column p <-> key 32 h2 + 31 - p
                        const float4 v = *reinterpret_cast<const float4*>(sk + lane * PITCH + 4 * mq);
                        const int jb = 32 * h2 + 31 - 4 * mq;
                        sv[jb] += v.x; sv[jb - 1] += v.y; sv[jb - 2] += v.z; sv[jb - 3] += v.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    if (j >= nk) sv[j] = -INFINITY;
                    mx = fmaxf(mx, sv[j]);
                }
                const float mnew = fmaxf(m, mx);
                const float corr = ex2(m - mnew);
                m = mnew;
                // ---- pass B: P = 2^(s - m), tf32, into the K-major tile
                float psum = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float p4[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float p = ex2(sv[4 * c + e] - m);
                        psum += p;
                        p4[e] = tf32q(p);
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sP + (uint32_t)(c >> 3) * (QT * 128) + sw_off(rloc, c & 7)), "f"(p4[0]),
                                 "f"(p4[1]), "f"(p4[2]), "f"(p4[3]) : "memory");
                }
                l = l * corr + psum;
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                // ---- O = O * corr + P V (the tile's product lands in the S columns, which are dead now)
                mbar_wait(o_full, (uint32_t)(t & 1));
                tc_fence_after();
                float ot[16];
                tmem_ld16f_nowait(trow, ot);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_read);
#pragma unroll
                for (int c = 0; c < 16; ++c) o[c] = fmaf(o[c], corr, ot[c]);
            }
            if (rok) {
                const float inv = 1.f / l;
                float4* cp = reinterpret_cast<float4*>(ctx + grow * CQ + h * D);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    cp[c] = make_float4(to_tf32(o[4 * c] * inv), to_tf32(o[4 * c + 1] * inv), to_tf32(o[4 * c + 2] * inv), to_tf32(o[4 * c + 3] * inv));
                if (lse) lse[grow * H + h] = m + lg2(l);
            }
        }
    } else if (warp == 4) {
        // ================================ issue warp ================================
        if (lane == 0) {
            const uint32_t idS = make_idesc_tf32(QT, KT, 0, 0), idR = make_idesc_tf32(QT, WIN, 0, 0), idO = make_idesc_tf32(QT, 16, 0, 0);
            long t = 0;
            for (int it = 0; it < my_items; ++it) {
                mbar_wait(q_full, (uint32_t)(it & 1));
                for (int kt = 0; kt < nkt; ++kt, ++t) {
                    const uint32_t par = (uint32_t)(t & 1);
                    mbar_wait(k_full, par);
                    mbar_wait(e_full, par);
                    if (t > 0) mbar_wait(o_read, (uint32_t)((t - 1) & 1));          // the previous tile's product has left the S columns
                    tc_fence_after();
                    const uint64_t qd = make_desc_sw128(sQ, 16, 1024);
                    const uint64_t kd = make_desc_sw128(sK, 16, 1024), ed = make_desc_sw128(sE, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_tf32(tmem_base, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idS, k);
                    umma_commit(k_free);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_tf32(tmem_base + 64, qd + (uint64_t)(2 * k), ed + (uint64_t)(2 * k), idR, k);
                    umma_commit(e_free);
                    umma_commit(s_full);
                    mbar_wait(p_full, par);
                    mbar_wait(v_full, par);
                    tc_fence_after();
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
                        const uint64_t pd = make_desc_sw128(sP + kc * (QT * 128), 16, 1024), vd = make_desc_sw128(sV + kc * (16 * 128), 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, pd + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idO, (kc | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(v_free);
                    umma_commit(o_full);
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ loader warp ================================
        long t = 0;
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int qt = item % nqt, h = (item / nqt) % H, s = item / (nqt * H);
            const long sbase = seq_base(g, s);
            const float* ksrc = qkv + h * D + CQ;
            const float* vsrc = qkv + h * D + 2 * CQ;
            for (int kt = 0; kt < nkt; ++kt, ++t) {
                const int j0 = kt * KT;
                const uint32_t parp = (uint32_t)((t - 1) & 1);
                // ---- K tile: 64 rows x 64 bytes into the first half of the swizzled rows
                if (t > 0) mbar_wait(k_free, parp);
                for (int idx = lane; idx < KT * 4; idx += 32) {
                    const int r = idx >> 2, c = idx & 3;
                    const bool ok = j0 + r < g.L;
                    const long row = sbase + (long)(ok ? j0 + r : 0) * g.tok_stride;
                    cp_async16(sK + sw_off(r, c), ksrc + row * LDQ + c * 4, ok ? 16u : 0u);
                }
                cp_async_commit();
                cp_async_wait<0>();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(k_full);
                // ---- E window: row w <-> distance (i0 - j0 - 63) + w
                if (t > 0) mbar_wait(e_free, parp);
                {
                    const int d0 = qt * QT - j0 - (KT - 1);
                    for (int idx = lane; idx < WIN * 4; idx += 32) {
                        const int w = idx >> 2, c = idx & 3;
                        const int e = clampi(d0 + w, -MAXPOS, MAXPOS) + MAXPOS;
                        cp_async16(sE + sw_off(w, c), E + e * D + c * 4, 16u);
                    }
                    cp_async_commit();
                    cp_async_wait<0>();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(e_full);
                }
                // ---- V tile, transposed: B operand of P V is [channel][key]: 16 rows x 64 keys (2 K chunks of 32 keys)
                if (t > 0) mbar_wait(v_free, parp);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r = half * 32 + lane;                // key within the tile
                    const bool ok = j0 + r < g.L;
                    const long row = sbase + (long)(ok ? j0 + r : 0) * g.tok_stride;
                    float4 vv[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) vv[c] = ok ? __ldg(reinterpret_cast<const float4*>(vsrc + row * LDQ) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float e4[4] = {vv[c].x, vv[c].y, vv[c].z, vv[c].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int ch = 4 * c + e;              // row of the transposed tile
                            const uint32_t a = sV + (uint32_t)half * (16 * 128) + (uint32_t)(ch * 128 + (((lane >> 2) ^ (ch & 7)) << 4) + (lane & 3) * 4);
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(e4[e]) : "memory");
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(v_full);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_ATT);
    }
}

int g_sms_att = 0;

}  // namespace

// tcgen05 forward (same outputs as cmgan_attention_fwd_tf32)
CMGAN_API int cmgan_attention_fwd_tc(const float* qkv, const float* E, int B, int T, int F, int axis, float* ctx, float* lse, void* stream) {
    CMGAN_REQUIRE(qkv && E && ctx, "cmgan_attention_fwd_tc: null pointer");
    CMGAN_REQUIRE(axis == 0 || axis == 1, "cmgan_attention_fwd_tc: axis must be 0 (time) or 1 (freq)");
    CMGAN_REQUIRE((((uintptr_t)qkv | (uintptr_t)ctx | (uintptr_t)E) & 15) == 0, "cmgan_attention_fwd_tc: pointers must be 16-byte aligned");
    SeqGeom g = make_seq_geom(B, T, F, axis);
    if (g.n_seq == 0 || g.L == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATT);
        CMGAN_REQUIRE(e == cudaSuccess, "cmgan_attention_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    if (g_sms_att == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms_att, cudaDevAttrMultiProcessorCount, dev);
    }
    const int nqt = cdiv(g.L, QT), nkt = cdiv(g.L, KT);
    const long n_items = (long)g.n_seq * H * nqt;
    CMGAN_REQUIRE(n_items < (1l << 30), "cmgan_attention_fwd_tc: too many work items");
    const int grid = (int)(n_items < 2L * g_sms_att ? n_items : 2L * g_sms_att);
    attn_fwd_tc_kernel<<<grid, NT, SMEM_ATT, (cudaStream_t)stream>>>(qkv, g, E, ctx, lse, (int)n_items, nqt, nkt);
    return cmgan_check_launch("attn_fwd_tc_kernel");
}
