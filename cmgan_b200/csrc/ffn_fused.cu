// Fused macaron feed-forward of the conformer block (reference conformer.py:54-72,136-148,211-212) for sm_100a:
//
//     out = x + alpha * drop2( W2 ( swish(W1 LN(x) + b1) * drop1 ) + b2 )                alpha = 0.5, C = 64, hidden = 256
//
// ONE kernel per feed-forward: a persistent CTA per SM walks 128-row tiles; the (128 x 256) hidden tile never leaves the SM --
// it is born in TMEM (tcgen05.mma, tf32 operands, fp32 accumulation), activated in registers and handed to the second
// contraction through a ring of K-major SWIZZLE_128B shared-memory chunks; both weight matrices (2 x 64 KB of pre-tiled tf32)
// stay resident in shared memory for the CTA's lifetime.  HBM traffic is the compulsory read of x and write of out.
//
//   warps 0-3   row owners (thread = row): load x, LayerNorm in registers (no shuffles), write the normalised tf32 row into the
//               K-major A tile; one tile later the same warps run the output epilogue of their rows (bias, dropout, alpha, residual).
//   warp 4      TMEM allocation; one lane issues every tcgen05.mma / tcgen05.commit:
//                 GEMM1  H[:, 64 q .. 64 q + 63] = xn W1^T in four N = 64 quarters (the activation warps start on quarter 0 while
//                        quarters 1-3 are still in the tensor pipe; quarter q of the NEXT tile is issued as soon as the activation
//                        warps have drained quarter q of this one),
//                 GEMM2  acc2 += a_chunk W2_chunk^T over eight K = 32 chunks as they arrive in the ring (double-buffered accumulator).
//   warp 5      weight images -> shared memory by cp.async.bulk, once.
//   warps 6-13  activation: tcgen05.ld 32 columns of H -> + b1 -> swish -> counter-based dropout -> round to tf32 -> st.shared into
//               the ring slot in the UMMA K-major swizzled layout -> fence.proxy.async -> mbarrier.
// All mbarrier waits are bounded (tc_ptx.cuh): a protocol bug traps instead of hanging the GPU.
#include "common.cuh"
#include "../../include/cmgan_b200.h"
#include "tc_ptx.cuh"

namespace {
using namespace cmgan_tc;

constexpr int BM = 128, C = 64, HID = 256;
constexpr int CHUNK_BYTES = BM * 128;                 // 128 rows x 32 floats (one 128-byte swizzle row per matrix row) = 16 KB
constexpr int W1_BYTES = 2 * HID * 128;               // 2 K-chunks x 256 rows x 128 B = 64 KB
constexpr int W2_BYTES = 8 * C * 128;                 // 8 K-chunks x  64 rows x 128 B = 64 KB
constexpr int XN_BYTES = 2 * CHUNK_BYTES;             // A tile of GEMM1: 128 x 64
constexpr int RING = 3;                               // hidden chunks in flight between the activation warps and GEMM2
constexpr int NTHREADS = 448;
constexpr int TMEM_COLS = 512;                        // H: 4 x 64, acc2: 2 x 64
constexpr int SMEM_FWD = 1024 + W1_BYTES + W2_BYTES + XN_BYTES + RING * CHUNK_BYTES + 2048;

struct FfnFwdArgs {
    const float* x; long long ldx;
    float* out; long long ldo;
    const float* ln_g; const float* ln_b;
    const float* W1p; const float* b1;
    const float* W2p; const float* b2;
    long long M;
    float alpha;
    unsigned long long seed1, seed2; unsigned int thr; float inv_keep;
    const unsigned long long* seed_dev;
};

// byte offset of (row r, 16-byte unit c of the row's 128 bytes) inside a K-major SWIZZLE_128B chunk
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) ffn_fwd_kernel(const __grid_constant__ FfnFwdArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW1 = base, sW2 = sW1 + W1_BYTES, sXn = sW2 + W2_BYTES, sRing = sXn + XN_BYTES;
    const uint32_t sPar = sRing + RING * CHUNK_BYTES;            // gamma[64] beta[64] b2[64] b1[256] floats = 1792 B
    float* par = reinterpret_cast<float*>(base_ptr + (sPar - base));
    const uint32_t bars = sPar + 1792;
    const uint32_t xn_full = bars, xn_empty = bars + 8, wready = bars + 16;
    auto hq_full = [&](int q) { return bars + 24u + 8u * q; };                   // [4]
    auto hid_full = [&](int s) { return bars + 56u + 8u * s; };                  // [RING]
    auto hid_empty = [&](int s) { return bars + 56u + 8u * (RING + s); };        // [RING]
    auto acc_full = [&](int b) { return bars + 56u + 16u * RING + 8u * b; };     // [2]
    auto acc_empty = [&](int b) { return bars + 72u + 16u * RING + 8u * b; };    // [2]
    const uint32_t tmem_ptr_addr = bars + 88u + 16u * RING;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (int)((g.M + BM - 1) / BM);
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        mbar_init(xn_full, 4); mbar_init(xn_empty, 1); mbar_init(wready, 1);
        for (int q = 0; q < 4; ++q) mbar_init(hq_full(q), 1);
        for (int s = 0; s < RING; ++s) { mbar_init(hid_full(s), 4); mbar_init(hid_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }
        fence_barrier_init();
    }
    for (int i = tid; i < 448; i += NTHREADS)
        par[i] = i < 64 ? __ldg(g.ln_g + i) : i < 128 ? __ldg(g.ln_b + i - 64) : i < 192 ? __ldg(g.b2 + i - 128) : __ldg(g.b1 + i - 192);
    if (warp == 4) tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    const float* gam = par; const float* bet = par + 64; const float* b2s = par + 128; const float* b1s = par + 192;

    if (warp < 4) {
        // ================================ row owners: LayerNorm producer + output epilogue ================================
        const uint32_t seed2_32 = cmgan_seed32(cmgan_eff_seed(g.seed2, g.seed_dev));
        const uint32_t thr16 = g.thr >> 16;
        const bool drop_on = g.thr != 0u;
        auto out_epilogue = [&](int lt) {
            const int buf = lt & 1;
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + tid;
            mbar_wait(acc_full(buf), (uint32_t)((lt >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(HID + buf * C) + ((uint32_t)(warp * 32) << 16);
            uint32_t r[64];
            tmem_ld16_nowait(taddr, r); tmem_ld16_nowait(taddr + 16, r + 16); tmem_ld16_nowait(taddr + 32, r + 32); tmem_ld16_nowait(taddr + 48, r + 48);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(buf));
            if (row < g.M) {
                const float* xr = g.x + row * g.ldx;
                float* orow = g.out + row * g.ldo;
                const uint32_t pair0 = (uint32_t)(((unsigned long long)row * C) >> 1);
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) {
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(xr) + c4);
                    float ds[4] = {1.f, 1.f, 1.f, 1.f};
                    if (drop_on) {
                        const uint32_t pr = pair0 + 2u * c4;
                        const uint32_t h0 = cmgan_mix32((pr * 0x9E3779B1u) ^ seed2_32), h1 = cmgan_mix32(((pr + 1u) * 0x9E3779B1u) ^ seed2_32);
                        ds[0] = (h0 & 0xFFFFu) >= thr16 ? g.inv_keep : 0.f; ds[1] = (h0 >> 16) >= thr16 ? g.inv_keep : 0.f;
                        ds[2] = (h1 & 0xFFFFu) >= thr16 ? g.inv_keep : 0.f; ds[3] = (h1 >> 16) >= thr16 ? g.inv_keep : 0.f;
                    }
                    float4 o;
                    o.x = fmaf(g.alpha * ds[0], __uint_as_float(r[4 * c4 + 0]) + b2s[4 * c4 + 0], xv.x);
                    o.y = fmaf(g.alpha * ds[1], __uint_as_float(r[4 * c4 + 1]) + b2s[4 * c4 + 1], xv.y);
                    o.z = fmaf(g.alpha * ds[2], __uint_as_float(r[4 * c4 + 2]) + b2s[4 * c4 + 2], xv.z);
                    o.w = fmaf(g.alpha * ds[3], __uint_as_float(r[4 * c4 + 3]) + b2s[4 * c4 + 3], xv.w);
                    reinterpret_cast<float4*>(orow)[c4] = o;
                }
            }
        };
        for (int lt = 0; lt < my_tiles; ++lt) {
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + tid;
            float v[64];
            if (row < g.M) {
                const float4* xr = reinterpret_cast<const float4*>(g.x + row * g.ldx);
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) { const float4 t = __ldg(xr + c4); v[4 * c4] = t.x; v[4 * c4 + 1] = t.y; v[4 * c4 + 2] = t.z; v[4 * c4 + 3] = t.w; }
            } else {
#pragma unroll
                for (int k = 0; k < 64; ++k) v[k] = 0.f;
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 64; ++k) s += v[k];
            const float mean = s * (1.f / 64.f);
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < 64; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(q * (1.f / 64.f) + 1e-5f);
            mbar_wait(xn_empty, (uint32_t)((lt & 1) ^ 1));          // GEMM1 of the previous tile has read the A tile
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = to_tf32(fmaf((v[4 * c + j] - mean) * rstd, gam[4 * c + j], bet[4 * c + j]));
                st_shared_v4(sXn + (uint32_t)(c >> 3) * CHUNK_BYTES + sw_off(tid, c & 7), o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(xn_full);
            if (lt > 0) out_epilogue(lt - 1);
        }
        if (my_tiles > 0) out_epilogue(my_tiles - 1);
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc64 = make_idesc_tf32(BM, 64, 0, 0);
            mbar_wait(wready, 0);
            auto issue_h_quarter = [&](int lt, int q) {              // H[:, 64 q ..] of tile lt (xn of that tile is in the A tile)
                if (q == 0) { mbar_wait(xn_full, (uint32_t)(lt & 1)); tc_fence_after(); }
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                    const uint64_t adesc = make_desc_sw128(sXn + kc * CHUNK_BYTES, 16, 1024);
                    const uint64_t bdesc = make_desc_sw128(sW1 + kc * (HID * 128) + q * (64 * 128), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(tmem_base + (uint32_t)(q * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc64, (kc | k) != 0 ? 1u : 0u);
                }
                umma_commit(hq_full(q));
                if (q == 3) umma_commit(xn_empty);
            };
            for (int q = 0; q < 4; ++q) issue_h_quarter(0, q);
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int buf = lt & 1;
                mbar_wait(acc_empty(buf), (uint32_t)(((lt >> 1) & 1) ^ 1));      // output epilogue of tile lt - 2 has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(HID + buf * C);
                for (int j = 0; j < 8; ++j) {
                    const long gch = (long)lt * 8 + j;
                    const int s = (int)(gch % RING);
                    mbar_wait(hid_full(s), (uint32_t)((gch / RING) & 1));
                    tc_fence_after();
                    const uint64_t adesc = make_desc_sw128(sRing + s * CHUNK_BYTES, 16, 1024);
                    const uint64_t bdesc = make_desc_sw128(sW2 + j * (C * 128), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc64, (j | k) != 0 ? 1u : 0u);
                    umma_commit(hid_empty(s));
                    // chunk 2 q + 1 in the ring means the activation warps have finished reading quarter q of H: refill it for the next tile
                    if ((j & 1) && lt + 1 < my_tiles) issue_h_quarter(lt + 1, j >> 1);
                }
                umma_commit(acc_full(buf));
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0) {
            mbar_arrive_expect_tx(wready, (uint32_t)(W1_BYTES + W2_BYTES));
            for (int i = 0; i < 4; ++i) bulk_g2s(sW1 + i * (W1_BYTES / 4), g.W1p + (long)i * (W1_BYTES / 16), (uint32_t)(W1_BYTES / 4), wready);
            for (int i = 0; i < 4; ++i) bulk_g2s(sW2 + i * (W2_BYTES / 4), g.W2p + (long)i * (W2_BYTES / 16), (uint32_t)(W2_BYTES / 4), wready);
        }
        __syncwarp();
    } else {
        // ================================ activation warps (6-13) ================================
        const int lq = warp & 3;                    // TMEM lane quarter this warp may touch
        const int half = (warp - 6) >> 2;           // quarters {half, half + 2} of H
        const int rloc = lq * 32 + lane;            // row within the tile
        const uint32_t seed1_32 = cmgan_seed32(cmgan_eff_seed(g.seed1, g.seed_dev));
        const uint32_t thr16 = g.thr >> 16;
        const bool drop_on = g.thr != 0u;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + rloc;
            for (int qq = 0; qq < 2; ++qq) {
                const int q = half + 2 * qq;
                mbar_wait(hq_full(q), (uint32_t)(lt & 1));
                tc_fence_after();
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int j = 2 * q + c2, n0 = q * 64 + c2 * 32;
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + (uint32_t)n0 + ((uint32_t)(lq * 32) << 16);
                    tmem_ld16_nowait(taddr, r); tmem_ld16_nowait(taddr + 16, r + 16);
                    tmem_wait_ld();
                    float a[32];
                    const uint32_t pair0 = (uint32_t)(((unsigned long long)row * HID + (unsigned long long)n0) >> 1);
#pragma unroll
                    for (int p = 0; p < 16; ++p) {
                        float d0 = 1.f, d1 = 1.f;
                        if (drop_on) {
                            const uint32_t h = cmgan_mix32(((pair0 + (uint32_t)p) * 0x9E3779B1u) ^ seed1_32);
                            d0 = (h & 0xFFFFu) >= thr16 ? g.inv_keep : 0.f; d1 = (h >> 16) >= thr16 ? g.inv_keep : 0.f;
                        }
                        const float v0 = __uint_as_float(r[2 * p]) + b1s[n0 + 2 * p], v1 = __uint_as_float(r[2 * p + 1]) + b1s[n0 + 2 * p + 1];
                        a[2 * p] = to_tf32(swishf_(v0) * d0);
                        a[2 * p + 1] = to_tf32(swishf_(v1) * d1);
                    }
                    const long gch = (long)lt * 8 + j;
                    const int s = (int)(gch % RING);
                    mbar_wait(hid_empty(s), (uint32_t)(((gch / RING) & 1) ^ 1));
                    const uint32_t dst = sRing + s * CHUNK_BYTES;
#pragma unroll
                    for (int c = 0; c < 8; ++c) st_shared_v4(dst + sw_off(rloc, c), a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(hid_full(s));
                }
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int g_sms = 0;
int num_sms() {
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}

}  // namespace

// out = x + alpha * drop(seed2)( W2 ( swish(W1 LN(x) + b1) * drop(seed1) ) + b2 ).  W1p / W2p: the weights re-tiled by cmgan_pack_weights
// (W1 (256, 64): sb_k = 1, sb_n = 64, N = 256, Cin = 64;  W2 (64, 256): sb_k = 1, sb_n = 256, N = 64, Cin = 256).  thr = p * 2^32 (0: no dropout).
// Dropout element indices are row * 256 + n (hidden) and row * 64 + c (output), i.e. the masks cmgan_dropout_mask exports for those seeds.
CMGAN_API int cmgan_ffn_fwd(const float* x, long long ldx, long long M, const float* ln_g, const float* ln_b, const float* W1p, const float* b1,
                            const float* W2p, const float* b2, float alpha, unsigned long long seed1, unsigned long long seed2, unsigned int thr,
                            float inv_keep, const unsigned long long* seed_dev, float* out, long long ldo, void* stream) {
    CMGAN_REQUIRE(x && out && ln_g && ln_b && W1p && b1 && W2p && b2, "cmgan_ffn_fwd: null pointer");
    CMGAN_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0, "cmgan_ffn_fwd: rows must be 16-byte aligned");
    CMGAN_REQUIRE((((uintptr_t)W1p | (uintptr_t)W2p) & 127) == 0, "cmgan_ffn_fwd: weight images must be 128-byte aligned");
    if (M == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ffn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD);
        if (e != cudaSuccess) { cmgan_set_error("cmgan_ffn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
        attr_set = true;
    }
    FfnFwdArgs a;
    a.x = x; a.ldx = ldx; a.out = out; a.ldo = ldo; a.ln_g = ln_g; a.ln_b = ln_b; a.W1p = W1p; a.b1 = b1; a.W2p = W2p; a.b2 = b2; a.M = M; a.alpha = alpha;
    a.seed1 = seed1; a.seed2 = seed2; a.thr = thr; a.inv_keep = inv_keep; a.seed_dev = seed_dev;
    const int ntiles = (int)((M + BM - 1) / BM);
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    ffn_fwd_kernel<<<grid, NTHREADS, SMEM_FWD, (cudaStream_t)stream>>>(a);
    return cmgan_check_launch("ffn_fwd_kernel");
}
