// Fused macaron feed-forward of the conformer block (reference conformer.py:54-72,136-148,211-212) for sm_100a:
//
//     out = x + alpha * drop2( W2 ( swish(W1 LN(x) + b1) * drop1 ) + b2 )                alpha = 0.5, C = 64, hidden = 256
//
// ONE kernel per feed-forward: a persistent CTA per SM walks 128-row tiles; the (128 x 256) hidden tile never leaves the SM --
// it is born in TMEM (tcgen05.mma, tf32 operands, fp32 accumulation), activated in registers and handed to the second
// contraction through a ring of K-major SWIZZLE_128B shared-memory chunks; both weight matrices (2 x 64 KB of pre-tiled tf32)
// stay resident in shared memory for the CTA's lifetime.  HBM traffic is the compulsory read of x and write of out.
//
//   warps 0-3   LayerNorm producers (thread = row): load x (one tile ahead of the tensor pipe), LayerNorm in registers (no shuffles),
//               write the normalised tf32 row into the K-major A tile.
//   warps 14-17 output epilogue (thread = row): acc2 from TMEM, bias, dropout, alpha, residual (x re-read from L2), store.
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
constexpr int NTHREADS = 576;                         // 18 warps: 4 LayerNorm + MMA + weights + 8 activation + 4 output
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
        // ================================ LayerNorm producers ================================
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
        }
    } else if (warp >= 14) {
        // ================================ output epilogue (warps 14-17) ================================
        const int lq = warp & 3;                    // TMEM lane quarter
        const int rloc = lq * 32 + lane;
        const uint32_t seed2_32 = cmgan_seed32(cmgan_eff_seed(g.seed2, g.seed_dev));
        const uint32_t thr16 = g.thr >> 16;
        const bool drop_on = g.thr != 0u;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int buf = lt & 1;
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + rloc;
            mbar_wait(acc_full(buf), (uint32_t)((lt >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(HID + buf * C) + ((uint32_t)(lq * 32) << 16);
            uint32_t r[64];
            tmem_ld16_nowait(taddr, r); tmem_ld16_nowait(taddr + 16, r + 16); tmem_ld16_nowait(taddr + 32, r + 32); tmem_ld16_nowait(taddr + 48, r + 48);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(buf));
            if (row < g.M) {
                float* orow = g.out + row * g.ldo;
                const float4* xr = reinterpret_cast<const float4*>(g.x + row * g.ldx);      // residual: read by this CTA a tile ago (L2)
                const uint32_t pair0 = (uint32_t)(((unsigned long long)row * C) >> 1);
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) {
                    const float4 xv = __ldg(xr + c4);
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
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc64 = make_idesc_tf32(BM, 64, 0, 0);
            mbar_wait(wready, 0);
            auto issue_h_quarter = [&](int lt, int q, bool block) -> bool {     // H[:, 64 q ..] of tile lt (xn of that tile is in the A tile)
                if (q == 0) {
                    if (block) mbar_wait(xn_full, (uint32_t)(lt & 1));
                    else if (!mbar_test(xn_full, (uint32_t)(lt & 1))) return false;
                    tc_fence_after();
                }
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
                return true;
            };
            for (int q = 0; q < 4; ++q) issue_h_quarter(0, q, true);
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int buf = lt & 1;
                const bool more = lt + 1 < my_tiles;
                int hq_next = 0;                                   // quarters of tile lt + 1 already in the tensor pipe
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
                    // chunk 2 q + 1 in the ring means the activation warps have finished reading quarter q of H: refill it for the next tile,
                    // but never stall GEMM2 on the LayerNorm producers (probe, do not wait)
                    while (more && hq_next < ((j + 1) >> 1)) {
                        if (!issue_h_quarter(lt + 1, hq_next, false)) break;
                        ++hq_next;
                    }
                }
                umma_commit(acc_full(buf));
                while (more && hq_next < 4) { issue_h_quarter(lt + 1, hq_next, true); ++hq_next; }
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
    } else if (warp >= 6) {
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


// =====================================================================================================================================
// Backward of the feed-forward module in one kernel (data gradients; the two weight-gradient GEMMs read what it writes):
//
//     xn = LN(x);  h = xn W1^T + b1;  s = sigmoid(h);  a = h s m1;                      (recomputed: the forward pass saved only x)
//     da = dz W2;  dh = da m1 s (1 + h (1 - s));  dxn = dh W1;  dx = LNbwd(dxn) + dout (+ res2);  dgamma, dbeta
//
// with dz = alpha * m2 * dout materialised by the producer of dout.  Outputs for the weight gradients: a, dh (M, 256), xn (M, 64), all
// rounded to tf32.  Per 128-row tile: H and DA are born in TMEM quarter by quarter (64 hidden columns), the activation warps turn them
// into a / dh, dh goes through a two-slot K-major ring into the third contraction (accumulator double-buffered in TMEM), and the
// epilogue warps run the LayerNorm backward of their rows straight from TMEM.
// Shared memory (fp32 operands are fat): W1 (64 KB, for H) and W1 in its K = hidden form (64 KB, for dxn) are resident, W2 (for DA)
// streams through three 8 KB slots (L2-resident), and xn / dz take turns in ONE 32 KB A tile.
// Roles are warpgroup-aligned so that setmaxnreg can move registers from the issue / copy warps to the epilogue warps:
//   WG0 warps 0-3   row producers: x -> LayerNorm -> A tile (+ xn to HBM, row statistics to smem), then dz -> A tile
//   WG1 warp 4      MMA issuer (event loop over: H quarter, DA quarter, dh chunk; probes instead of blocking waits)
//       warp 5      weight copies (resident images once, W2 pieces continuously)
//   WG2-3 warps 8-15  activation (two groups alternating over the quarters)
//   WG4 warps 16-19 LayerNorm-backward epilogue
constexpr int NT_BWD = 640;
constexpr int W2_PIECE = 64 * 128;                  // one (quarter, K chunk) piece of the DA weight: 64 rows x 128 B
constexpr int NPIECE = 3;
constexpr int SMEM_BWD = 1024 + W1_BYTES + W2_BYTES /* W1 in K = hidden form */ + XN_BYTES + 2 * CHUNK_BYTES + NPIECE * W2_PIECE + 2 * BM * 8 + 1536 + 256;

struct FfnBwdArgs {
    const float* x; long long ldx;
    const float* dz; long long lddz;
    const float* dout; long long lddo;
    const float* res2; long long ldr2;
    float* dx; long long lddx;
    float* a_out; float* dh_out; float* xn_out;
    const float* ln_g; const float* ln_b; const float* b1;
    const float* W1p; const float* W2tp; const float* W1tp;
    float* dgamma; float* dbeta;
    long long M;
    unsigned long long seed1; unsigned int thr; float inv_keep;
    const unsigned long long* seed_dev;
};

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// column sums over the 32 lanes of a warp of a 32-entry per-lane array: afterwards v[0] holds the total of entry `lane`
__device__ __forceinline__ void warp_transpose_sum32(float v[32], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool up = lane & 16;
        const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 8;
        const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 4;
        const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 2;
        const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}

__global__ void __launch_bounds__(NT_BWD, 1) ffn_bwd_kernel(const __grid_constant__ FfnBwdArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW1 = base, sW1t = sW1 + W1_BYTES, sA = sW1t + W2_BYTES, sRing = sA + XN_BYTES, sW2 = sRing + 2 * CHUNK_BYTES;
    const uint32_t sStat = sW2 + NPIECE * W2_PIECE;                 // float2 [2][128]: (mean, rstd) of the rows of the tile
    const uint32_t sPar = sStat + 2 * BM * 8;                       // gamma[64] beta[64] b1[256]
    float2* stat = reinterpret_cast<float2*>(base_ptr + (sStat - base));
    float* par = reinterpret_cast<float*>(base_ptr + (sPar - base));
    const uint32_t bars = sPar + 1536;
    const uint32_t xn_full = bars, dz_full = bars + 8, a_free = bars + 16, wready = bars + 24;
    auto h_full = [&](int q) { return bars + 32u + 8u * q; };        // [4]
    auto da_full = [&](int s) { return bars + 64u + 8u * s; };       // [2]
    auto hid_full = [&](int s) { return bars + 80u + 8u * s; };      // [2]
    auto hid_empty = [&](int s) { return bars + 96u + 8u * s; };     // [2]
    auto acc_full = [&](int b) { return bars + 112u + 8u * b; };     // [2]
    auto acc_empty = [&](int b) { return bars + 128u + 8u * b; };    // [2]
    auto w2_full = [&](int s) { return bars + 144u + 8u * s; };      // [3]
    auto w2_empty = [&](int s) { return bars + 168u + 8u * s; };     // [3]
    const uint32_t tmem_ptr_addr = bars + 192u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (int)((g.M + BM - 1) / BM);
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int NQ = 4 * my_tiles;

    if (tid == 0) {
        mbar_init(xn_full, 4); mbar_init(dz_full, 4); mbar_init(a_free, 1); mbar_init(wready, 1);
        for (int q = 0; q < 4; ++q) mbar_init(h_full(q), 1);
        for (int s = 0; s < 2; ++s) { mbar_init(da_full(s), 1); mbar_init(hid_full(s), 4); mbar_init(hid_empty(s), 1); mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 4); }
        for (int s = 0; s < NPIECE; ++s) { mbar_init(w2_full(s), 1); mbar_init(w2_empty(s), 1); }
        fence_barrier_init();
    }
    for (int i = tid; i < 384; i += NT_BWD) par[i] = i < 64 ? __ldg(g.ln_g + i) : i < 128 ? __ldg(g.ln_b + i - 64) : __ldg(g.b1 + i - 128);
    if (warp == 4) tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    const float* gam = par; const float* bet = par + 64; const float* b1s = par + 128;
    // TMEM columns: H quarters [0, 256), DA slots [256, 384), dxn accumulators [384, 512)
    constexpr uint32_t T_DA = 256, T_ACC = 384;

    if (warp < 4) {
        // ================================ WG0: row producers ================================
        for (int lt = 0; lt < my_tiles; ++lt) {
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + tid;
            const bool ok = row < g.M;
            float v[64];
            if (ok) {
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
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = to_tf32(fmaf((v[k] - mean) * rstd, gam[k], bet[k]));
            if (ok) {
                float4* xo = reinterpret_cast<float4*>(g.xn_out + row * C);
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) xo[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            }
            if (lt >= 2) mbar_wait(acc_empty(lt & 1), (uint32_t)(((lt - 2) >> 1) & 1));     // the epilogue of tile lt - 2 has read its row statistics
            stat[(lt & 1) * BM + tid] = make_float2(mean, rstd);
            if (lt > 0) mbar_wait(a_free, 1u);                       // dz of the previous tile has been consumed (phase 2 lt - 1)
#pragma unroll
            for (int c = 0; c < 16; ++c) st_shared_v4(sA + (uint32_t)(c >> 3) * CHUNK_BYTES + sw_off(tid, c & 7), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(xn_full);
            if (ok) {
                const float4* zr = reinterpret_cast<const float4*>(g.dz + row * g.lddz);
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) { const float4 t = __ldg(zr + c4); v[4 * c4] = t.x; v[4 * c4 + 1] = t.y; v[4 * c4 + 2] = t.z; v[4 * c4 + 3] = t.w; }
            } else {
#pragma unroll
                for (int k = 0; k < 64; ++k) v[k] = 0.f;
            }
            mbar_wait(a_free, 0u);                                   // H of this tile has read xn (phase 2 lt)
#pragma unroll
            for (int c = 0; c < 16; ++c) st_shared_v4(sA + (uint32_t)(c >> 3) * CHUNK_BYTES + sw_off(tid, c & 7), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(dz_full);
        }
    } else if (warp < 8) {
        // ================================ WG1: MMA issuer (warp 4), weight copies (warp 5) ================================
        if (warp == 4 && lane == 0 && my_tiles > 0) {
            const uint32_t idesc64 = make_idesc_tf32(BM, 64, 0, 0);
            mbar_wait(wready, 0);
            int Gh = 0, Gd = 0;              // next H / DA quarter to issue (global quarter index = 4 tile + quarter)
            long cj = 0;                     // next dh chunk to consume (8 per tile)
            const long NC = 8L * my_tiles;
            int xn_tile = -1, dz_tile = -1;  // tiles whose xn / dz have been seen in the A tile
            uint32_t idle = 0;
            while (cj < NC) {
                bool progress = false;
                const int cq = (int)(cj >> 1);                       // quarters completely consumed by the activation warps
                // ---- H quarter: region q free once quarter Gh - 4 is consumed; xn of its tile must be in the A tile
                if (Gh < NQ && (Gh < 4 || cq > Gh - 4)) {
                    const int t = Gh >> 2, q = Gh & 3;
                    if (xn_tile < t && mbar_test(xn_full, (uint32_t)(t & 1))) { xn_tile = t; tc_fence_after(); }
                    if (xn_tile >= t) {
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
                            const uint64_t adesc = make_desc_sw128(sA + kc * CHUNK_BYTES, 16, 1024);
                            const uint64_t bdesc = make_desc_sw128(sW1 + kc * (HID * 128) + q * (64 * 128), 16, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_tf32(tmem_base + (uint32_t)(q * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc64, (kc | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(h_full(q));
                        if (q == 3) umma_commit(a_free);             // phase 2 t: xn consumed, dz may take the A tile
                        ++Gh; progress = true;
                    }
                }
                // ---- DA quarter: slot Gd % 2 free once quarter Gd - 2 is consumed; dz of its tile in the A tile; both W2 pieces landed
                if (Gd < NQ && (Gd < 2 || cq > Gd - 2)) {
                    const int t = Gd >> 2, q = Gd & 3;
                    if (dz_tile < t && mbar_test(dz_full, (uint32_t)(t & 1))) { dz_tile = t; tc_fence_after(); }
                    const long p0 = 2L * Gd, p1 = p0 + 1;
                    if (dz_tile >= t && mbar_test(w2_full((int)(p0 % NPIECE)), (uint32_t)((p0 / NPIECE) & 1)) &&
                        mbar_test(w2_full((int)(p1 % NPIECE)), (uint32_t)((p1 / NPIECE) & 1))) {
                        tc_fence_after();
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
                            const long p = p0 + kc;
                            const uint64_t adesc = make_desc_sw128(sA + kc * CHUNK_BYTES, 16, 1024);
                            const uint64_t bdesc = make_desc_sw128(sW2 + (uint32_t)(p % NPIECE) * W2_PIECE, 16, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_tf32(tmem_base + T_DA + (uint32_t)((Gd & 1) * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc64, (kc | k) != 0 ? 1u : 0u);
                            umma_commit(w2_empty((int)(p % NPIECE)));
                        }
                        umma_commit(da_full(Gd & 1));
                        if (q == 3) umma_commit(a_free);             // phase 2 t + 1: dz consumed, xn of the next tile may take the A tile
                        ++Gd; progress = true;
                    }
                }
                // ---- dh chunk -> dxn accumulator
                {
                    const int lt = (int)(cj >> 3), j = (int)(cj & 7), buf = lt & 1, sl = (int)(cj & 1);
                    bool ready = mbar_test(hid_full(sl), (uint32_t)((cj >> 1) & 1));
                    if (ready && j == 0) ready = mbar_test(acc_empty(buf), (uint32_t)(((lt >> 1) & 1) ^ 1));
                    if (ready) {
                        tc_fence_after();
                        const uint64_t adesc = make_desc_sw128(sRing + sl * CHUNK_BYTES, 16, 1024);
                        const uint64_t bdesc = make_desc_sw128(sW1t + j * (C * 128), 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32(tmem_base + T_ACC + (uint32_t)(buf * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc64, (j | k) != 0 ? 1u : 0u);
                        umma_commit(hid_empty(sl));
                        if (j == 7) umma_commit(acc_full(buf));
                        ++cj; progress = true;
                    }
                }
                if (progress) idle = 0;
                else if (++idle > (1u << 27)) __trap();           // protocol bug: fail loudly instead of hanging the device
            }
        } else if (warp == 5 && lane == 0 && my_tiles > 0) {
            mbar_arrive_expect_tx(wready, (uint32_t)(W1_BYTES + W2_BYTES));
            for (int i = 0; i < 4; ++i) bulk_g2s(sW1 + i * (W1_BYTES / 4), g.W1p + (long)i * (W1_BYTES / 16), (uint32_t)(W1_BYTES / 4), wready);
            for (int i = 0; i < 4; ++i) bulk_g2s(sW1t + i * (W2_BYTES / 4), g.W1tp + (long)i * (W2_BYTES / 16), (uint32_t)(W2_BYTES / 4), wready);
            const long NP = 2L * NQ;
            for (long p = 0; p < NP; ++p) {                         // piece p = (quarter G = p / 2, K chunk p % 2) of the DA weight image
                const int sl = (int)(p % NPIECE), q = (int)((p >> 1) & 3), kc = (int)(p & 1);
                mbar_wait(w2_empty(sl), (uint32_t)(((p / NPIECE) & 1) ^ 1));
                mbar_arrive_expect_tx(w2_full(sl), (uint32_t)W2_PIECE);
                bulk_g2s(sW2 + sl * W2_PIECE, g.W2tp + ((long)kc * (HID * 128) + (long)q * W2_PIECE) / 4, (uint32_t)W2_PIECE, w2_full(sl));
            }
        }
        __syncwarp();
    } else if (warp < 16) {
        // ================================ WG2-3: activation ================================
        const int lq = warp & 3, grp = (warp - 8) >> 2, rloc = lq * 32 + lane;
        const uint32_t seed1_32 = cmgan_seed32(cmgan_eff_seed(g.seed1, g.seed_dev));
        const uint32_t thr16 = g.thr >> 16;
        const bool drop_on = g.thr != 0u;
        for (int G = grp; G < NQ; G += 2) {
            const int lt = G >> 2, q = G & 3;
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + rloc;
            const bool ok = row < g.M;
            mbar_wait(h_full(q), (uint32_t)(lt & 1));
            mbar_wait(da_full(grp), (uint32_t)((G >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                const long cj = 8L * lt + 2 * q + c2;
                const uint32_t dst = sRing + c2 * CHUNK_BYTES;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                     // 16 hidden columns at a time (register budget)
                    const int n0 = q * 64 + c2 * 32 + hf * 16;
                    uint32_t rh[16], rd[16];
                    tmem_ld16_nowait(tmem_base + (uint32_t)n0 + ((uint32_t)(lq * 32) << 16), rh);
                    tmem_ld16_nowait(tmem_base + T_DA + (uint32_t)(grp * 64 + c2 * 32 + hf * 16) + ((uint32_t)(lq * 32) << 16), rd);
                    tmem_wait_ld();
                    float av[16], dv[16];
                    const uint32_t pair0 = (uint32_t)(((unsigned long long)row * HID + (unsigned long long)n0) >> 1);
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        float d0 = 1.f, d1 = 1.f;
                        if (drop_on) {
                            const uint32_t h = cmgan_mix32(((pair0 + (uint32_t)p) * 0x9E3779B1u) ^ seed1_32);
                            d0 = (h & 0xFFFFu) >= thr16 ? g.inv_keep : 0.f; d1 = (h >> 16) >= thr16 ? g.inv_keep : 0.f;
                        }
                        const float h0 = __uint_as_float(rh[2 * p]) + b1s[n0 + 2 * p], h1 = __uint_as_float(rh[2 * p + 1]) + b1s[n0 + 2 * p + 1];
                        const float s0 = sigmoidf_(h0), s1 = sigmoidf_(h1);
                        av[2 * p] = to_tf32(h0 * s0 * d0);
                        av[2 * p + 1] = to_tf32(h1 * s1 * d1);
                        dv[2 * p] = to_tf32(__uint_as_float(rd[2 * p]) * d0 * (s0 * (1.f + h0 * (1.f - s0))));
                        dv[2 * p + 1] = to_tf32(__uint_as_float(rd[2 * p + 1]) * d1 * (s1 * (1.f + h1 * (1.f - s1))));
                    }
                    // the ring first: mbarrier.arrive releases every earlier write of the thread, global ones included -- stores to HBM issued before it
                    // would have to drain (MEMBAR) before the tensor pipe may see the chunk
                    if (hf == 0) mbar_wait(hid_empty(c2), (uint32_t)(((cj >> 1) & 1) ^ 1));        // the third contraction has read the previous occupant
#pragma unroll
                    for (int c = 0; c < 4; ++c) st_shared_v4(dst + sw_off(rloc, hf * 4 + c), dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
                    if (hf == 1) {
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(hid_full(c2));
                    }
                    if (ok) {
                        float4* ao = reinterpret_cast<float4*>(g.a_out + row * HID + n0);
                        float4* go = reinterpret_cast<float4*>(g.dh_out + row * HID + n0);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            ao[c] = make_float4(av[4 * c], av[4 * c + 1], av[4 * c + 2], av[4 * c + 3]);
                            go[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
                        }
                    }
                }

            }
        }
    } else {
        // ================================ WG4: LayerNorm-backward epilogue ================================
        // thread = row.  Two sweeps over the row in halves of 32 channels; dxn is simply re-read from TMEM (cheap) and x from L2, so only
        // 2 x 32 values are live at a time.
        const int lq = warp & 3, rloc = lq * 32 + lane;
        float accg0 = 0.f, accg1 = 0.f, accb0 = 0.f, accb1 = 0.f;       // lane's share of dgamma / dbeta: channels lane and 32 + lane
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int buf = lt & 1;
            const long row = ((long)blockIdx.x + (long)lt * gridDim.x) * BM + rloc;
            const bool ok = row < g.M;
            const float4* xr = reinterpret_cast<const float4*>(g.x + (ok ? row : 0) * g.ldx);
            mbar_wait(acc_full(buf), (uint32_t)((lt >> 1) & 1));
            tc_fence_after();
            const float2 st = stat[(lt & 1) * BM + rloc];
            const uint32_t taddr = tmem_base + T_ACC + (uint32_t)(buf * 64) + ((uint32_t)(lq * 32) << 16);
            float c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int hv = 0; hv < 2; ++hv) {
                float r[32];
                tmem_ld16f_nowait(taddr + hv * 32, r); tmem_ld16f_nowait(taddr + hv * 32 + 16, r + 16);
                float4 xv[8];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) xv[c4] = __ldg(xr + hv * 8 + c4);
                tmem_wait_ld();
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float xs[4] = {xv[c4].x, xv[c4].y, xv[c4].z, xv[c4].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float gk = r[4 * c4 + j] * gam[32 * hv + 4 * c4 + j];
                        c1 += gk; c2 = fmaf(gk, (xs[j] - st.x) * st.y, c2);
                    }
                }
            }
            if (!ok) { c1 = 0.f; c2 = 0.f; }
            c1 *= (1.f / 64.f); c2 *= (1.f / 64.f);
#pragma unroll
            for (int hv = 0; hv < 2; ++hv) {
                float r[32], xh[32];
                tmem_ld16f_nowait(taddr + hv * 32, r); tmem_ld16f_nowait(taddr + hv * 32 + 16, r + 16);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 t = __ldg(xr + hv * 8 + c4);
                    xh[4 * c4] = (t.x - st.x) * st.y; xh[4 * c4 + 1] = (t.y - st.x) * st.y; xh[4 * c4 + 2] = (t.z - st.x) * st.y; xh[4 * c4 + 3] = (t.w - st.x) * st.y;
                }
                tmem_wait_ld();
                if (hv == 1) {                                       // last TMEM read of this accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(buf));
                }
                if (ok) {
                    const float4* dor = reinterpret_cast<const float4*>(g.dout + row * g.lddo) + hv * 8;
                    const float4* r2 = g.res2 ? reinterpret_cast<const float4*>(g.res2 + row * g.ldr2) + hv * 8 : nullptr;
                    float4* dxr = reinterpret_cast<float4*>(g.dx + row * g.lddx) + hv * 8;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float4 o = __ldg(dor + c4);
                        if (r2) { const float4 e = __ldg(r2 + c4); o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
                        const int k = 4 * c4, kg = 32 * hv + k;
                        o.x += st.y * (r[k + 0] * gam[kg + 0] - c1 - xh[k + 0] * c2);
                        o.y += st.y * (r[k + 1] * gam[kg + 1] - c1 - xh[k + 1] * c2);
                        o.z += st.y * (r[k + 2] * gam[kg + 2] - c1 - xh[k + 2] * c2);
                        o.w += st.y * (r[k + 3] * gam[kg + 3] - c1 - xh[k + 3] * c2);
                        dxr[c4] = o;
                    }
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) { r[k] = ok ? r[k] : 0.f; xh[k] *= r[k]; }
                warp_transpose_sum32(xh, lane);                      // dgamma: column sums of dxn * xhat
                warp_transpose_sum32(r, lane);                       // dbeta:  column sums of dxn
                if (hv == 0) { accg0 += xh[0]; accb0 += r[0]; } else { accg1 += xh[0]; accb1 += r[0]; }
            }
        }
        if (my_tiles > 0) {
            atomicAdd(g.dgamma + lane, accg0); atomicAdd(g.dgamma + 32 + lane, accg1);
            atomicAdd(g.dbeta + lane, accb0); atomicAdd(g.dbeta + 32 + lane, accb1);
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

// Data gradients of the same module (see ffn_bwd_kernel): dx = LNbwd((dz W2 (.) act'(h)) W1) + dout (+ res2), dgamma / dbeta accumulated, and
// the operands of the two weight-gradient GEMMs written out: a = swish(h) * drop (M, 256), dh (M, 256), xn = LN(x) (M, 64).
// Weight images (cmgan_pack_weights): W1p as in cmgan_ffn_fwd; W2tp = W2 (64, 256) with sb_k = 256, sb_n = 1, N = 256, Cin = 64;
// W1tp = W1 (256, 64) with sb_k = 64, sb_n = 1, N = 64, Cin = 256.
CMGAN_API int cmgan_ffn_bwd(const float* x, long long ldx, const float* dz, long long lddz, const float* dout, long long lddo, const float* res2,
                            long long ldr2, long long M, const float* ln_g, const float* ln_b, const float* W1p, const float* b1, const float* W2tp,
                            const float* W1tp, unsigned long long seed1, unsigned int thr, float inv_keep, const unsigned long long* seed_dev,
                            float* dx, long long lddx, float* a_out, float* dh_out, float* xn_out, float* dgamma, float* dbeta, void* stream) {
    CMGAN_REQUIRE(x && dz && dout && dx && a_out && dh_out && xn_out && ln_g && ln_b && W1p && b1 && W2tp && W1tp && dgamma && dbeta, "cmgan_ffn_bwd: null pointer");
    CMGAN_REQUIRE(ldx % 4 == 0 && lddz % 4 == 0 && lddo % 4 == 0 && ldr2 % 4 == 0 && lddx % 4 == 0, "cmgan_ffn_bwd: leading dimensions must be multiples of 4");
    CMGAN_REQUIRE((((uintptr_t)x | (uintptr_t)dz | (uintptr_t)dout | (uintptr_t)res2 | (uintptr_t)dx | (uintptr_t)a_out | (uintptr_t)dh_out | (uintptr_t)xn_out) & 15) == 0,
                  "cmgan_ffn_bwd: rows must be 16-byte aligned");
    CMGAN_REQUIRE((((uintptr_t)W1p | (uintptr_t)W2tp | (uintptr_t)W1tp) & 127) == 0, "cmgan_ffn_bwd: weight images must be 128-byte aligned");
    if (M == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD);
        if (e != cudaSuccess) { cmgan_set_error("cmgan_ffn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
        attr_set = true;
    }
    FfnBwdArgs a;
    a.x = x; a.ldx = ldx; a.dz = dz; a.lddz = lddz; a.dout = dout; a.lddo = lddo; a.res2 = res2; a.ldr2 = ldr2; a.dx = dx; a.lddx = lddx;
    a.a_out = a_out; a.dh_out = dh_out; a.xn_out = xn_out; a.ln_g = ln_g; a.ln_b = ln_b; a.b1 = b1; a.W1p = W1p; a.W2tp = W2tp; a.W1tp = W1tp;
    a.dgamma = dgamma; a.dbeta = dbeta; a.M = M; a.seed1 = seed1; a.thr = thr; a.inv_keep = inv_keep; a.seed_dev = seed_dev;
    const int ntiles = (int)((M + BM - 1) / BM);
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    ffn_bwd_kernel<<<grid, NT_BWD, SMEM_BWD, (cudaStream_t)stream>>>(a);
    return cmgan_check_launch("ffn_bwd_kernel");
}
