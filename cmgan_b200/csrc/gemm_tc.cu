// tf32 tensor-core implementation of the row-parallel GEMM contract (gemm_args.h) for sm_100a.
// Persistent, warp-specialised: each CTA loops over 128-row output tiles; the accumulator lives in TMEM and is double-buffered so
// that the epilogue of tile i overlaps the loads and MMAs of tile i+1.  Two shapes: 2 CTAs/SM x 10 warps (4 epilogue warps) for
// narrow N, 1 CTA/SM x 14 warps (8 epilogue warps) otherwise.
//
//   warps 0-3  A producers, three modes:
//              * TMA (cfg.tma = 1): dense row-major A, one thread issues cp.async.bulk.tensor.2d boxes of 128 rows x 32 floats that
//                land directly in the K-major SWIZZLE_128B layout; rows past M are zero-filled by the unit.
//              * TMA patches (cfg.tma = 2, template PATCH): same-size convolutions; a tile is a 16 x 8 (h x w) patch of one image and
//                a chunk is one (tap, 32 channels) box of a 4-D (C, W, H, B) tensor map with the tap's (dy, dx) added to the
//                coordinates -- padding is the unit's out-of-bounds zero fill.
//                With TMA every stage of the ring is in flight; LDGSTS producers saturate near 16 GB/s per SM.
//              * cp.async (LDGSTS 16 B, zero fill) for strided / transposed convolutions, and LDG -> registers -> transform ->
//                st.shared for operands with a prologue (BatchNorm+Swish / Swish+dropout / dropout / LayerNorm / InstanceNorm+PReLU).
//   warp 4     TMEM allocation (2 x N columns); one lane issues tcgen05.mma (kind::tf32, M = 128, N = 16..256, K = 8) and
//              tcgen05.commit (A-stage release, accumulator ready).
//   warp 5     weight tiles by cp.async.bulk (UBLKCP) of pre-tiled, pre-swizzled (N x 128 B) blocks: all K chunks once per CTA when
//              the whole weight fits in shared memory ("resident", every conformer GEMM), else per K chunk through the stage ring
//              ("streamed", the dilated dense convolutions).
//   warps 6+   epilogue (compile-time kind): two tcgen05.ld 32x32b.x16 in flight -> per-warp shared-memory staging -> coalesced
//              float4 rows: bias, dropout, residual, activation gradients, Swish dual output -> global; auxiliary operands are
//              prefetched one to two 8-row batches ahead; then the accumulator buffer is released.
//
// The weight operand is re-tiled once per call by pack_b_kernel into the scratch the caller passes (any source layout:
// Linear (N,K), Conv2d (N,C,kh,kw), and the transposed forms used for data gradients).
// All waits are bounded (a protocol bug traps instead of hanging the GPU).
#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched from the driver at run time)

#include "common.cuh"
#include "../../include/cmgan_b200.h"
#include "gemm_device.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace cmgan_gemm;
using namespace cmgan_tc;

constexpr int BM = 128;              // rows per tile = UMMA M
constexpr int KC = 32;               // floats per K chunk = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * KC * 4;   // 16 KB
constexpr int NPROD = 128;           // producer threads (warps 0-3)
constexpr int NTHREADS4 = 320, NTHREADS8 = 448;     // 4 or 8 epilogue warps
constexpr int SLAB = 64;             // epilogue column slab
constexpr int STG_LD = SLAB + 4;     // staging row stride (floats): conflict-free 128-bit accesses
constexpr int STG_BYTES4 = 4 * 32 * STG_LD * 4;     // 34816: per-warp staging of 32 rows x (64 + 4) floats
constexpr int STG_BYTES8 = 8 * 32 * STG_LD * 4;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int RESIDENT_MAX = 96 * 1024;

// ---- weight re-tiling ------------------------------------------------------------------------------
// out[chunk][n][swizzled 32 floats], chunk = tap * (Cin/32) + kc;  rows n >= N are zero
__global__ void pack_b_kernel(const float* __restrict__ B, long sb_tap, long sb_k, long sb_n, int Cin, int ntaps, int N, int BN,
                              float* __restrict__ out) {
    long total = (long)ntaps * (Cin / KC) * BN * KC;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int kk = (int)(i % KC); long t = i / KC; int n = (int)(t % BN); long chunk = t / BN;
    int cpt = Cin / KC;
    int tap = (int)(chunk / cpt), kc = (int)(chunk % cpt);
    float v = 0.f;
    if (n < N) v = to_tf32(__ldg(B + (long)tap * sb_tap + (long)(kc * KC + kk) * sb_k + (long)n * sb_n));
    int c = kk >> 2, j = kk & 3;
    out[(chunk * BN + n) * KC + ((c ^ (n & 7)) << 2) + j] = v;
}

// every weight of a network in one launch (after the optimiser step): blockIdx.y = descriptor
__global__ void pack_all_kernel(const CmganPackDesc* __restrict__ descs) {
    const CmganPackDesc d = descs[blockIdx.y];
    const int Cin = (int)d.Cin, N = (int)d.N;
    const long total = d.ntaps * (Cin / KC) * (long)N * KC;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int kk = (int)(i % KC); long t = i / KC; int n = (int)(t % N); long chunk = t / N;
        int cpt = Cin / KC;
        int tap = (int)(chunk / cpt), kc = (int)(chunk % cpt);
        float v = to_tf32(__ldg(d.src + (long)tap * d.sb_tap + (long)(kc * KC + kk) * d.sb_k + (long)n * d.sb_n));
        int c = kk >> 2, j = kk & 3;
        d.dst[(chunk * N + n) * KC + ((c ^ (n & 7)) << 2) + j] = v;
    }
}

// tma: 0 = cp.async / register producers, 1 = dense 2-D tensor map, 2 = same-size convolution: tiles are 16 x 8 (h x w) patches of one image
// (pw = 8 positions along w, ph = 16 lines), fetched through a 4-D tensor map with the tap offset added to the coordinates
struct TcCfg { int BN, stages, tmem_cols, resident, ntiles, tma, nfx, nty, W, H; };
constexpr int PW = 8, PH = 16;

template <bool ASYNC_A, bool EPI8, int EPI, bool PATCH>
__global__ void __launch_bounds__(EPI8 ? NTHREADS8 : NTHREADS4, (ASYNC_A && !EPI8) ? 2 : 1) gemm_rows_tc_kernel(const __grid_constant__ CmganGemmArgs g, const float* __restrict__ Bp,
                                                                    const TcCfg cfg, const __grid_constant__ CUtensorMap tmA) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int BN = cfg.BN, stages = cfg.stages;
    const int b_tile_bytes = BN * KC * 4;
    const int cpt = g.Cin / KC;
    const int nchunks = cpt * g.ntaps;
    const uint32_t sA = base;
    const uint32_t sB = sA + stages * A_STAGE_BYTES;
    const uint32_t b_region = (uint32_t)(cfg.resident ? nchunks : stages) * b_tile_bytes;
    const uint32_t sStg = sB + b_region;
    const uint32_t bars = sStg + (EPI8 ? STG_BYTES8 : STG_BYTES4);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (stages + s); };
    const uint32_t tfull_bar = bars + 8u * (2 * stages);          // [2]
    const uint32_t tempty_bar = tfull_bar + 16u;                  // [2]
    const uint32_t bready_bar = tempty_bar + 16u;
    const uint32_t tmem_ptr_addr = bready_bar + 8u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = cfg.ntiles;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), (cfg.tma ? 1 : NPROD) + (cfg.resident ? 0 : 1)); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar + 8u * b, 1); mbar_init(tempty_bar + 8u * b, EPI8 ? 8 : 4); }
        mbar_init(bready_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr_addr, (uint32_t)cfg.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    const uint32_t acc_stride = (uint32_t)(cfg.tmem_cols / 2);

    if (warp < 4) {
        // ================================ A producers ================================
        const int c = tid & 7;            // 16-byte chunk within the 128-byte row
        const int rr = tid >> 3;          // rows rr, rr+16, ..., rr+112
        uint32_t dst_off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int r = rr + 16 * i; dst_off[i] = r * 128 + ((c ^ (r & 7)) << 4); }
        const long total = (long)my_tiles * nchunks;

        if (ASYNC_A && cfg.tma) {
            // dense row-major A (no gather): one thread drives TMA, a 128-row x 32-float box per K chunk written straight into the
            // SWIZZLE_128B layout (rows past M are zero-filled by the unit); every stage of the ring can be in flight
            if (tid == 0) {
                for (long q = 0; q < total; ++q) {
                    const int lt = (int)(q / nchunks), ch = (int)(q - (long)lt * nchunks);
                    const int s = (int)(q % stages);
                    const uint32_t par = (uint32_t)((q / stages) & 1);
                    mbar_wait(empty_bar(s), par ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), (uint32_t)A_STAGE_BYTES);
                    const int tile = blockIdx.x + lt * gridDim.x;
                    if (!PATCH) {
                        tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, ch * KC, tile * BM, full_bar(s));
                    } else {        // patch (bimg, ty, fx): rows r = 8 * line + position; padding and ragged edges come back as zeros
                        const int fx = tile % cfg.nfx, ty = (tile / cfg.nfx) % cfg.nty, bimg = tile / (cfg.nfx * cfg.nty);
                        const int tap = ch / cpt, kc = ch - tap * cpt;
                        tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, kc * KC, fx * PW + g.dx[tap], ty * PH + g.dy[tap], bimg, full_bar(s));
                    }
                }
            }
            __syncwarp();
        } else if (ASYNC_A) {
            const int LAG = stages >= 3 ? 2 : 1;
            long rowoff[8];
            RowInfo ri[8];
            int cur_tile = -1, cur_tap = -1;
            for (long q = 0; q < total + LAG; ++q) {
                if (q < total) {
                    const int lt = (int)(q / nchunks), ch = (int)(q - (long)lt * nchunks);
                    const int s = (int)(q % stages);
                    const uint32_t par = (uint32_t)((q / stages) & 1);
                    const int tap = ch / cpt, k0 = (ch - tap * cpt) * KC + c * 4;
                    if (lt != cur_tile) {
                        cur_tile = lt; cur_tap = -1;
                        const int m0 = (blockIdx.x + lt * gridDim.x) * BM;
#pragma unroll
                        for (int i = 0; i < 8; ++i) ri[i] = decode_row(g, m0 + rr + 16 * i);
                    }
                    if (tap != cur_tap) {
                        cur_tap = tap;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            long r = in_row_of(g, ri[i], tap);
                            rowoff[i] = r < 0 ? -1 : g.tap_off[tap] + r * g.lda;
                        }
                    }
                    mbar_wait(empty_bar(s), par ^ 1u);
                    const uint32_t sbase = sA + s * A_STAGE_BYTES;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool ok = rowoff[i] >= 0;
                        cp_async16(sbase + dst_off[i], g.A + (ok ? rowoff[i] + k0 : 0), ok ? 16u : 0u);
                    }
                }
                cp_async_commit();
                const long done = q - LAG;
                if (done >= 0) {
                    if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
                    fence_proxy_async();
                    mbar_arrive(full_bar((int)(done % stages)));
                }
            }
        } else {
            float mean[8], rstd[8];
            RowInfo ri[8];
            int cur_tile = -1;
            for (long q = 0; q < total; ++q) {
                const int lt = (int)(q / nchunks), ch = (int)(q - (long)lt * nchunks);
                const int s = (int)(q % stages);
                const uint32_t par = (uint32_t)((q / stages) & 1);
                const int tap = ch / cpt, k0 = (ch - tap * cpt) * KC + c * 4;
                if (lt != cur_tile) {
                    cur_tile = lt;
                    const int m0 = (blockIdx.x + lt * gridDim.x) * BM;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        ri[i] = decode_row(g, m0 + rr + 16 * i);
                        mean[i] = 0.f; rstd[i] = 1.f;
                        if (g.pro == CMGAN_PRO_LN) {        // ntaps == 1: in_row is constant over the K loop
                            long r = in_row_of(g, ri[i], 0);
                            if (r >= 0) { float2 st = __ldg(reinterpret_cast<const float2*>(g.p0) + r); mean[i] = st.x; rstd[i] = st.y; }
                        }
                    }
                }
                ChunkParams cp;
                load_chunk_params(g, k0, cp);
                float4 v[8];
                long rows[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    rows[i] = in_row_of(g, ri[i], tap);
                    v[i] = rows[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(g.A + g.tap_off[tap] + rows[i] * g.lda + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (rows[i] >= 0) v[i] = transform4(g, v[i], rows[i], k0, mean[i], rstd[i], cp);
                mbar_wait(empty_bar(s), par ^ 1u);
                const uint32_t sbase = sA + s * A_STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + dst_off[i]), "f"(to_tf32(v[i].x)), "f"(to_tf32(v[i].y)),
                                 "f"(to_tf32(v[i].z)), "f"(to_tf32(v[i].w)) : "memory");
                fence_proxy_async();
                mbar_arrive(full_bar(s));
            }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN, 0, 0);
            if (cfg.resident) mbar_wait(bready_bar, 0);
            long q = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int buf = lt & 1;
                mbar_wait(tempty_bar + 8u * buf, (uint32_t)(((lt >> 1) & 1) ^ 1));      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * acc_stride;
                for (int ch = 0; ch < nchunks; ++ch, ++q) {
                    const int s = (int)(q % stages);
                    const uint32_t par = (uint32_t)((q / stages) & 1);
                    mbar_wait(full_bar(s), par);
                    tc_fence_after();
                    const uint64_t adesc = make_desc_sw128(sA + s * A_STAGE_BYTES, 16, 1024);
                    const uint64_t bdesc = make_desc_sw128(sB + (cfg.resident ? ch : s) * b_tile_bytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k)       // tf32: K = 8 per instruction = 32 bytes along the swizzled row
                        umma_tf32(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ch | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(s));             // frees the stage once the MMAs above have read it
                }
                umma_commit(tfull_bar + 8u * buf);         // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ================================ weight-tile loader (TMA bulk copies) ================================
        if (lane == 0) {
            if (cfg.resident) {
                mbar_arrive_expect_tx(bready_bar, (uint32_t)(nchunks * b_tile_bytes));
                for (int ch = 0; ch < nchunks; ++ch)
                    bulk_g2s(sB + ch * b_tile_bytes, Bp + (long)ch * BN * KC, (uint32_t)b_tile_bytes, bready_bar);
            } else {
                const long total = (long)my_tiles * nchunks;
                for (long q = 0; q < total; ++q) {
                    const int ch = (int)(q % nchunks);
                    const int s = (int)(q % stages);
                    const uint32_t par = (uint32_t)((q / stages) & 1);
                    mbar_wait(empty_bar(s), par ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), (uint32_t)b_tile_bytes);
                    bulk_g2s(sB + s * b_tile_bytes, Bp + (long)ch * BN * KC, (uint32_t)b_tile_bytes, full_bar(s));
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue (warps 6-9, or 6-13 with EPI8) ================================
        // a warp may only touch its TMEM lane quarter (warp % 4); with EPI8 two warps share a quarter and alternate 64-column slabs.
        // Per slab: 4 x tcgen05.ld in flight -> one wait -> 16 st.shared.v4 (own row) -> the warp re-reads the slab as coalesced
        // 256-byte row segments (16 lanes x float4, 2 rows per pass), applies the compile-time epilogue and stores.
        const int q4 = warp & 3;
        const int ew = warp - 6;
        const int half = EPI8 ? (ew >> 2) : 0;
        constexpr int NH = EPI8 ? 2 : 1;
        constexpr int PD = EPI8 ? 2 : 1;       // auxiliary-operand prefetch distance in batches (register budget: 128 vs 96 per thread)
        constexpr int RING = EPI8 ? 4 : 2;     // divides the 4 batches of a slab, so slots are compile-time constants
        float* stg = reinterpret_cast<float*>(base_ptr + (sStg - base)) + ew * 32 * STG_LD;
        const int col4 = (lane & 15) * 4;
        const int rsub = lane >> 4;
        const int nslabs = (BN + SLAB - 1) / SLAB;
        constexpr bool DROPS = EPI == CMGAN_EPI_DROP_RES || EPI == CMGAN_EPI_DSWISH_DROP || EPI == CMGAN_EPI_SWISH_DUAL;
        const uint32_t seed32 = DROPS ? cmgan_seed32(eff_seed(g)) : 0u;
        const uint32_t thr16 = g.drop_thr >> 16;
        const bool drop_on = DROPS && g.drop_thr != 0u;
        constexpr bool patch = PATCH;          // compile-time: the flat-tile epilogue keeps its simple row arithmetic
        const float inv_keep = g.inv_keep, alpha = g.alpha;
        // auxiliary operand read at the output position: R (DROP_RES, optional), aux (DSWISH_DROP / DBNSWISH), old C (ACC)
        const float* xbase = nullptr;
        long ldx = 0;
        if (EPI == CMGAN_EPI_DROP_RES) { xbase = g.R; ldx = g.ldr; }
        else if (EPI == CMGAN_EPI_DSWISH_DROP || EPI == CMGAN_EPI_DBNSWISH) { xbase = g.aux; ldx = g.ldaux; }
        else if (EPI == CMGAN_EPI_ACC) { xbase = g.C; ldx = g.ldc; }
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int buf = lt & 1;
            // rows of this TMEM lane quarter.  Flat tiles: 32 consecutive rows.  Patch tiles (cfg.tma == 2): 4 image lines x 8 positions.
            // Pass ps handles quarter rows 2 ps + rsub; its row index is mfirst + (ps >> 2) * hi_rows + 2 * (ps & 3).
            const int tile = blockIdx.x + lt * gridDim.x;
            long mfirst;
            int hi_rows, vhi, vlo;          // valid: flat: 8 (ps >> 2) + 2 (ps & 3) + rsub < vlo;  patch: (ps >> 2) < vhi && 2 (ps & 3) + rsub < vlo
            if (patch) {
                const int fx = tile % cfg.nfx, ty = (tile / cfg.nfx) % cfg.nty, bimg = tile / (cfg.nfx * cfg.nty);
                const int y0 = ty * PH + q4 * 4, x0 = fx * PW;
                mfirst = ((long)bimg * cfg.H + y0) * cfg.W + x0 + rsub;
                hi_rows = cfg.W; vhi = cfg.H - y0; vlo = cfg.W - x0;
            } else {
                const int mrow0 = tile * BM + q4 * 32;
                mfirst = (long)mrow0 + rsub;
                hi_rows = 8; vhi = 4; vlo = g.M - mrow0;
            }
            auto row_ok = [&](int ps) { return patch ? ((ps >> 2) < vhi && 2 * (ps & 3) + rsub < vlo) : (2 * ps + rsub < vlo); };
            auto row_delta = [&](int ps) { return (long)(ps >> 2) * hi_rows + 2 * (ps & 3); };
            const bool any_row = row_ok(0);
            // the auxiliary operand does not depend on the accumulator: its first loads are issued before waiting for the MMAs,
            // later batches (4 passes = 8 rows each) one batch ahead of their use
            float4 ex[RING][4];                                 // slot = batch index % RING; PD batches of loads in flight
            auto prefetch = [&](int sl, int b4, float4* dst) {
                const int n = sl * SLAB + col4;
                if (xbase == nullptr || n >= BN) return;
                const float* xp = xbase + mfirst * ldx + n;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (row_ok(b4 * 4 + u)) dst[u] = __ldg(reinterpret_cast<const float4*>(xp + row_delta(b4 * 4 + u) * ldx));
            };
            if (half < nslabs) { prefetch(half, 0, ex[0]); if (PD == 2) prefetch(half, 1, ex[1]); }
            mbar_wait(tfull_bar + 8u * buf, (uint32_t)((lt >> 1) & 1));
            tc_fence_after();
            const uint32_t trow = tmem_base + buf * acc_stride + ((uint32_t)(q4 * 32) << 16);
            bool released = false;
            for (int sl = half; sl < nslabs; sl += NH) {
                const int n0 = sl * SLAB;
                const int ncols = min(SLAB, BN - n0);
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {              // 32 columns at a time: two tcgen05.ld in flight per wait
                    if (hq * 32 >= ncols) break;
                    uint32_t r[32];
                    tmem_ld16_nowait(trow + (uint32_t)(n0 + hq * 32), r);
                    if (hq * 32 + 16 < ncols) tmem_ld16_nowait(trow + (uint32_t)(n0 + hq * 32 + 16), r + 16);
                    tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (hq * 32 + q * 16 < ncols) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4*>(stg + lane * STG_LD + hq * 32 + q * 16 + 4 * j) =
                                    make_uint4(r[16 * q + 4 * j], r[16 * q + 4 * j + 1], r[16 * q + 4 * j + 2], r[16 * q + 4 * j + 3]);
                        }
                }
                if (sl + NH >= nslabs) {                      // this warp's last slab is out of TMEM: the accumulator buffer may be overwritten
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar + 8u * buf);
                    released = true;
                }
                __syncwarp();
                if (col4 < ncols && any_row) {
                    const int n = n0 + col4;
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), e0v = bias4, e1v = bias4;
                    if (g.bias) bias4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                    if (EPI == CMGAN_EPI_DBNSWISH) { e0v = __ldg(reinterpret_cast<const float4*>(g.e0 + n)); e1v = __ldg(reinterpret_cast<const float4*>(g.e1 + n)); }
                    float* cptr = g.C ? g.C + mfirst * g.ldc + n : nullptr;
                    float* c2ptr = EPI == CMGAN_EPI_SWISH_DUAL ? g.C2 + mfirst * g.ldc2 + n : nullptr;
                    const float* sptr = stg + rsub * STG_LD + col4;
                    const long ldc = g.ldc, ldc2 = g.ldc2;
                    const uint32_t pair = (uint32_t)(((unsigned long long)mfirst * (unsigned long long)g.N + (unsigned long long)n) >> 1);
                    const uint32_t phalf = (uint32_t)g.N >> 1;       // one row further = N / 2 pairs further
#pragma unroll
                    for (int b4 = 0; b4 < 4; ++b4) {
                        if (b4 + PD < 4) prefetch(sl, b4 + PD, ex[(b4 + PD) % RING]);
                        else if (sl + NH < nslabs) prefetch(sl + NH, b4 + PD - 4, ex[(b4 + PD) % RING]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int ps = b4 * 4 + u;
                            if (!row_ok(ps)) continue;
                            const long rd = row_delta(ps);
                            const float4 a = *reinterpret_cast<const float4*>(sptr + ps * 2 * STG_LD);
                            float v[4] = {a.x + bias4.x, a.y + bias4.y, a.z + bias4.z, a.w + bias4.w};
                            float ds[4] = {1.f, 1.f, 1.f, 1.f};
                            if (DROPS && drop_on) {
                                const uint32_t pr = pair + (uint32_t)rd * phalf;
                                const uint32_t h0 = cmgan_mix32((pr * 0x9E3779B1u) ^ seed32), h1 = cmgan_mix32(((pr + 1u) * 0x9E3779B1u) ^ seed32);
                                ds[0] = (h0 & 0xFFFFu) >= thr16 ? inv_keep : 0.f; ds[1] = (h0 >> 16) >= thr16 ? inv_keep : 0.f;
                                ds[2] = (h1 & 0xFFFFu) >= thr16 ? inv_keep : 0.f; ds[3] = (h1 >> 16) >= thr16 ? inv_keep : 0.f;
                            }
                            const float4 xe = ex[b4 % RING][u];
                            const float x[4] = {xe.x, xe.y, xe.z, xe.w};
                            if (EPI == CMGAN_EPI_SWISH_DUAL) {
                                if (cptr) *reinterpret_cast<float4*>(cptr + rd * ldc) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] = to_tf32(swishf_(v[j]) * ds[j]);      // operand of the next contraction: round, the tensor core would truncate
                                *reinterpret_cast<float4*>(c2ptr + rd * ldc2) = make_float4(v[0], v[1], v[2], v[3]);
                                continue;
                            }
                            if (EPI == CMGAN_EPI_DROP_RES) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] = alpha * v[j] * ds[j] + (xbase ? x[j] : 0.f);
                            } else if (EPI == CMGAN_EPI_DSWISH_DROP) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] = to_tf32(v[j] * dswishf_(x[j]) * ds[j]);   // feeds the next data-gradient GEMM and a weight-gradient GEMM
                            } else if (EPI == CMGAN_EPI_DBNSWISH) {
                                const float sa[4] = {e0v.x, e0v.y, e0v.z, e0v.w}, sb[4] = {e1v.x, e1v.y, e1v.z, e1v.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] = v[j] * dswishf_(fmaf(x[j], sa[j], sb[j]));
                            } else if (EPI == CMGAN_EPI_ACC) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] = alpha * v[j] + x[j];
                            }
                            *reinterpret_cast<float4*>(cptr + rd * ldc) = make_float4(v[0], v[1], v[2], v[3]);
                        }
                    }
                } else if (sl + NH < nslabs) {
                    prefetch(sl + NH, 0, ex[0]);
                    if (PD == 2) prefetch(sl + NH, 1, ex[1]);
                }
                __syncwarp();
            }
            if (!released) {          // no slab for this warp (narrow N): still part of the release count
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + 8u * buf);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)cfg.tmem_cols);
    }
}

int tc_supported(const CmganGemmArgs* a) {
    if (a->N % 16 || a->N < 16 || a->N > 256) return 0;
    if (a->Cin % KC) return 0;
    if (a->lda % 4 || ((uintptr_t)a->A & 15)) return 0;
    if (a->C && (a->ldc % 4 || ((uintptr_t)a->C & 15))) return 0;
    if (a->epi == CMGAN_EPI_SWISH_DUAL && (a->ldc2 % 4 || ((uintptr_t)a->C2 & 15))) return 0;
    if (a->bias && ((uintptr_t)a->bias & 15)) return 0;
    for (int t = 0; t < a->ntaps; ++t)
        if (a->tap_off[t] % 4) return 0;
    if (!a->ws || a->ws_floats < (long long)a->N * a->Cin * a->ntaps) return 0;
    if ((uintptr_t)a->ws & 127) return 0;
    if (a->pro == CMGAN_PRO_LN && (((uintptr_t)a->p1 & 15) || ((uintptr_t)a->p2 & 15) || a->ntaps != 1)) return 0;
    if (a->pro == CMGAN_PRO_BN_SWISH && (((uintptr_t)a->p0 & 15) || ((uintptr_t)a->p1 & 15))) return 0;
    if (a->pro == CMGAN_PRO_IN_PRELU && (((uintptr_t)a->p0 & 15) || ((uintptr_t)a->p1 & 15) || ((uintptr_t)a->p2 & 15) || a->pstride % 4)) return 0;
    return 1;
}

int g_num_sms = 0;

using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encoder() {
    static PFN_encodeTiled encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    return encode;
}

template <bool ASYNC_A, bool EPI8, int EPI, bool PATCH = false>
int launch_variant(const CmganGemmArgs& a, const TcCfg& cfg, int grid, size_t smem, cudaStream_t st, const CUtensorMap& tm) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_rows_tc_kernel<ASYNC_A, EPI8, EPI, PATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) { cmgan_set_error("gemm_rows_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
        attr_set = true;
    }
    gemm_rows_tc_kernel<ASYNC_A, EPI8, EPI, PATCH><<<grid, EPI8 ? NTHREADS8 : NTHREADS4, smem, st>>>(a, a.ws, cfg, tm);
    return 0;
}

}  // namespace

// tf32 tensor-core path of cmgan_gemm_rows (same contract).  Returns 1 if the shape is not covered (caller falls back).
int cmgan_gemm_rows_tc_launch(const CmganGemmArgs* a, cudaStream_t st) {
    if (!tc_supported(a)) return 1;
    TcCfg cfg;
    cfg.BN = a->N;
    const int b_tile = cfg.BN * KC * 4;
    const int nchunks = (a->Cin / KC) * a->ntaps;
    cfg.resident = (long)nchunks * b_tile <= RESIDENT_MAX ? 1 : 0;
    cfg.tmem_cols = 64;
    while (cfg.tmem_cols < 2 * cfg.BN) cfg.tmem_cols <<= 1;
    // configuration A: two co-resident CTAs per SM with 4 epilogue warps each (twice the loads in flight);
    // configuration B: one CTA per SM with 8 epilogue warps (wide N: TMEM / shared memory allow only one CTA)
    const int resident_bytes = cfg.resident ? nchunks * b_tile : 0;
    const int per_stage = A_STAGE_BYTES + (cfg.resident ? 0 : b_tile);
    int fixed = 1024 /*alignment*/ + STG_BYTES4 + 256 /*barriers*/ + resident_bytes;
    int ctas = (a->pro == CMGAN_PRO_NONE && cfg.tmem_cols <= 256 && fixed + 3 * per_stage <= SMEM_LIMIT / 2) ? 2 : 1;
    const bool epi8 = ctas == 1;
    if (epi8) fixed += STG_BYTES8 - STG_BYTES4;
    cfg.stages = (SMEM_LIMIT / ctas - fixed) / per_stage;
    if (cfg.stages > 6) cfg.stages = 6;
    if (cfg.stages < 2) return 1;
    cfg.ntiles = cdiv(a->M, BM);
    const size_t smem = (size_t)fixed + (size_t)cfg.stages * per_stage;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // dense row-major A without prologue: describe it to TMA (box = 32 floats x 128 rows, SWIZZLE_128B, zero fill out of bounds)
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    cfg.tma = 0;
    cfg.nfx = cfg.nty = 1; cfg.W = cfg.H = 0;
    if (a->pro == CMGAN_PRO_NONE && !a->conv && a->ntaps == 1 && a->lda % 4 == 0) {
        PFN_encodeTiled encode = get_encoder();
        if (encode) {
            const cuuint64_t gdim[2] = {(cuuint64_t)a->Cin, (cuuint64_t)a->M};
            const cuuint64_t gstride[1] = {(cuuint64_t)a->lda * sizeof(float)};
            const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BM};
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->A + a->tap_off[0]), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) cfg.tma = 1;
        }
    }
    // same-size convolution (taps = coordinate offsets of a (C, W, H, B) tensor, padding = out-of-bounds zero fill): 16 x 8 patch tiles
    bool same_off = true;
    for (int t = 1; t < a->ntaps; ++t) same_off = same_off && a->tap_off[t] == a->tap_off[0];
    if (a->pro == CMGAN_PRO_NONE && (a->epi == CMGAN_EPI_NONE || a->epi == CMGAN_EPI_ACC) && a->conv && a->mul_y == 1 && a->mul_x == 1 && a->div_y == 1 && a->div_x == 1 && a->OH == a->IH &&
        a->OW == a->IW && same_off && a->M % ((long long)a->OH * a->OW) == 0) {
        PFN_encodeTiled encode = get_encoder();
        if (encode) {
            const long long Bn = a->M / ((long long)a->OH * a->OW);
            const cuuint64_t gdim[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->IW, (cuuint64_t)a->IH, (cuuint64_t)Bn};
            const cuuint64_t gstride[3] = {(cuuint64_t)a->lda * sizeof(float), (cuuint64_t)a->IW * a->lda * sizeof(float),
                                           (cuuint64_t)a->IH * a->IW * a->lda * sizeof(float)};
            const cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)PW, (cuuint32_t)PH, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a->A + a->tap_off[0]), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const long long nfx = cdiv(a->OW, PW), nty = cdiv(a->OH, PH);
            if (r == CUDA_SUCCESS && Bn * nfx * nty < (1ll << 30)) {
                cfg.tma = 2; cfg.nfx = (int)nfx; cfg.nty = (int)nty; cfg.W = a->OW; cfg.H = a->OH;
                cfg.ntiles = (int)(Bn * nfx * nty);
            }
        }
    }
    if (!a->b_packed) {
        long total = (long)nchunks * cfg.BN * KC;
        pack_b_kernel<<<cdiv(total, 256), 256, 0, st>>>(a->B, a->sb_tap, a->sb_k, a->sb_n, a->Cin, a->ntaps, a->N, cfg.BN, a->ws);
        if (cmgan_check_launch("pack_b_kernel")) return -1;
    }
    const int grid = cfg.ntiles < ctas * g_num_sms ? cfg.ntiles : ctas * g_num_sms;
    const int variant = a->pro != CMGAN_PRO_NONE ? 0 : (epi8 ? 1 : 2);
    int rc = -2;
#define CMGAN_TC_LAUNCH(E)                                                                                            \
    case E:                                                                                                           \
        rc = variant == 0   ? launch_variant<false, true, E>(*a, cfg, grid, smem, st, tm)                             \
             : variant == 1 ? launch_variant<true, true, E>(*a, cfg, grid, smem, st, tm)                              \
                            : launch_variant<true, false, E>(*a, cfg, grid, smem, st, tm);                            \
        break;
    if (cfg.tma == 2) {      // patch tiles (same-size convolutions): forward (plain) and data-gradient (accumulating) epilogues
        if (a->epi == CMGAN_EPI_NONE)
            rc = epi8 ? launch_variant<true, true, CMGAN_EPI_NONE, true>(*a, cfg, grid, smem, st, tm)
                      : launch_variant<true, false, CMGAN_EPI_NONE, true>(*a, cfg, grid, smem, st, tm);
        else
            rc = epi8 ? launch_variant<true, true, CMGAN_EPI_ACC, true>(*a, cfg, grid, smem, st, tm)
                      : launch_variant<true, false, CMGAN_EPI_ACC, true>(*a, cfg, grid, smem, st, tm);
    } else
    switch (a->epi) {
        CMGAN_TC_LAUNCH(CMGAN_EPI_NONE)
        CMGAN_TC_LAUNCH(CMGAN_EPI_DROP_RES)
        CMGAN_TC_LAUNCH(CMGAN_EPI_DSWISH_DROP)
        CMGAN_TC_LAUNCH(CMGAN_EPI_DBNSWISH)
        CMGAN_TC_LAUNCH(CMGAN_EPI_ACC)
        CMGAN_TC_LAUNCH(CMGAN_EPI_SWISH_DUAL)
        default: break;
    }
#undef CMGAN_TC_LAUNCH
    if (rc == -2) { cmgan_set_error("gemm_rows_tc: unknown epilogue %d", a->epi); return -1; }
    if (rc) return -1;
    return cmgan_check_launch("gemm_rows_tc_kernel");
}

// Re-tile n weights (device table of CmganPackDesc) for the tensor-core path in one launch: called once after the optimiser step, so that the
// GEMM launches of the next step find their B operand ready (CmganGemmArgs.b_packed = 1).
// one weight -> its K-major SWIZZLE_128B tile image (the layout cmgan_ffn_fwd / the b_packed GEMM path consume): N * Cin * ntaps floats
CMGAN_API int cmgan_pack_weight(const float* src, float* dst, long long sb_tap, long long sb_k, long long sb_n, int Cin, int ntaps, int N, void* stream) {
    CMGAN_REQUIRE(src && dst && Cin > 0 && Cin % KC == 0 && ntaps >= 1 && N > 0, "cmgan_pack_weight: bad arguments (Cin must be a multiple of %d)", KC);
    const long total = (long)ntaps * Cin * N;
    pack_b_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, sb_tap, sb_k, sb_n, Cin, ntaps, N, N, dst);
    return cmgan_check_launch("pack_b_kernel");
}

CMGAN_API int cmgan_pack_weights(const CmganPackDesc* descs, int n, void* stream) {
    CMGAN_REQUIRE(descs || n == 0, "cmgan_pack_weights: null table");
    if (n == 0) return 0;
    pack_all_kernel<<<dim3(48, n), 256, 0, (cudaStream_t)stream>>>(descs);
    return cmgan_check_launch("pack_all_kernel");
}
