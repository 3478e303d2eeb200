// tf32 tensor-core implementation of the row-parallel GEMM contract (gemm_args.h) for sm_100a:
// tcgen05.mma (kind::tf32, M = 128, N = 16..256) with the accumulator in TMEM, warp-specialised:
//
//   warps 0-3  A producers, then epilogue.  global (LDG.128, gathered rows of the implicit convolution) -> registers
//              -> prologue (LayerNorm / InstanceNorm+PReLU / BatchNorm+Swish / Swish+dropout) -> round to tf32 ->
//              st.shared in the canonical K-major SWIZZLE_128B layout -> fence.proxy.async -> mbarrier arrive.
//              After the main loop the same warps read the accumulator (tcgen05.ld 32x32b) and run the fused epilogue.
//   warp 4     TMEM allocation; one lane issues tcgen05.mma and tcgen05.commit (stage release / accumulator ready).
//   warp 5     one lane issues the weight-tile loads: cp.async.bulk (TMA bulk copy, UBLKCP) of a pre-tiled,
//              pre-swizzled (N x 128 B) block per K chunk, completing on the stage's mbarrier.
//
// The weight operand is re-tiled once per call by pack_b_kernel into the scratch the caller passes (any source layout:
// Linear (N,K), Conv2d (N,C,kh,kw), and the transposed forms used for data gradients).
// One CTA computes a 128 x N output tile; two CTAs are co-resident per SM so one tile's epilogue overlaps the
// other's main loop.  All waits are bounded (a protocol bug traps instead of hanging the GPU).
#include "common.cuh"
#include "../../include/cmgan_b200.h"
#include "gemm_device.cuh"

namespace {
using namespace cmgan_gemm;

constexpr int BM = 128;              // rows per CTA tile = UMMA M
constexpr int KC = 32;               // floats per K chunk = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * KC * 4;   // 16 KB
constexpr int NPROD = 128;           // producer / epilogue threads (warps 0-3)
constexpr int NTHREADS = 192;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 28)) __trap();        // protocol bug: fail loudly instead of hanging the device
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// K-major SWIZZLE_128B shared-memory descriptor (cute::UMMA::SmemDescriptor, version 1): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset >> 4 bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset >> 4  bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

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

// ---- main kernel --------------------------------------------------------------------------------------
struct SmemLayout { int stages; int b_stage_bytes; };

__global__ void __launch_bounds__(NTHREADS, 2) gemm_rows_tc_kernel(const __grid_constant__ CmganGemmArgs g, const float* __restrict__ Bp,
                                                                    int BN, int stages, int tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // SWIZZLE_128B tiles need 1024-byte alignment
    const int b_stage_bytes = BN * KC * 4;
    const uint32_t sA = base;
    const uint32_t sB = base + stages * A_STAGE_BYTES;
    const uint32_t bars = sB + stages * b_stage_bytes;                  // full[stages], empty[stages], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (stages + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * stages);
    const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM;
    const int cpt = g.Cin / KC;
    const int nchunks = cpt * g.ntaps;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), NPROD + 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr_addr, (uint32_t)tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp < 4) {
        // ================================ A producers ================================
        const int c = tid & 7;            // 16-byte chunk within the 128-byte row
        const int rr = tid >> 3;          // rows rr, rr+16, ..., rr+112
        RowInfo ri[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ri[i] = decode_row(g, m0 + rr + 16 * i);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % stages;
            const uint32_t par = (uint32_t)((ch / stages) & 1);
            const int tap = ch / cpt, k0 = (ch - tap * cpt) * KC + c * 4;
            float v[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) load_a4<4>(g, in_row_of(g, ri[i], tap), tap, k0, v[i]);
            mbar_wait(empty_bar(s), par ^ 1u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rr + 16 * i;
                const uint32_t dst = sA + s * A_STAGE_BYTES + r * 128 + ((c ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(to_tf32(v[i][0])), "f"(to_tf32(v[i][1])),
                             "f"(to_tf32(v[i][2])), "f"(to_tf32(v[i][3])) : "memory");
            }
            fence_proxy_async();
            mbar_arrive(full_bar(s));
        }
        // ================================ epilogue ================================
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const long m = (long)m0 + warp * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < BN; n0 += 16) {
            float acc[16];
            tmem_ld16(trow + (uint32_t)n0, acc);
            if (m < g.M) {
                float* cp = g.C + m * g.ldc + n0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n0 + j;
                    if (n < g.N) {
                        float vv = acc[j] + (g.bias ? __ldg(g.bias + n) : 0.f);
                        acc[j] = epilogue(g, vv, m, n, cp + j);
                    }
                }
                if (n0 + 16 <= g.N && (g.ldc & 3) == 0 && (((uintptr_t)g.C) & 15) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < g.N) cp[j] = acc[j];
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BN);
            for (int ch = 0; ch < nchunks; ++ch) {
                const int s = ch % stages;
                const uint32_t par = (uint32_t)((ch / stages) & 1);
                mbar_wait(full_bar(s), par);
                tc_fence_after();
                const uint64_t adesc = make_desc(sA + s * A_STAGE_BYTES);
                const uint64_t bdesc = make_desc(sB + s * b_stage_bytes);
#pragma unroll
                for (int k = 0; k < KC / 8; ++k)       // tf32: K = 8 per instruction = 32 bytes along the swizzled row
                    umma_tf32(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ch | k) != 0 ? 1u : 0u);
                umma_commit(empty_bar(s));             // frees the stage once the MMAs above have read it
            }
            umma_commit(tmem_full_bar);                // accumulator complete
        }
        __syncwarp();
    } else {
        // ================================ weight-tile loader (TMA bulk copies) ================================
        if (lane == 0) {
            for (int ch = 0; ch < nchunks; ++ch) {
                const int s = ch % stages;
                const uint32_t par = (uint32_t)((ch / stages) & 1);
                mbar_wait(empty_bar(s), par ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), (uint32_t)b_stage_bytes);
                bulk_g2s(sB + s * b_stage_bytes, Bp + (long)ch * BN * KC, (uint32_t)b_stage_bytes, full_bar(s));
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}

int tc_supported(const CmganGemmArgs* a) {
    if (a->N % 16 || a->N < 16 || a->N > 256) return 0;
    if (a->Cin % KC) return 0;
    if (a->lda % 4 || ((uintptr_t)a->A & 15)) return 0;
    for (int t = 0; t < a->ntaps; ++t)
        if (a->tap_off[t] % 4) return 0;
    if (!a->ws || a->ws_floats < (long long)a->N * a->Cin * a->ntaps) return 0;
    if ((uintptr_t)a->ws & 127) return 0;
    return 1;
}

}  // namespace

// tf32 tensor-core path of cmgan_gemm_rows (same contract).  Returns 1 if the shape is not covered (caller falls back).
int cmgan_gemm_rows_tc_launch(const CmganGemmArgs* a, cudaStream_t st) {
    if (!tc_supported(a)) return 1;
    const int BN = a->N;
    const int b_stage = BN * KC * 4;
    int stages = (100 * 1024 - 2048) / (A_STAGE_BYTES + b_stage);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    const int nchunks = (a->Cin / KC) * a->ntaps;
    if (stages > nchunks) stages = nchunks < 2 ? 2 : nchunks;
    int tmem_cols = 32;
    while (tmem_cols < BN) tmem_cols <<= 1;
    const size_t smem = (size_t)stages * (A_STAGE_BYTES + b_stage) + 1024 /*alignment*/ + 8 * (2 * stages + 2) + 16;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(110 * 1024));
        if (e != cudaSuccess) { cmgan_set_error("gemm_rows_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
        smem_set = 110 * 1024;
    }
    long total = (long)nchunks * BN * KC;
    pack_b_kernel<<<cdiv(total, 256), 256, 0, st>>>(a->B, a->sb_tap, a->sb_k, a->sb_n, a->Cin, a->ntaps, a->N, BN, a->ws);
    if (cmgan_check_launch("pack_b_kernel")) return -1;
    gemm_rows_tc_kernel<<<cdiv(a->M, BM), NTHREADS, smem, st>>>(*a, a->ws, BN, stages, tmem_cols);
    return cmgan_check_launch("gemm_rows_tc_kernel");
}
