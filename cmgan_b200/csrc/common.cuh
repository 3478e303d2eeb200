// Shared device/host helpers for the cmgan_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CMGAN_API extern "C" __attribute__((visibility("default")))

// ---- error channel --------------------------------------------------------------------------------
void cmgan_set_error(const char* fmt, ...);
int cmgan_check_launch(const char* what);   // cudaGetLastError() -> 0 / -1 (+ message)

#define CMGAN_REQUIRE(cond, ...)                       \
    do {                                               \
        if (!(cond)) {                                 \
            cmgan_set_error(__VA_ARGS__);              \
            return -1;                                 \
        }                                              \
    } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- math ----------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
// d/dx [x * sigmoid(x)]
__device__ __forceinline__ float dswishf_(float x) {
    float s = sigmoidf_(x);
    return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- dropout --------------------------------------------------------------------------------------
// Counter-based keep decision: a function of (seed, element index) only, so the backward pass (and the
// test-side mask export) regenerate exactly the mask the forward pass applied.  splitmix64 finaliser.
__host__ __device__ __forceinline__ uint32_t cmgan_hash(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
// returns 0 (dropped) or 1/(1-p) (kept); thr = p * 2^32 (0 => dropout disabled => 1)
__host__ __device__ __forceinline__ float cmgan_drop_scale(uint64_t seed, uint64_t idx, uint32_t thr, float inv_keep) {
    if (thr == 0u) return 1.0f;
    return cmgan_hash(seed, idx) >= thr ? inv_keep : 0.0f;
}

// ---- sequence geometry -----------------------------------------------------------------------------
// Activations are channel-last rows (b, t, f) -> row index (b*T + t)*F + f.  A "sequence" is either all
// t for fixed (b, f) (time axis) or all f for fixed (b, t) (frequency axis); both are described by
//   row(s, l) = (s / n_inner) * outer_stride + (s % n_inner) * inner_stride + l * tok_stride
struct SeqGeom {
    int n_seq;          // number of sequences
    int L;              // tokens per sequence
    int n_inner;        // time: F      freq: T
    long outer_stride;  // time: T*F    freq: T*F
    long inner_stride;  // time: 1      freq: F
    long tok_stride;    // time: F      freq: 1
};
__host__ __device__ __forceinline__ long seq_base(const SeqGeom& g, int s) {
    return (long)(s / g.n_inner) * g.outer_stride + (long)(s % g.n_inner) * g.inner_stride;
}
static inline SeqGeom make_seq_geom(int B, int T, int F, int axis /*0 = time, 1 = freq*/) {
    SeqGeom g;
    if (axis == 0) { g.n_seq = B * F; g.L = T; g.n_inner = F; g.outer_stride = (long)T * F; g.inner_stride = 1; g.tok_stride = F; }
    else           { g.n_seq = B * T; g.L = F; g.n_inner = T; g.outer_stride = (long)T * F; g.inner_stride = F; g.tok_stride = 1; }
    return g;
}
