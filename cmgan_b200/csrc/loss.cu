// Generator / discriminator losses of the reference Trainer (train.py:124-174) as fused reductions that also emit the
// gradients (the loss is a closed form of its inputs, so no autograd graph is needed), and a flat AdamW step (train.py:63-66).
#include "common.cuh"
#include "../../include/cmgan_b200.h"

namespace {

__device__ __forceinline__ double block_sum_d(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;      // valid in thread 0
}

// spectral terms.  er, ei: estimate (n elements, (B,1,T,F) memory); cr, ci: clean compressed spectrum, element i of batch b at
// cr[b*cb + (i % per)] (planes of a (B,2,T,F) tensor).  acc[0] += sum (er-cr)^2 + (ei-ci)^2;  acc[1] += sum (|e|-|c|)^2.
// d_er/d_ei = w_ri*2(e-c)/n + w_mag*2(|e|-|c|) e/|e| / n;  est_mag / clean_mag optionally written ((B,1,T,F) memory).
__global__ void spec_loss_kernel(const float* __restrict__ er, const float* __restrict__ ei, const float* __restrict__ cr,
                                 const float* __restrict__ ci, long per, long cb, long n, float w_ri, float w_mag, double* __restrict__ acc,
                                 float* __restrict__ d_er, float* __restrict__ d_ei, float* __restrict__ est_mag, float* __restrict__ clean_mag) {
    __shared__ double sm[32];
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double s_ri = 0.0, s_mag = 0.0;
    if (i < n) {
        long b = i / per, r = i - b * per;
        float a = er[i], c = ei[i], p = __ldg(cr + b * cb + r), q = __ldg(ci + b * cb + r);
        float em = sqrtf(a * a + c * c), cm = sqrtf(p * p + q * q);
        float dr = a - p, di = c - q, dm = em - cm;
        s_ri = (double)dr * dr + (double)di * di;
        s_mag = (double)dm * dm;
        float inv_n = 1.0f / (float)n;
        float gm = em > 0.f ? 2.f * w_mag * dm / em * inv_n : 0.f;      // d|e|/de = e/|e| (0 at the origin, as torch.sqrt's backward would be inf*0)
        if (d_er) { d_er[i] = 2.f * w_ri * dr * inv_n + gm * a; d_ei[i] = 2.f * w_ri * di * inv_n + gm * c; }
        if (est_mag) est_mag[i] = em;
        if (clean_mag) clean_mag[i] = cm;
    }
    double t0 = block_sum_d(s_ri, sm);
    double t1 = block_sum_d(s_mag, sm);
    if (threadIdx.x == 0) { atomicAdd(acc, t0); atomicAdd(acc + 1, t1); }
}

// time term: acc[2] += sum |ea - clean|;  d_ea = w_t * sign(ea - clean) / n
__global__ void time_loss_kernel(const float* __restrict__ ea, long lde, const float* __restrict__ clean, long ldc, int B, int L, float w_t,
                                 double* __restrict__ acc, float* __restrict__ d_ea) {
    __shared__ double sm[32];
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long n = (long)B * L;
    double s = 0.0;
    if (i < n) {
        long b = i / L, k = i - b * L;
        float d = ea[b * lde + k] - __ldg(clean + b * ldc + k);
        s = fabs((double)d);
        if (d_ea) d_ea[b * lde + k] = w_t * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / (float)n;
    }
    double t = block_sum_d(s, sm);
    if (threadIdx.x == 0) atomicAdd(acc + 2, t);
}

// loss = w_ri*acc0/n_spec + w_mag*acc1/n_spec + w_t*acc2/n_time + w_gan*mean((fake-1)^2);  d_fake = w_gan*2(fake-1)/B
__global__ void gen_loss_finalize_kernel(const double* __restrict__ acc, double n_spec, double n_time, float w_ri, float w_mag, float w_t,
                                         float w_gan, const float* __restrict__ fake, int B, float* __restrict__ loss, float* __restrict__ d_fake) {
    double g = 0.0;
    if (fake)
        for (int b = 0; b < B; ++b) { double d = (double)fake[b] - 1.0; g += d * d; if (d_fake) d_fake[b] = (float)(w_gan * 2.0 * d / B); }
    loss[0] = (float)(w_ri * acc[0] / n_spec + w_mag * acc[1] / n_spec + w_t * acc[2] / n_time + (fake ? w_gan * g / B : 0.0));
}

// discriminator loss (train.py:168-170): mean((d_max-1)^2) + mean((d_enh - target)^2) and its gradients
__global__ void disc_loss_kernel(const float* __restrict__ d_max, const float* __restrict__ d_enh, const float* __restrict__ target, int B,
                                 float* __restrict__ loss, float* __restrict__ g_max, float* __restrict__ g_enh) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) {
        double a = (double)d_max[b] - 1.0, e = (double)d_enh[b] - (double)target[b];
        s += a * a + e * e;
        g_max[b] = (float)(2.0 * a / B);
        g_enh[b] = (float)(2.0 * e / B);
    }
    loss[0] = (float)(s / B);
}

// d_er += d_mag * er/|e|  (gradient of est_mag = sqrt(er^2 + ei^2) arriving from the discriminator);  d_mag: (B,1,F,T)-shaped with strides
__global__ void mag_bwd_add_kernel(const float* __restrict__ er, const float* __restrict__ ei, const float* __restrict__ d_mag, long gb, long gt,
                                   long gf, int T, int F, long n, float* __restrict__ d_er, float* __restrict__ d_ei) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = (int)(i % F); long bt = i / F; int t = (int)(bt % T); long b = bt / T;
    float a = er[i], c = ei[i];
    float m = sqrtf(a * a + c * c);
    if (m > 0.f) {
        float g = __ldg(d_mag + b * gb + t * gt + f * gf) / m;
        d_er[i] += g * a; d_ei[i] += g * c;
    }
}

// AdamW over flat buffers (torch.optim.AdamW defaults: decoupled weight decay, bias correction)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float bc2, const unsigned long long* __restrict__ step_dev,
                             const float* __restrict__ lr_dev) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (lr_dev) lr = __ldg(lr_dev);        // learning rate as a device scalar: a schedule (StepLR, train.py:248-253) works under CUDA-graph replay
    if (step_dev) { float t = (float)__ldg(step_dev); bc1 = 1.f - powf(b1, t); bc2 = 1.f - powf(b2, t); }
    float gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    pi *= 1.f - lr * wd;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * mi / denom;
}

}  // namespace

// acc: 3 doubles, zeroed by the caller
CMGAN_API int cmgan_spec_loss(const float* er, const float* ei, const float* cr, const float* ci, long long per, long long cb, long long n,
                              float w_ri, float w_mag, double* acc, float* d_er, float* d_ei, float* est_mag, float* clean_mag, void* stream) {
    CMGAN_REQUIRE(er && ei && cr && ci && acc && per > 0, "cmgan_spec_loss: bad arguments");
    if (n == 0) return 0;
    spec_loss_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(er, ei, cr, ci, per, cb, n, w_ri, w_mag, acc, d_er, d_ei, est_mag, clean_mag);
    return cmgan_check_launch("spec_loss_kernel");
}

CMGAN_API int cmgan_time_loss(const float* ea, long long lde, const float* clean, long long ldc, int B, int L, float w_t, double* acc, float* d_ea,
                              void* stream) {
    CMGAN_REQUIRE(ea && clean && acc, "cmgan_time_loss: null pointer");
    if ((long)B * L == 0) return 0;
    time_loss_kernel<<<cdiv((long)B * L, 256), 256, 0, (cudaStream_t)stream>>>(ea, lde, clean, ldc, B, L, w_t, acc, d_ea);
    return cmgan_check_launch("time_loss_kernel");
}

CMGAN_API int cmgan_gen_loss_finalize(const double* acc, double n_spec, double n_time, float w_ri, float w_mag, float w_t, float w_gan,
                                      const float* fake, int B, float* loss, float* d_fake, void* stream) {
    CMGAN_REQUIRE(acc && loss, "cmgan_gen_loss_finalize: null pointer");
    gen_loss_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc, n_spec, n_time, w_ri, w_mag, w_t, w_gan, fake, B, loss, d_fake);
    return cmgan_check_launch("gen_loss_finalize_kernel");
}

CMGAN_API int cmgan_disc_loss(const float* d_max, const float* d_enh, const float* target, int B, float* loss, float* g_max, float* g_enh,
                              void* stream) {
    CMGAN_REQUIRE(d_max && d_enh && target && loss && g_max && g_enh, "cmgan_disc_loss: null pointer");
    disc_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_max, d_enh, target, B, loss, g_max, g_enh);
    return cmgan_check_launch("disc_loss_kernel");
}

CMGAN_API int cmgan_mag_bwd_add(const float* er, const float* ei, const float* d_mag, long long gb, long long gt, long long gf, int B, int T, int F,
                                float* d_er, float* d_ei, void* stream) {
    CMGAN_REQUIRE(er && ei && d_mag && d_er && d_ei, "cmgan_mag_bwd_add: null pointer");
    long n = (long)B * T * F;
    if (n == 0) return 0;
    mag_bwd_add_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(er, ei, d_mag, gb, gt, gf, T, F, n, d_er, d_ei);
    return cmgan_check_launch("mag_bwd_add_kernel");
}

CMGAN_API int cmgan_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, float wd,
                          int step, const unsigned long long* step_dev, const float* lr_dev, void* stream) {
    CMGAN_REQUIRE(p && g && m && v && (step >= 1 || step_dev), "cmgan_adamw: bad arguments");
    if (n == 0) return 0;
    float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adamw_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2, step_dev, lr_dev);
    return cmgan_check_launch("adamw_kernel");
}
