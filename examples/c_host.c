/* A non-Python host of libcmgan_b200.so (plain C99): enumerates the parameter block of the module-level entry point, sizes the workspace
 * and -- when a CUDA device and a weight file are given -- runs TSCNet.forward (generator.py:174-196, inference mode) through
 * cmgan_tscnet_fwd.  Build:  gcc -std=c99 -Iinclude examples/c_host.c -o c_host -Lcmgan_b200 -lcmgan_b200 -Wl,-rpath,$PWD/cmgan_b200
 * (add -DWITH_CUDA -I/usr/local/cuda/include -L/usr/local/cuda/lib64 -lcudart for the forward).
 * Without WITH_CUDA only the host-side queries run (no GPU needed): this is what tests/test_c_host.py builds and checks. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cmgan_b200.h"

#ifdef WITH_CUDA
#include <cuda_runtime.h>
#endif

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 1, T = argc > 2 ? atoi(argv[2]) : 321, F = 201;
    const int n = cmgan_tscnet_param_count();
    const long long total = cmgan_tscnet_param_floats();
    printf("params %d tensors %lld floats\n", n, total);
    for (int i = 0; i < n; ++i) {
        const char* key;
        long long off, cnt;
        if (cmgan_tscnet_param_info(i, &key, &off, &cnt)) { fprintf(stderr, "%s\n", cmgan_last_error()); return 1; }
        if (i < 3 || i == n - 1) printf("  [%d] %s offset %lld numel %lld\n", i, key, off, cnt);
    }
    const long long ws = cmgan_tscnet_workspace_bytes(B, T, F, 1);
    if (ws < 0) { fprintf(stderr, "%s\n", cmgan_last_error()); return 1; }
    printf("workspace B=%d T=%d tf32: %lld bytes\n", B, T, ws);
    if (cmgan_tscnet_workspace_bytes(B, T, 200, 1) >= 0) { fprintf(stderr, "a 200-bin input must be rejected\n"); return 1; }
    printf("rejected F=200: %s\n", cmgan_last_error());
#ifdef WITH_CUDA
    /* parameters: a raw little-endian float32 dump of the block (cmgan_b200.module_abi.pack_params(...).cpu().numpy().tofile(path)) */
    if (argc > 3) {
        float *params, *x, *re, *im;
        void* wsp;
        float* host = (float*)malloc((size_t)total * 4);
        FILE* f = fopen(argv[3], "rb");
        if (!f || fread(host, 4, (size_t)total, f) != (size_t)total) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
        fclose(f);
        const size_t nx = (size_t)B * 2 * T * F, ny = (size_t)B * T * F;
        cudaMalloc((void**)&params, (size_t)total * 4);
        cudaMalloc((void**)&x, nx * 4);
        cudaMalloc((void**)&re, ny * 4);
        cudaMalloc((void**)&im, ny * 4);
        cudaMalloc(&wsp, (size_t)ws);
        cudaMemcpy(params, host, (size_t)total * 4, cudaMemcpyHostToDevice);
        cudaMemset(x, 0, nx * 4);
        /* contiguous (B, 2, T, F): strides in elements */
        if (cmgan_tscnet_fwd(params, x, 2LL * T * F, (long long)T * F, F, 1, B, T, F, re, im, wsp, ws, 1, 0)) {
            fprintf(stderr, "%s\n", cmgan_last_error());
            return 1;
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "device error\n"); return 1; }
        printf("forward ok\n");
    }
#endif
    return 0;
}
