#include <cstdio>
#include <cuda_runtime.h>
// throughput of scalar vs packed FP32 adds/FMAs on sm_100a, and whether packed ops free issue slots
__global__ void k_scalar(float* out, int iters, float a, float b) {
    float x[16];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = x[i] + (u & 1 ? a : b);
    }
    float r = 0; for (int i = 0; i < 16; i++) r += x[i];
    if (r == 123.456f) out[0] = r;
}
__global__ void k_packed(float* out, int iters, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fadd2_rn(x[i], (u & 1) ? aa : bb);
    }
    float r = 0; for (int i = 0; i < 8; i++) r += x[i].x + x[i].y;
    if (r == 123.456f) out[0] = r;
}
// packed adds interleaved with integer ALU work: do the packed ops leave issue slots for the ALU?
__global__ void k_mix_scalar(float* out, int iters, float a, int m) {
    float x[16]; int y[8];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int i = 0; i < 8; i++) y[i] = threadIdx.x * 3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = x[i] + a;
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = (y[i] ^ m) + i;
        }
    }
    float r = 0; for (int i = 0; i < 16; i++) r += x[i]; int q = 0; for (int i = 0; i < 8; i++) q += y[i];
    if (r == 123.456f || q == 12345) out[0] = r + q;
}
__global__ void k_mix_packed(float* out, int iters, float a, int m) {
    float2 x[8]; int y[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    for (int i = 0; i < 8; i++) y[i] = threadIdx.x * 3 + i;
    float2 aa = make_float2(a, a);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fadd2_rn(x[i], aa);
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = (y[i] ^ m) + i;
        }
    }
    float r = 0; for (int i = 0; i < 8; i++) r += x[i].x + x[i].y; int q = 0; for (int i = 0; i < 8; i++) q += y[i];
    if (r == 123.456f || q == 12345) out[0] = r + q;
}
template <typename K, typename... A> float run(K k, A... args) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 8, 256>>>(args...); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<148 * 8, 256>>>(args...); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 64);
    int iters = 4096;
    double adds = 64.0 * iters * 148 * 8 * 256;
    float ms;
    ms = run(k_scalar, d, iters, 1e-7f, 2e-7f); printf("scalar FADD : %.3f ms  %.1f Tadd/s\n", ms, adds / ms / 1e9);
    ms = run(k_packed, d, iters, 1e-7f, 2e-7f); printf("packed FADD2: %.3f ms  %.1f Tadd/s\n", ms, adds / ms / 1e9);
    ms = run(k_mix_scalar, d, iters, 1e-7f, 5); printf("mix scalar  : %.3f ms (64 FADD + 64 ALU ops per iter)\n", ms);
    ms = run(k_mix_packed, d, iters, 1e-7f, 5); printf("mix packed  : %.3f ms (32 FADD2 + 64 ALU ops per iter)\n", ms);
    return 0;
}
