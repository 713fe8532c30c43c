// Probe: do FP64 tensor (DMMA) and FP64 vector (DFMA) instructions run on independent pipes on this GPU?
// Three register-resident loops per warp: DMMA only, DFMA only, both interleaved; reports TFLOP/s of each.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: DMMA, 1: DFMA, 2: both
__global__ void __launch_bounds__(256) probe(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16], f[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0.0, f[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE != 1) dmma(c[2 * i], c[2 * i + 1], a, b);
            if (MODE != 0) {
                // per DMMA (8 issue clocks of the FP64 unit if shared): give the vector pipe the same amount of work
#pragma unroll
                for (int j = 0; j < 8; j++) f[(i * 2 + j) & 15] = fma(f[(i * 2 + j) & 15], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(double* d, int blocks, int iters, double flop_per_thread_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(d, iters / 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return flop_per_thread_iter * iters * blocks * 256.0 / (ms * 1e-3) / 1e12;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 4, iters = 20000;
    double* d; cudaMalloc(&d, sizeof(double) * blocks * 256);
    // per thread and iteration: 8 DMMA = 8 * 512 flop / 32 lanes = 128 flop; 64 DFMA = 128 flop
    double t0 = run<0>(d, blocks, iters, 128.0);
    double t1 = run<1>(d, blocks, iters, 128.0);
    double t2 = run<2>(d, blocks, iters, 256.0);
    printf("DMMA only %.2f TF | DFMA only %.2f TF | both interleaved %.2f TF (sum of parts %.2f)\n", t0, t1, t2, t0 + t1);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
