// Microbenchmark: FP64 DFMA vs DMMA (mma.sync.m8n8k4.f64) peak on the bench GPU, plus a
// STREAM-like complex128 copy.  Not part of the product; it provides the FP64 roofline
// denominator that MEASURED_PEAKS.json lacks.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_kernel(double* out, int iters) {
    double c0[4][2] = {{0}}, a = threadIdx.x * 1e-9, b = 1.0000001;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j][0]), "+d"(c0[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int j = 0; j < 4; ++j) s += c0[j][0] + c0[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms %d\n", p.name, p.multiProcessorCount);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {128, 256, 512, 1024}) {
        int blocks = p.multiProcessorCount * (2048 / threads);
        int iters = 20000;
        dfma_kernel<<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)blocks * threads;
        printf("DFMA threads %4d blocks %4d: %.2f TFLOP/s (%.3f ms)\n", threads, blocks, fl / ms * 1e-9, ms);
        dmma_kernel<<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 8 * 8 * 4 * 4 * iters * (double)blocks * (threads / 32);
        printf("DMMA threads %4d blocks %4d: %.2f TFLOP/s (%.3f ms)\n", threads, blocks, fl / ms * 1e-9, ms);
    }
    size_t n = (size_t)1 << 28;  // 4 GiB per buffer of double2
    double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 0, n * 16);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); copy_kernel<<<148 * 16, 512>>>(a, b, n); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("copy double2 grid-stride: %.1f GB/s\n", 2.0 * n * 16 / ms * 1e-6);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
