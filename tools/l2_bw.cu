// L2 -> shared-memory bandwidth a kernel of the k_matvec_tiled / k_matvec_dmma kind can draw (TMA bulk copies of ket-row
// sized chunks into a shared-memory ring, no arithmetic): the ceiling behind `l1tex__m_xbar2l1tex_read_bytes` in
// profiles/.  Build and run on the GPU box:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_bw tools/l2_bw.cu && /tmp/l2_bw
// Prints GB/s for a working set inside L2 (32 MB) and one far outside (4 GB, HBM), for several chunk sizes, ring depths and
// CTAs per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_pull(const char* __restrict__ src, size_t span, int chunk, int stages, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm);
    unsigned char* ring = sm + 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        const size_t nchunks = span / chunk;
        size_t c = ((size_t)blockIdx.x * 2654435761u) % nchunks;
        int st = 0;
        for (int it = 0; it < iters + stages; ++it) {
            const int ph = ((it - stages) / stages) & 1;       // the stage's use index this wait completes
            if (it >= stages) {
                // wait for the copy that used this stage
                asm volatile(
                    "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(
                        smem_u32(&bar[st])),
                    "r"((unsigned)ph)
                    : "memory");
            }
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar[st])), "r"((unsigned)chunk)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                 smem_u32(ring + (size_t)st * chunk)),
                             "l"(src + c * chunk), "r"((unsigned)chunk), "r"(smem_u32(&bar[st]))
                             : "memory");
                c += gridDim.x;
                if (c >= nchunks) c -= nchunks;
            }
            if (++st == stages) st = 0;
        }
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    char* buf = nullptr;
    const size_t big = (size_t)4 << 30;
    cudaMalloc(&buf, big);
    cudaMemset(buf, 1, big);
    cudaFuncSetAttribute(k_pull, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t spans[2] = {(size_t)32 << 20, big};
    for (int sp = 0; sp < 2; ++sp)
        for (int chunk : {2048, 8192, 32768})
            for (int cps : {1, 2, 4}) {
                const int stages = 96 * 1024 / cps / chunk < 2 ? 2 : (96 * 1024 / cps / chunk > 8 ? 8 : 96 * 1024 / cps / chunk);
                const size_t smem = 128 + (size_t)stages * chunk;
                const int grid = sms * cps;
                const int iters = (int)(((size_t)24 << 30) / grid / chunk > 200000 ? 200000 : ((size_t)24 << 30) / grid / chunk);
                k_pull<<<grid, 32, smem>>>(buf, spans[sp], chunk, stages, 64);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                k_pull<<<grid, 32, smem>>>(buf, spans[sp], chunk, stages, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)grid * iters * chunk;
                printf("%s span %5zu MB chunk %6d B x %d stages, %d CTA/SM: %8.1f GB/s  (%s)\n", sp ? "HBM" : "L2 ", spans[sp] >> 20,
                       chunk, stages, cps, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
    // op-rate ceiling: many small copies per SM (what a ket-row staging kernel with one copy per state and product issues)
    for (int chunk : {256, 512, 1024, 2048})
        for (int cps : {4, 8, 16, 32}) {
            const int stages = 8;
            const size_t smem = 128 + (size_t)stages * chunk;
            const int grid = sms * cps;
            const int iters = 20000;
            k_pull<<<grid, 32, smem>>>(buf, spans[0], chunk, stages, 64);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_pull<<<grid, 32, smem>>>(buf, spans[0], chunk, stages, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double ops = (double)grid * iters;
            printf("ops  chunk %5d B x %d stages, %2d CTA/SM: %7.1f M copies/s per SM (%5.1f cycles at 1.9 GHz), %8.1f GB/s  (%s)\n", chunk,
                   stages, cps, ops / ms / 1e3 / sms, 1.9e9 / (ops / ms * 1e3 / sms), ops * chunk / ms / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
