// Developer microbenchmark: TMEM load throughput (tcgen05.ld 32x32b.x32) per SM with 4 / 8 / 16 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_tmem tools/microbench_tmem.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define ITERS 2048
__global__ void bench(float* out, long long* cycles, int nld) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + uint32_t((warp >> 2) * 64 % 512);
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        uint32_t r[32];
        for (int k = 0; k < nld; ++k) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(base + uint32_t((k & 1) * 32))
                : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[it & 31]);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {128, 256, 512}) for (int nld : {1, 2, 4}) {
        bench<<<148, threads>>>(out, cyc, nld);
        cudaDeviceSynchronize();
        bench<<<148, threads>>>(out, cyc, nld);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += double(h[i]) / 148;
        double bytes = double(ITERS) * nld * threads * 32 * 4;
        printf("LDTM.x32 threads/SM=%4d loads-in-flight=%d cycles=%9.0f  bytes/clk/SM=%7.1f (%s)\n", threads, nld, c, bytes / c, cudaGetErrorString(e));
    }
    return 0;
}
