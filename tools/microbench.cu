// Developer microbenchmarks for the softmax inner loop of the attention kernel (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu && tools/microbench
// Measures per-SM throughput (results per clock per SM) of: MUFU ex2 f32 / f16x2 / bf16x2, packed fp32 FMA, and the
// polynomial exp2 emulation, with 8 warps per SM (2 per SMSP — what the attention kernel's softmax runs with) and with 16.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 4096

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { return (uint64_t(__float_as_uint(hi)) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ float lo2(uint64_t v) { return __uint_as_float(uint32_t(v)); }
__device__ __forceinline__ float hi2(uint64_t v) { return __uint_as_float(uint32_t(v >> 32)); }

// polynomial 2^x for a pair (Cody-Waite split, degree-3 minimax on [0,1))
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t x) {
    const uint64_t magic = pack2(12582912.f, 12582912.f);
    float a = fmaxf(lo2(x), -127.f), b = fmaxf(hi2(x), -127.f);
    uint64_t xc = pack2(a, b);
    uint64_t t = add2_rm(xc, magic);
    uint64_t fl = sub2(t, magic);
    uint64_t f = sub2(xc, fl);
    uint64_t p = fma2(f, pack2(0.077119089663f, 0.077119089663f), pack2(0.227564394474f, 0.227564394474f));
    p = fma2(p, f, pack2(0.695146143436f, 0.695146143436f));
    p = fma2(p, f, pack2(1.0f, 1.0f));
    uint32_t r0 = uint32_t(p) + (uint32_t(t) << 23);
    uint32_t r1 = uint32_t(p >> 32) + (uint32_t(t >> 32) << 23);
    return (uint64_t(r1) << 32) | r0;
}

template <int MODE>
__global__ void bench(float* out, long long* cycles, float seed) {
    float acc[8];
    uint32_t h[8];
    uint64_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = seed + 0.001f * (threadIdx.x + i);
        h[i] = 0x3c003c00u + threadIdx.x + i;
        q[i] = pack2(acc[i], -acc[i]);
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) acc[i] = ex2_f32(acc[i]) - 1.0f;
            if (MODE == 1) h[i] = ex2_f16x2(h[i]) ^ 0x80008000u;
            if (MODE == 2) h[i] = ex2_bf16x2(h[i]) ^ 0x80008000u;
            if (MODE == 3) q[i] = fma2(q[i], q[(i + 1) & 7], q[(i + 2) & 7]);
            if (MODE == 4) acc[i] = fmaf(acc[i], acc[(i + 1) & 7], acc[(i + 2) & 7]);
            if (MODE == 5) q[i] = ex2_poly2(q[i]) ^ 0x8000000080000000ull;
            if (MODE == 6) {  // 50/50 mix: one MUFU f32 pair + one polynomial pair
                if (i & 1) q[i] = ex2_poly2(q[i]) ^ 0x8000000080000000ull;
                else q[i] = pack2(ex2_f32(lo2(q[i])) - 1.f, ex2_f32(hi2(q[i])) - 1.f);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i] + __uint_as_float(h[i]) + lo2(q[i]) + hi2(q[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, double results_per_op) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    bench<MODE><<<148, threads>>>(out, cyc, -1.5f);
    cudaDeviceSynchronize();
    bench<MODE><<<148, threads>>>(out, cyc, -1.5f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += double(h[i]) / 148;
    double per_clk = double(ITERS) * 8 * threads * results_per_op / c;
    printf("%-28s threads/SM=%4d  cycles=%10.0f  results/clk/SM=%7.2f  (%s)\n", name, threads, c, per_clk, cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int threads : {256, 512}) {
        run<0>("ex2.approx.ftz.f32", threads, 1);
        run<1>("ex2.approx.f16x2", threads, 2);
        run<2>("ex2.approx.ftz.bf16x2", threads, 2);
        run<3>("fma.rn.f32x2", threads, 2);
        run<4>("fma.rn.f32", threads, 1);
        run<5>("poly ex2 (f32x2)", threads, 2);
        run<6>("mix MUFU f32 + poly 50/50", threads, 2);
    }
    return 0;
}
