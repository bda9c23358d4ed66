// FP64 calibration microbenchmarks for B200 (sm_100a): DFMA vs DMMA shapes, plus HBM copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_dfma(double* out, int iters) {
    double a[16];
    double x = threadIdx.x * 1e-9, y = 1.0000001;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], y, x);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
    double c[16][2];
    double a = threadIdx.x * 1e-9, b = 1.0000001;
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, int iters) {  // m16n8k8: A 4 regs, B 2 regs, C 4 regs
    double c[8][4];
    double a0 = threadIdx.x * 1e-9, a1 = 0.5, a2 = 0.25, a3 = 0.125, b0 = 1.0000001, b1 = 0.999;
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {  // m16n8k16: A 8 regs, B 4 regs, C 4 regs
    double c[8][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = 1.0 + i * 1e-7;
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

template <typename F>
float time_it(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms=%d cc=%d.%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.major, p.minor, p.clockRate);
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 4096;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 256, blocks = sms * (warps_per_sm * 32 / threads > 0 ? warps_per_sm * 32 / threads : 1);
        if (warps_per_sm * 32 < threads) { threads = warps_per_sm * 32; blocks = sms; }
        double total_threads = (double)threads * blocks;
        float ms;
        ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, iters); }, 5);
        printf("warps/SM=%2d  DFMA        : %8.2f TFLOP/s (%.3f ms)\n", warps_per_sm, 2.0 * 16 * iters * total_threads / ms / 1e9, ms);
        ms = time_it([&] { k_dmma884<<<blocks, threads>>>(out, iters); }, 5);
        printf("warps/SM=%2d  DMMA m8n8k4  : %8.2f TFLOP/s (%.3f ms)\n", warps_per_sm, 2.0 * 256 * 16 * iters * (total_threads / 32) / ms / 1e9, ms);
        ms = time_it([&] { k_dmma1688<<<blocks, threads>>>(out, iters); }, 5);
        printf("warps/SM=%2d  DMMA m16n8k8 : %8.2f TFLOP/s (%.3f ms)\n", warps_per_sm, 2.0 * 1024 * 8 * iters * (total_threads / 32) / ms / 1e9, ms);
        ms = time_it([&] { k_dmma16816<<<blocks, threads>>>(out, iters); }, 5);
        printf("warps/SM=%2d  DMMA m16n8k16: %8.2f TFLOP/s (%.3f ms)\n", warps_per_sm, 2.0 * 2048 * 8 * iters * (total_threads / 32) / ms / 1e9, ms);
    }
    // sustained (2 s) DMMA + DFMA to see the power-capped clock
    {
        int threads = 256, blocks = sms * 4;
        double total_threads = (double)threads * blocks;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0; float ms = 0;
        do { for (int i = 0; i < 20; i++) k_dmma884<<<blocks, threads>>>(out, iters); n += 20;
             CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); } while (ms < 2000);
        printf("sustained DMMA m8n8k4: %8.2f TFLOP/s over %.0f ms\n", 2.0 * 256 * 16 * iters * (total_threads / 32) * n / ms / 1e9, ms);
        CK(cudaEventRecord(e0)); n = 0;
        do { for (int i = 0; i < 20; i++) k_dfma<<<blocks, threads>>>(out, iters); n += 20;
             CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); } while (ms < 2000);
        printf("sustained DFMA       : %8.2f TFLOP/s over %.0f ms\n", 2.0 * 16 * iters * total_threads * n / ms / 1e9, ms);
    }
    // HBM copy
    {
        size_t n = (size_t)1 << 30;  // 1 Gi doubles /2 -> 8 GiB read + 8 GiB write
        double2 *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
        CK(cudaMemset(a, 1, n * 8));
        float ms = time_it([&] { k_copy<<<sms * 16, 512>>>(a, b, n / 2); }, 5);
        printf("copy kernel 8 GiB: %8.1f GB/s (%.3f ms)\n", 2.0 * n * 8 / ms / 1e6, ms);
        ms = time_it([&] { cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice); }, 5);
        printf("cudaMemcpy D2D 8 GiB: %8.1f GB/s (%.3f ms)\n", 2.0 * n * 8 / ms / 1e6, ms);
        cudaFree(a); cudaFree(b);
    }
    return 0;
}
