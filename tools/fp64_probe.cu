// FP64 peak probe for the roofline denominators: DFMA (vector pipe) and DMMA m8n8k4 (mma.sync, the sm_100 FP64 tensor
// path) issue rates, register operands only.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double x)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dadd_kernel(double *out, int iters, double x)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(x));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmul_kernel(double *out, int iters, double x)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(x));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void shfl_kernel(double *out, int iters, double x)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters, double x)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    double a = x, b = 1.0 + 1e-12 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 512, blocks = sms * 4, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 5; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999);
            else if (which == 1) dmma_kernel<<<blocks, threads>>>(out, iters, 0.999999);
            else if (which == 2) dadd_kernel<<<blocks, threads>>>(out, iters, 1e-9);
            else if (which == 3) dmul_kernel<<<blocks, threads>>>(out, iters, 0.999999);
            else shfl_kernel<<<blocks, threads>>>(out, iters / 4, 0.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        // DFMA: 16 FMA per thread per iteration; DMMA: 8 mma per warp per iteration, 8*8*4 FMA each
        const char *names[] = {"dfma", "dmma_m8n8k4", "dadd", "dmul", "shfl64"};
        if (which <= 1) {
            const double fma = which == 0 ? (double)threads * blocks * iters * 16 : (double)threads / 32 * blocks * iters * 8 * 256;
            printf("{\"probe\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f}\n", names[which], best, 2 * fma / (best * 1e-3) / 1e12);
        } else {
            // thread-level operations per clock per SM at 1.92 GHz (shfl64 = one 64-bit value = two SHFL.32)
            const double ops = which == 4 ? (double)threads * blocks * (iters / 4) * 8 : (double)threads * blocks * iters * 16;
            printf("{\"probe\": \"%s\", \"ms\": %.3f, \"thread_ops_per_clk_per_sm\": %.1f}\n", names[which], best,
                   ops / (best * 1e-3) / sms / 1.92e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
