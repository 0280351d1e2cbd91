// Micro-benchmark behind the flux-assembly kernel design: sum P buffers spaced `stride` elements apart
// (one double2 per thread, grid-stride), optionally with the phi read-modify-write and norm of the real
// kernel, on zero-filled or random data, with few or many resident warps.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <bool RMW>
__global__ void __launch_bounds__(256, 6) sum_kernel(const double *base, long long stride_elems, int P, long long n, double *out, double *norm, const double *old)
{
    double num = 0, den = 0;
    for (long long i = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 2LL * gridDim.x * blockDim.x) {
        double2 x = *reinterpret_cast<const double2 *>(base + i);
#pragma unroll 4
        for (int p = 1; p < P; ++p) {
            const double2 y = *reinterpret_cast<const double2 *>(base + p * stride_elems + i);
            x.x += y.x; x.y += y.y;
        }
        if (RMW) {
            const double2 o = *reinterpret_cast<const double2 *>(old + i);
            num += (x.x - o.x) * (x.x - o.x) + (x.y - o.y) * (x.y - o.y);
            den += x.x * x.x + x.y * x.y;
        }
        *reinterpret_cast<double2 *>(out + i) = x;
    }
    if (RMW && num + den == -1.0) norm[0] = num;
}
__global__ void fill_kernel(double *p, long long n, unsigned seed)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long h = (i + seed) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        p[i] = 0.5 + (double)(h & 0xfffff) * 1e-6;
    }
}
int main(int argc, char **argv)
{
    const long long n = 70013440;           // G*Zp of the 1M-cell, 70-group slab
    const int P = 16;
    double *buf, *out, *norm, *big = nullptr, *old;
    const size_t extra = argc > 1 ? (size_t)atof(argv[1]) * (1ull << 30) : 0;   // optional ballast allocation (GiB)
    if (extra) { cudaMalloc(&big, extra); cudaMemset(big, 1, extra); }
    cudaMalloc(&buf, (size_t)(P * n) * 8);
    cudaMalloc(&out, (size_t)n * 8);
    cudaMalloc(&norm, 8);
    cudaMalloc(&old, (size_t)n * 8);
    cudaMemset(old, 0, (size_t)n * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int random = 0; random < 2; ++random) {
        if (random) { fill_kernel<<<4096, 256>>>(buf, P * n, 1); fill_kernel<<<4096, 256>>>(out, n, 7); }
        else { cudaMemset(buf, 0, (size_t)(P * n) * 8); cudaMemset(out, 0, (size_t)n * 8); }
        for (int rmw = 0; rmw < 3; ++rmw)
            for (int blocks : {148 * 6}) {
                float ms = 0;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(a);
                    if (rmw == 1) sum_kernel<true><<<blocks, 256>>>(buf, n, P, n, out, norm, out);       // in place
                    else if (rmw == 2) sum_kernel<true><<<blocks, 256>>>(buf, n, P, n, out, norm, old);  // ping-pong
                    else sum_kernel<false><<<blocks, 256>>>(buf, n, P, n, out, norm, old);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    cudaEventElapsedTime(&ms, a, b);
                }
                printf("ballast=%zuGiB data=%s rmw=%d blocks=%d: %.3f ms  %.0f GB/s\n", extra >> 30, random ? "random" : "zeros", rmw,
                       blocks, ms, (P + 1 + (rmw > 0)) * n * 8 / ms / 1e6);
            }
    }
    return 0;
}
