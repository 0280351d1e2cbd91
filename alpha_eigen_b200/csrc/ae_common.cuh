// Shared device/host helpers for the sm_100a slab S_N kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "alpha_eigen_b200.h"

namespace ae {

constexpr int kL = AE_LANE_CELLS;        // consecutive cells per lane
constexpr int kTile = AE_TILE;           // cells per warp tile
static_assert(kTile == 32 * kL, "tile = one warp x kL cells");

// thread-local error string behind ae_last_error()
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define AE_CUDA(call)                                                             \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return ::ae::cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)

#define AE_TRY(call)                         \
    do {                                     \
        int st_ = (call);                    \
        if (st_ != AE_OK) return st_;        \
    } while (0)

// storage offset inside a tile of the j-th cell (0..7) of `lane`
__host__ __device__ inline int tile_offset(int lane, int j) { return (j >> 1) * 64 + lane * 2 + (j & 1); }

// logical cell of storage index s (inverse of ae_cell_offset)
__host__ __device__ inline int64_t logical_cell(int64_t s)
{
    int64_t tile = s / kTile;
    int r = (int)(s % kTile);
    int k = r >> 6, lane = (r & 63) >> 1, e = r & 1;
    return tile * kTile + lane * kL + 2 * k + e;
}

// block-wide sum of two doubles (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ inline void block_sum2(double &a, double &b, double *smem /* >= 64 doubles */)
{
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) { smem[warp] = a; smem[32 + warp] = b; }
    __syncthreads();
    if (warp == 0) {
        a = lane < nw ? smem[lane] : 0.0;
        b = lane < nw ? smem[32 + lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
}

}  // namespace ae
