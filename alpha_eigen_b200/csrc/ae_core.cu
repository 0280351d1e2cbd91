// Error reporting, ABI introspection and the cell storage map.
#include <stdarg.h>
#include <string.h>
#include "ae_common.cuh"

namespace ae {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return AE_CUDA_ERROR;
}

}  // namespace ae

extern "C" int ae_abi_version(void) { return AE_ABI_VERSION; }
extern "C" const char *ae_last_error(void) { return ae::g_err; }
extern "C" int64_t ae_padded_cells(int64_t Z) { return (Z + AE_TILE - 1) / AE_TILE * AE_TILE; }
extern "C" int64_t ae_padded_cells_sharded(int64_t Z, int32_t nranks)
{
    const int64_t unit = (int64_t)AE_TILE * (nranks > 0 ? nranks : 1);
    return (Z + unit - 1) / unit * unit;
}
extern "C" int64_t ae_cell_offset(int64_t z)
{
    const int64_t tile = z / AE_TILE;
    const int r = (int)(z % AE_TILE);
    return tile * AE_TILE + ae::tile_offset(r / AE_LANE_CELLS, r % AE_LANE_CELLS);
}

// ---- layout conversion kernels (boundary only; not on the iteration path) -------------------
namespace ae {

// cells: logical [Z][G] <-> storage [G][Zp].  A 32x32 tile goes through shared memory so that both the
// logical side (G fastest) and the storage side (cell fastest) are accessed with unit stride.
// Z: logical cells of the (sub)range, Zp: its storage cells (whole tiles), ld: row stride of the storage array.
__global__ void __launch_bounds__(256) pack_cells_kernel(const double *__restrict__ logical, int64_t Z, int64_t Zp, int64_t ld,
                                                         int G, double *__restrict__ storage, int unpack)
{
    __shared__ double tile[32][33];
    const int64_t s0 = (int64_t)blockIdx.x * 32;        // storage cells of this tile
    const int g0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
    double *lw = const_cast<double *>(logical);
    if (!unpack) {
        for (int r = ty; r < 32; r += 8) {              // r: cell within tile, tx: group -> coalesced over G
            const int64_t s = s0 + r;
            const int64_t z = s < Zp ? logical_cell(s) : Z;
            const int g = g0 + tx;
            tile[r][tx] = (z < Z && g < G) ? logical[z * G + g] : 0.0;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {              // r: group within tile, tx: cell -> coalesced over cells
            const int g = g0 + r;
            const int64_t s = s0 + tx;
            if (g < G && s < Zp) storage[(int64_t)g * ld + s] = tile[tx][r];
        }
    } else {
        for (int r = ty; r < 32; r += 8) {
            const int g = g0 + r;
            const int64_t s = s0 + tx;
            tile[tx][r] = (g < G && s < Zp) ? storage[(int64_t)g * ld + s] : 0.0;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int64_t s = s0 + r;
            const int64_t z = s < Zp ? logical_cell(s) : Z;
            const int g = g0 + tx;
            if (z < Z && g < G) lw[z * G + g] = tile[r][tx];
        }
    }
}

__global__ void pack_psi_kernel(const double *__restrict__ logical, int64_t Z, int64_t Zp, int At, int G,
                                const int32_t *__restrict__ aidx, int A, double *__restrict__ storage, int unpack)
{
    const int64_t n = (int64_t)G * A * Zp;
    double *lw = const_cast<double *>(logical);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / Zp;
        const int g = (int)(row / A), a = (int)(row % A);
        const int64_t z = logical_cell(i % Zp);
        const int64_t j = (z * At + aidx[a]) * G + g;
        if (unpack) { if (z < Z) lw[j] = storage[i]; }
        else storage[i] = z < Z ? logical[j] : 0.0;
    }
}

static int grid_for(int64_t n) { int64_t b = (n + 255) / 256; return (int)(b < 148 * 32 ? b : 148 * 32); }

}  // namespace ae

extern "C" int ae_pack_cells(const double *logical, int64_t Z, int64_t Zp, int32_t G, double *storage, void *stream)
{
    ae::pack_cells_kernel<<<dim3((unsigned)((Zp + 31) / 32), (G + 31) / 32), 256, 0, (cudaStream_t)stream>>>(logical, Z, Zp, Zp,
                                                                                                            G, storage, 0);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

extern "C" int ae_unpack_cells(const double *storage, int64_t Z, int64_t Zp, int32_t G, double *logical, void *stream)
{
    ae::pack_cells_kernel<<<dim3((unsigned)((Zp + 31) / 32), (G + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
        logical, Z, Zp, Zp, G, const_cast<double *>(storage), 1);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

extern "C" int ae_pack_psi(const double *logical, int64_t Z, int64_t Zp, int32_t A_total, int32_t G,
                           const int32_t *angle_index, int32_t A, double *storage, void *stream)
{
    ae::pack_psi_kernel<<<ae::grid_for((int64_t)G * A * Zp), 256, 0, (cudaStream_t)stream>>>(logical, Z, Zp, A_total, G,
                                                                                            angle_index, A, storage, 0);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

extern "C" int ae_unpack_psi(const double *storage, int64_t Z, int64_t Zp, int32_t A_total, int32_t G,
                             const int32_t *angle_index, int32_t A, double *logical, void *stream)
{
    ae::pack_psi_kernel<<<ae::grid_for((int64_t)G * A * Zp), 256, 0, (cudaStream_t)stream>>>(logical, Z, Zp, A_total, G,
                                                                                            angle_index, A,
                                                                                            const_cast<double *>(storage), 1);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

// Cell-range variants for cell-sharded host I/O: `storage` points at the first cell of the range inside rows
// of stride ld; the range holds `cells` storage cells (whole tiles) of which the first `count` are real.
extern "C" int ae_pack_cells_range(const double *logical, int64_t count, int64_t cells, int64_t ld, int32_t G,
                                   double *storage, void *stream)
{
    if (cells % AE_TILE || count > cells) { ae::set_error("ae_pack_cells_range: range must be whole tiles"); return AE_BAD_ARG; }
    ae::pack_cells_kernel<<<dim3((unsigned)((cells + 31) / 32), (G + 31) / 32), 256, 0, (cudaStream_t)stream>>>(logical, count, cells,
                                                                                                               ld, G, storage, 0);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

extern "C" int ae_unpack_cells_range(const double *storage, int64_t count, int64_t cells, int64_t ld, int32_t G,
                                     double *logical, void *stream)
{
    if (cells % AE_TILE || count > cells) { ae::set_error("ae_unpack_cells_range: range must be whole tiles"); return AE_BAD_ARG; }
    ae::pack_cells_kernel<<<dim3((unsigned)((cells + 31) / 32), (G + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
        logical, count, cells, ld, G, const_cast<double *>(storage), 1);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}
