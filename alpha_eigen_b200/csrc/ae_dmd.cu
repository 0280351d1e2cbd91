// DMD on the device: tall-skinny Gram contraction + small dense eigen-solve.
//
// Replaces TimeUnit.compute_alpha_eigenvalues (reference time_unit.py:61-92).  With snapshot rows
// s_0..s_K (K+1 consecutive steps, each a vector of length M in HBM), X = [s_0..s_{K-1}]^T,
// Y = [s_1..s_K]^T.  Everything the reference needs from the thin SVD X = U S V^T follows from
// the (K+1)x(K+1) Gram matrix C = S S^T:   X^T X = C[0:K,0:K] = V S^2 V^T,
//   Atilde = U^T Y V S^-1 = S^-1 V^T (X^T Y) V S^-1,   X^T Y = C[0:K,1:K+1],
// so U (M x K) is never formed.  The Gram contraction is the only O(M)
// step; it is done with FP64 tensor-core DMMA (mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind)
// and reduced deterministically.  The K x K symmetric and non-symmetric eigenproblems go to
// cuSOLVER (plain library calls on tiny matrices).
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "ae_common.cuh"
#include "ae_dmd_small.cuh"

namespace ae {

constexpr int kGN = 128;            // padded snapshot count handled by one CTA (n <= 128)
constexpr int kGK = 32;             // columns (cells) staged per smem stage
constexpr int kGStride = kGK + 4;   // padded row stride in doubles: conflict-free DMMA fragment loads

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Each CTA (8 warps) owns a contiguous range of columns and accumulates the Gram matrix of it -- its UPPER block
// triangle only (the matrix is symmetric): the 16 strips of 8 rows have 16, 15, ..., 1 tiles of 8 x 8 on or above the
// diagonal, so warp w takes strips w and 15 - w, 17 tiles, the same for every warp (the full square would be 32).
// The A fragment (row i0+lane/4, k0+lane%4) and the B fragment (col j0+lane/4, k0+lane%4) are the same access pattern
// on the staged tile, because B = S^T.
__global__ void __launch_bounds__(256) gram_kernel(const double *__restrict__ snap, int64_t ld, int64_t len, int n,
                                                   int64_t cols_per_cta, double *__restrict__ partial)
{
    extern __shared__ double gram_tiles[];              // two stages of kGN x kGStride: the next one loads while this one is used
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t c_begin = (int64_t)blockIdx.x * cols_per_cta;
    const int64_t c_end = min(len, c_begin + cols_per_cta);
    const int s0 = warp, s1 = 15 - warp;                // this warp's two row strips (s0 < s1)
    double acc[2][16][2];                               // [strip][column tile]: tiles b >= strip are live
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 16; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // stage rows 0..127 x kGK columns (zero beyond n / c_end) with 8-byte asynchronous copies: 32 consecutive doubles per row
    auto stage = [&](int buf, int64_t c0) {
        double *tile = gram_tiles + buf * (kGN * kGStride);
        for (int idx = tid; idx < kGN * kGK; idx += 256) {
            const int r = idx / kGK, k = idx % kGK;
            const int64_t c = c0 + k;
            const bool live = r < n && c < c_end;
            const double *src = live ? snap + (int64_t)r * ld + c : snap;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + r * kGStride + k);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(live ? 8 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (c_begin < c_end) stage(0, c_begin);
    int it = 0;
    for (int64_t c0 = c_begin; c0 < c_end; c0 += kGK, ++it) {
        const bool more = c0 + kGK < c_end;
        if (more) {
            stage((it + 1) & 1, c0 + kGK);              // its previous user finished at the barrier that ended iteration it-1
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double *tile = gram_tiles + (it & 1) * (kGN * kGStride);
#pragma unroll
        for (int k0 = 0; k0 < kGK; k0 += 4) {
            const double a0 = tile[(s0 * 8 + (lane >> 2)) * kGStride + k0 + (lane & 3)];
            const double a1 = tile[(s1 * 8 + (lane >> 2)) * kGStride + k0 + (lane & 3)];
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                if (b >= s0) {                          // warp-uniform
                    const double bf = tile[(b * 8 + (lane >> 2)) * kGStride + k0 + (lane & 3)];
                    dmma_8x8x4(acc[0][b][0], acc[0][b][1], a0, bf);
                    if (b >= s1) dmma_8x8x4(acc[1][b][0], acc[1][b][1], a1, bf);
                }
            }
        }
        __syncthreads();
    }
    // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}; only the live tiles are written (the reduction mirrors them)
    double *out = partial + (size_t)blockIdx.x * kGN * kGN;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int strip = a ? s1 : s0;
            if (b < strip) continue;
            const int r = strip * 8 + (lane >> 2), c = b * 8 + 2 * (lane & 3);
            out[r * kGN + c] = acc[a][b][0];
            out[r * kGN + c + 1] = acc[a][b][1];
        }
}

__global__ void gram_reduce_kernel(const double *__restrict__ partial, int nparts, int n, double *__restrict__ gram)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    int r = idx / n, c = idx % n;
    if ((r >> 3) > (c >> 3)) { const int t = r; r = c; c = t; }     // lower block triangle: the mirror tile
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * kGN * kGN + r * kGN + c];
    gram[idx] = s;
}

// Atilde = S^-1 V^T Cxy V S^-1 for the r retained modes (column-major output for cuSOLVER).
// V: eigenvectors of Cxx as returned by syevd (column-major K x K, eigenvalues ascending), so
// retained mode j (descending sigma) is column K-1-j.
__global__ void atilde_kernel(const double *__restrict__ gram, int n /*K+1*/, const double *__restrict__ V,
                              const double *__restrict__ inv_sigma, int K, int r, double *__restrict__ T /*K x r*/,
                              double *__restrict__ At /* r x r col-major */)
{
    // T = Cxy V_r   (K x r):  T[i][j] = sum_k C[i][k+1] V[k][col(j)]
    for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < K * r; idx += blockDim.x * gridDim.x) {
        const int i = idx / r, j = idx % r, col = K - 1 - j;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s = fma(gram[i * n + k + 1], V[(size_t)col * K + k], s);
        T[idx] = s;
    }
    __threadfence();
    __syncthreads();    // single-block launch: T complete
    for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {
        const int i = idx % r, j = idx / r, coli = K - 1 - i;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s = fma(V[(size_t)coli * K + k], T[k * r + j], s);
        At[idx] = inv_sigma[i] * s * inv_sigma[j];
    }
}

// ---------------------------------------------------------------------------------------------
// R-only TSQR.  The Gram route squares the condition number: singular values below ~1e-8 sigma_1
// drown in rounding, while the reference's rank rule (time_unit.py:75-76) looks 1e-13 deep.  A
// Householder QR of the tall matrix T = S^T (len x n) is backward stable, and only its n x n
// triangular factor is needed (Q is never formed).
//
// Each CTA keeps the running triangle R[n][ldw] in shared memory and the current block of kTsqrB rows of T in
// REGISTERS: 512 threads = 128 column slots x 4 row parts, thread (j, part) holds rows 4b + part (b = 0..23) of
// column j.  One "append" runs n Householder steps; step k only touches row k of R and the block rows (the rest
// of column k is already zero):
//   the four threads of column k publish the reflector x (their 96 block entries) in a small shared buffer, sum
//   its square and derive (tau, 1/(alpha - beta), beta) exactly as LAPACK's dlarfg does;   -- one barrier --
//   every thread of a column j > k forms its part of x.A[:,j] from registers, the four parts meet with two
//   shuffles, row k of R and the 24 register rows are updated.
// The reflector buffer is double buffered by the parity of k, so one barrier per step suffices.  (The first
// version kept the block in shared memory with one thread per column: every step streamed the whole trailing block
// through shared memory twice, 4.3 us per step at 4 warps per SM; this one takes ~0.5 us.)
constexpr int kTsqrParts = 4, kTsqrRegRows = 24, kTsqrB = kTsqrParts * kTsqrRegRows;    // 96 rows per block
constexpr int kTsqrThreads = kGN * kTsqrParts;                                          // 512
static_assert(kGN == 128 && kTsqrThreads == 512, "thread layout assumes 128 column slots");

template <bool ROW_MAJOR>
__global__ void __launch_bounds__(kTsqrThreads, 1) tsqr_kernel(const double *__restrict__ in, int64_t ld, int64_t len, int n,
                                                               int64_t rows_per_cta, double *__restrict__ r_out)
{
    extern __shared__ double R[];                   // [n][ldw], then x[2][kTsqrB], then hh[2][4]
    const int ldw = n | 1;
    double *xbuf = R + n * ldw;
    double *hbuf = xbuf + 2 * kTsqrB;
    const int tid = threadIdx.x, lane = tid & 31;
    const int part = lane & 3, j = (tid >> 5) * 8 + (lane >> 2);       // the 4 parts of a column are adjacent lanes
    const unsigned quad = 0xFu << (lane & ~3);
    const int64_t m_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t m_end = min(len, m_begin + rows_per_cta);
    for (int i = tid; i < n * ldw; i += blockDim.x) R[i] = 0.0;
    for (int64_t m0 = m_begin; m0 < m_end; m0 += kTsqrB) {
        double a[kTsqrRegRows];
#pragma unroll
        for (int b = 0; b < kTsqrRegRows; ++b) {
            const int64_t m = m0 + 4 * b + part;
            a[b] = (j < n && m < m_end) ? __ldg(ROW_MAJOR ? in + m * ld + j : in + (int64_t)j * ld + m) : 0.0;
        }
        __syncthreads();                            // R zeroed / the previous block's last step is complete
        for (int k = 0; k < n; ++k) {
            double *x = xbuf + (k & 1) * kTsqrB, *hh = hbuf + (k & 1) * 4;
            if (j == k) {                           // dlarfg on (R[k][k], block entries of column k)
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int b = 0; b < kTsqrRegRows; b += 4) {
                    s0 = fma(a[b], a[b], s0); s1 = fma(a[b + 1], a[b + 1], s1);
                    s2 = fma(a[b + 2], a[b + 2], s2); s3 = fma(a[b + 3], a[b + 3], s3);
                }
#pragma unroll
                for (int b = 0; b < kTsqrRegRows; ++b) x[4 * b + part] = a[b];
                double ss = (s0 + s1) + (s2 + s3);
                ss += __shfl_xor_sync(quad, ss, 1);
                ss += __shfl_xor_sync(quad, ss, 2);
                if (part == 0) {
                    const double alpha = R[k * ldw + k];
                    if (ss == 0.0) {
                        hh[0] = 0.0; hh[1] = 0.0;
                    } else {
                        // dlarfg's beta = -sign(alpha) |(alpha, x)|, tau = (beta - alpha)/beta = 1 + |alpha|/norm and
                        // 1/(alpha - beta) = sign(alpha)/(|alpha| + norm) from one rsqrt and one reciprocal: this scalar
                        // chain is what the other 15 warps wait for at the barrier of every step
                        const double n2 = fma(alpha, alpha, ss), ri = rsqrt(n2), nrm = n2 * ri, aa = fabs(alpha);
                        hh[0] = fma(aa, ri, 1.0);
                        hh[1] = copysign(__drcp_rn(aa + nrm), alpha);
                        R[k * ldw + k] = -copysign(nrm, alpha);
                    }
                }
            }
            __syncthreads();
            const double tau = hh[0], inv = hh[1];
            if (j > k && j < n && tau != 0.0) {
                double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                for (int b = 0; b < kTsqrRegRows; b += 4) {
                    d0 = fma(x[4 * b + part], a[b], d0); d1 = fma(x[4 * b + 4 + part], a[b + 1], d1);
                    d2 = fma(x[4 * b + 8 + part], a[b + 2], d2); d3 = fma(x[4 * b + 12 + part], a[b + 3], d3);
                }
                double dot = (d0 + d1) + (d2 + d3);
                dot += __shfl_xor_sync(quad, dot, 1);
                dot += __shfl_xor_sync(quad, dot, 2);
                const double rkj = R[k * ldw + j];
                const double w = tau * fma(inv, dot, rkj);
                const double wi = w * inv;
                __syncwarp(quad);                   // all four parts have read R[k][j]
                if (part == 0) R[k * ldw + j] = rkj - w;
#pragma unroll
                for (int b = 0; b < kTsqrRegRows; ++b) a[b] = fma(-wi, x[4 * b + part], a[b]);
            }
        }
    }
    __syncthreads();
    double *out = r_out + (size_t)blockIdx.x * n * n;
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, jj = idx % n;
        out[idx] = jj >= i ? R[i * ldw + jj] : 0.0;
    }
}

static size_t tsqr_smem(int n) { return ((size_t)n * (n | 1) + 2 * kTsqrB + 8) * sizeof(double); }

static int tsqr_launch(const double *in, int64_t ld, int64_t len, int n, bool row_major, int64_t rows_per_cta, int ctas,
                       double *r_out, cudaStream_t st)
{
    const size_t smem = tsqr_smem(n);
    if (row_major) {
        AE_CUDA(cudaFuncSetAttribute((const void *)tsqr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tsqr_kernel<true><<<ctas, kTsqrThreads, smem, st>>>(in, ld, len, n, rows_per_cta, r_out);
    } else {
        AE_CUDA(cudaFuncSetAttribute((const void *)tsqr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tsqr_kernel<false><<<ctas, kTsqrThreads, smem, st>>>(in, ld, len, n, rows_per_cta, r_out);
    }
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

// column-major copies for cuSOLVER: Ax = R[:, 0:K]  ((K+1) x K), Ry = R[:, 1:K+1]
__global__ void split_r_kernel(const double *__restrict__ R, int n, double *__restrict__ Ax, double *__restrict__ Ry)
{
    const int K = n - 1;
    for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < n * K; idx += blockDim.x * gridDim.x) {
        const int i = idx % n, j = idx / n;
        Ax[idx] = R[i * n + j];
        Ry[idx] = R[i * n + j + 1];
    }
}

// Atilde = U_r^T Ry V_r S_r^-1  (time_unit.py:85-89), column-major r x r.  Single block.
__global__ void atilde_from_svd_kernel(const double *__restrict__ U, int ldu, const double *__restrict__ VT, int ldvt,
                                       const double *__restrict__ Ry, int n, const double *__restrict__ inv_sigma, int r,
                                       double *__restrict__ T /* n x r */, double *__restrict__ At)
{
    const int K = n - 1;
    for (int idx = threadIdx.x; idx < n * r; idx += blockDim.x) {       // T[i][b] = sum_j Ry[i][j] V[j][b] / s_b
        const int i = idx % n, b = idx / n;
        double s = 0.0;
        for (int j = 0; j < K; ++j) s = fma(Ry[i + j * n], VT[b + j * ldvt], s);
        T[idx] = s * inv_sigma[b];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {       // At[a][b] = sum_i U[i][a] T[i][b]
        const int a = idx % r, b = idx / r;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = fma(U[i + a * ldu], T[i + b * n], s);
        At[idx] = s;
    }
}

// One cuSOLVER handle per host thread (creation costs milliseconds and synchronises the device, which would
// serialise the concurrent member solves of an ensemble).
struct SolverCtx {
    cusolverDnHandle_t h = nullptr;
    cusolverDnParams_t params = nullptr;
    ~SolverCtx()
    {
        if (params) cusolverDnDestroyParams(params);
        if (h) cusolverDnDestroy(h);
    }
};
static thread_local SolverCtx t_solver;

static int cusolver_fail(cusolverStatus_t st, const char *what)
{
    set_error("%s failed with cusolver status %d", what, (int)st);
    return AE_CUDA_ERROR;
}
#define AE_CUSOLVER(call)                                              \
    do {                                                               \
        cusolverStatus_t s_ = (call);                                  \
        if (s_ != CUSOLVER_STATUS_SUCCESS) { status = cusolver_fail(s_, #call); goto done; } \
    } while (0)
static int solver_ctx(cudaStream_t st, cusolverDnHandle_t *h, cusolverDnParams_t *params)
{
    int status = AE_OK;
    if (!t_solver.h) {
        cusolverStatus_t s_ = cusolverDnCreate(&t_solver.h);
        if (s_ != CUSOLVER_STATUS_SUCCESS) return cusolver_fail(s_, "cusolverDnCreate");
        s_ = cusolverDnCreateParams(&t_solver.params);
        if (s_ != CUSOLVER_STATUS_SUCCESS) return cusolver_fail(s_, "cusolverDnCreateParams");
    }
    cusolverStatus_t s2 = cusolverDnSetStream(t_solver.h, st);
    if (s2 != CUSOLVER_STATUS_SUCCESS) return cusolver_fail(s2, "cusolverDnSetStream");
    *h = t_solver.h;
    *params = t_solver.params;
    return status;
}

#define AE_CUDA_G(call)                                                \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) { status = cuda_fail(e_, #call, __FILE__, __LINE__); goto done; } \
    } while (0)

// ---- windows wider than kGN snapshots --------------------------------------------------------------------------
// The hand-written Gram and TSQR kernels keep one 128-column operand per CTA (registers / shared memory), which covers
// the benchmark windows (K + 1 = 116 at 200 steps).  The reference takes any step count (time_unit.py:62-72); wider
// windows go through plain library calls on the same data: the snapshot rows [n][ld] ARE the column-major len x n tall
// matrix, so its R factor is cusolverDnDgeqrf on a scratch copy (blocked Householder, backward stable like the TSQR
// kernel) and its Gram matrix one cublasDgemm.  Slower per column than the kernels; the answers obey the same tests.
static thread_local cublasHandle_t t_blas = nullptr;
static int blas_ctx(cudaStream_t st, cublasHandle_t *h)
{
    if (!t_blas && cublasCreate(&t_blas) != CUBLAS_STATUS_SUCCESS) { set_error("cublasCreate failed"); return AE_CUDA_ERROR; }
    if (cublasSetStream(t_blas, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return AE_CUDA_ERROR; }
    *h = t_blas;
    return AE_OK;
}

// upper triangle of a column-major m x n factored matrix (ld = lda) -> row-major n x n, zeros below the diagonal
__global__ void extract_r_kernel(const double *__restrict__ a, int64_t lda, int n, int64_t rows, double *__restrict__ r)
{
    for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < n * n; idx += blockDim.x * gridDim.x) {
        const int i = idx / n, j = idx % n;
        r[idx] = (j >= i && i < rows) ? a[(int64_t)j * lda + i] : 0.0;
    }
}

// R factor of the column-major tall matrix `a` (len x n, leading dimension lda), which is overwritten.
static int geqrf_r(double *a, int64_t lda, int64_t len, int n, double *r_out, cudaStream_t st)
{
    int status = AE_OK;
    cusolverDnHandle_t h = nullptr;
    cusolverDnParams_t params = nullptr;
    double *tau = nullptr, *work = nullptr;
    int *dinfo = nullptr, lwork = 0, info = 0;
    if ((status = solver_ctx(st, &h, &params)) != AE_OK) return status;
    if (len > 2147483647LL || lda > 2147483647LL) { set_error("ae_tsqr: window wider than %d snapshots needs fewer than 2^31 rows per rank", kGN); return AE_BAD_ARG; }
    AE_CUDA_G(cudaMallocAsync((void **)&tau, sizeof(double) * n, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dinfo, sizeof(int), st));
    AE_CUSOLVER(cusolverDnDgeqrf_bufferSize(h, (int)len, n, a, (int)lda, &lwork));
    AE_CUDA_G(cudaMallocAsync((void **)&work, sizeof(double) * (size_t)lwork, st));
    AE_CUSOLVER(cusolverDnDgeqrf(h, (int)len, n, a, (int)lda, tau, work, lwork, dinfo));
    AE_CUDA_G(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaStreamSynchronize(st));
    if (info != 0) { set_error("geqrf info=%d", info); status = AE_CUDA_ERROR; goto done; }
    extract_r_kernel<<<64, 256, 0, st>>>(a, lda, n, len, r_out);
    AE_CUDA_G(cudaGetLastError());
done:
    cudaFreeAsync(tau, st); cudaFreeAsync(work, st); cudaFreeAsync(dinfo, st);
    return status;
}

// row-major [rows][n] -> column-major rows x n (the stacked triangles of a merge are row-major)
__global__ void transpose_rows_kernel(const double *__restrict__ in, int64_t rows, int n, double *__restrict__ out)
{
    for (int64_t idx = threadIdx.x + (int64_t)blockIdx.x * blockDim.x; idx < rows * n; idx += (int64_t)blockDim.x * gridDim.x) {
        const int64_t i = idx / n;
        const int j = (int)(idx % n);
        out[(int64_t)j * rows + i] = in[idx];
    }
}

}  // namespace ae

using namespace ae;

// The DMD entry points take their scratch from the stream-ordered allocator.  By default the pool hands everything back to the
// driver at the next synchronisation, so every call pays the allocation again (~5 ms for the 38 MB of Gram partials): keep up
// to 256 MB cached.
static void keep_pool_memory()
{
    static bool done = false;
    if (done) return;
    done = true;
    int dev = 0;
    cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) { cudaGetLastError(); return; }
    unsigned long long current = 0, want = 256ull << 20;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &current) == cudaSuccess && current < want)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
    cudaGetLastError();
}

extern "C" int ae_gram(const double *snap, int64_t ld, int64_t len, int32_t n, double *gram, void *stream)
{
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    if (!snap || !gram || n < 1 || len < 1 || ld < len) { set_error("ae_gram: bad argument (n=%d)", n); return AE_BAD_ARG; }
    if (n > kGN) {                                      // wide window: one library GEMM, G = B^T B with B = snap as len x n column-major
        cublasHandle_t bh = nullptr;
        AE_TRY(blas_ctx(st, &bh));
        if (len > 2147483647LL || ld > 2147483647LL) { set_error("ae_gram: window wider than %d snapshots needs fewer than 2^31 rows per rank", kGN); return AE_BAD_ARG; }
        const double one = 1.0, zero = 0.0;
        if (cublasDgemm(bh, CUBLAS_OP_T, CUBLAS_OP_N, n, n, (int)len, &one, snap, (int)ld, snap, (int)ld, &zero, gram, n) != CUBLAS_STATUS_SUCCESS) {
            set_error("cublasDgemm failed");
            return AE_CUDA_ERROR;
        }
        return AE_OK;
    }
    int nparts = 148 * 2;
    int64_t cols = (len + nparts - 1) / nparts;
    cols = (cols + kGK - 1) / kGK * kGK;
    nparts = (int)((len + cols - 1) / cols);
    constexpr int gram_smem = 2 * kGN * kGStride * (int)sizeof(double);
    static bool gram_attr = false;
    if (!gram_attr) {
        AE_CUDA(cudaFuncSetAttribute((const void *)gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem));
        gram_attr = true;
    }
    double *partial = nullptr;
    AE_CUDA(cudaMallocAsync((void **)&partial, (size_t)nparts * kGN * kGN * sizeof(double), st));
    gram_kernel<<<nparts, 256, gram_smem, st>>>(snap, ld, len, n, cols, partial);
    gram_reduce_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(partial, nparts, n, gram);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(partial, st);
    AE_CUDA(e);
    return AE_OK;
}

extern "C" int ae_dmd_from_gram(const double *gram_dev, int32_t K, double dt, double rank_tol, double *alpha_re,
                                double *alpha_im, double *sigma_out, int32_t *rank, void *stream)
{
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    if (!gram_dev || K < 1 || !alpha_re || !alpha_im || !rank) {
        set_error("ae_dmd_from_gram: bad argument");
        return AE_BAD_ARG;
    }
    int status = AE_OK;
    const int n = K + 1;
    cusolverDnHandle_t h = nullptr;
    cusolverDnParams_t params = nullptr;
    double *dV = nullptr, *dW = nullptr, *dwork = nullptr, *dT = nullptr, *dAt = nullptr, *dInv = nullptr;
    double *dEig = nullptr;
    void *dbuf = nullptr, *hbuf = nullptr;
    int *dinfo = nullptr;
    int lwork = 0, info = 0, r = 0;
    size_t dbytes = 0, hbytes = 0;
    double *lam = (double *)malloc(sizeof(double) * K), *sig = (double *)malloc(sizeof(double) * K),
           *inv = (double *)malloc(sizeof(double) * K), *W = (double *)malloc(sizeof(double) * 2 * K);

    if ((status = solver_ctx(st, &h, &params)) != AE_OK) goto done;
    AE_CUDA_G(cudaMallocAsync((void **)&dV, sizeof(double) * K * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dW, sizeof(double) * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dinfo, sizeof(int), st));
    // Cxx = gram[0:K,0:K] (symmetric, so row/column-major agree)
    AE_CUDA_G(cudaMemcpy2DAsync(dV, sizeof(double) * K, gram_dev, sizeof(double) * n, sizeof(double) * K, K,
                                cudaMemcpyDeviceToDevice, st));
    AE_CUSOLVER(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, K, dV, K, dW, &lwork));
    AE_CUDA_G(cudaMallocAsync((void **)&dwork, sizeof(double) * (size_t)lwork, st));
    AE_CUSOLVER(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, K, dV, K, dW, dwork, lwork, dinfo));
    AE_CUDA_G(cudaMemcpyAsync(lam, dW, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaStreamSynchronize(st));
    if (info != 0) { set_error("syevd info=%d", info); status = AE_CUDA_ERROR; goto done; }
    {
        // singular values of X, descending; rank rule of time_unit.py:75-76
        double total = 0.0, run = 0.0;
        // The Gram matrix carries ~K*eps*sigma_1^2 of rounding, i.e. its small eigenvalues are noise
        // below sigma ~ 1e-7 sigma_1: such values are excluded before the reference's rule is applied.
        for (int j = 0; j < K; ++j) sig[j] = sqrt(fmax(lam[K - 1 - j], 0.0));
        int kept = 0;
        for (int j = 0; j < K; ++j) { if (sig[j] > 1e-7 * sig[0]) { kept = j + 1; total += sig[j]; } else break; }
        for (int j = 0; j < kept; ++j) {
            run += sig[j];
            if ((1.0 - run / total > rank_tol || j + 1 < kept) && sig[j] > 0.0) r = j + 1; else break;
        }
        if (r == 0 && kept > 0) r = 1;
        if (r < 1) { set_error("ae_dmd_from_gram: no singular value survives the rank rule"); status = AE_BAD_ARG; goto done; }
        for (int j = 0; j < r; ++j) inv[j] = 1.0 / sig[j];
        if (sigma_out) memcpy(sigma_out, sig, sizeof(double) * K);
    }
    AE_CUDA_G(cudaMallocAsync((void **)&dT, sizeof(double) * K * r, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dAt, sizeof(double) * r * r, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dInv, sizeof(double) * r, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dEig, sizeof(double) * 2 * r, st));
    AE_CUDA_G(cudaMemcpyAsync(dInv, inv, sizeof(double) * r, cudaMemcpyHostToDevice, st));
    atilde_kernel<<<1, 1024, 0, st>>>(gram_dev, n, dV, dInv, K, r, dT, dAt);
    AE_CUDA_G(cudaGetLastError());
    // eigenvalues of the real non-symmetric r x r matrix (time_unit.py:90)
    AE_CUSOLVER(cusolverDnXgeev_bufferSize(h, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, r, CUDA_R_64F,
                                           dAt, r, CUDA_C_64F, dEig, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1,
                                           CUDA_R_64F, &dbytes, &hbytes));
    AE_CUDA_G(cudaMallocAsync(&dbuf, dbytes ? dbytes : 8, st));
    hbuf = malloc(hbytes ? hbytes : 8);
    AE_CUSOLVER(cusolverDnXgeev(h, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, r, CUDA_R_64F, dAt, r,
                                CUDA_C_64F, dEig, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dbuf,
                                dbytes, hbuf, hbytes, dinfo));
    AE_CUDA_G(cudaMemcpyAsync(W, dEig, sizeof(double) * 2 * r, cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaStreamSynchronize(st));
    if (info != 0) { set_error("geev info=%d", info); status = AE_CUDA_ERROR; goto done; }
    for (int j = 0; j < r; ++j) {
        // alpha = (1 - 1/lambda)/dt  (time_unit.py:91), complex arithmetic
        const double re = W[2 * j], im = W[2 * j + 1], d = re * re + im * im;
        alpha_re[j] = (1.0 - re / d) / dt;
        alpha_im[j] = (im / d) / dt;
    }
    *rank = r;
done:
    cudaFreeAsync(dV, st); cudaFreeAsync(dW, st); cudaFreeAsync(dwork, st); cudaFreeAsync(dT, st); cudaFreeAsync(dAt, st); cudaFreeAsync(dInv, st); cudaFreeAsync(dEig, st);
    cudaFreeAsync(dbuf, st); cudaFreeAsync(dinfo, st);
    free(hbuf); free(lam); free(sig); free(inv); free(W);
    return status;
}

// Triangles are merged by a tree of further QRs (stacked triangles are just another tall matrix, row-major);
// fan-in 6 keeps every CTA of a level at ~7 blocks of rows.
constexpr int kTsqrFanIn = 6;

static int tsqr_merge_tree(double *in, double *scratch, int count, int n, double *r_out, cudaStream_t st)
{
    while (count > 1) {
        const int next = count <= kTsqrFanIn + 2 ? 1 : (count + kTsqrFanIn - 1) / kTsqrFanIn;
        const int per = (count + next - 1) / next;
        double *dst = next == 1 ? r_out : scratch;
        AE_TRY(tsqr_launch(in, n, (int64_t)count * n, n, true, (int64_t)per * n, next, dst, st));
        count = next;
        double *t = in; in = scratch; scratch = t;
    }
    return AE_OK;
}

extern "C" int ae_tsqr(const double *snap, int64_t ld, int64_t len, int32_t n, double *r_out, void *stream)
{
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    if (!snap || !r_out || n < 1 || len < 1) { set_error("ae_tsqr: bad argument (n=%d)", n); return AE_BAD_ARG; }
    if (n > kGN) {                                      // wide window: library QR on a scratch copy of the window
        double *copy = nullptr;
        AE_CUDA(cudaMallocAsync((void **)&copy, sizeof(double) * (size_t)len * n, st));
        cudaError_t e = cudaMemcpy2DAsync(copy, sizeof(double) * len, snap, sizeof(double) * ld, sizeof(double) * len, n,
                                          cudaMemcpyDeviceToDevice, st);
        int rc = e == cudaSuccess ? geqrf_r(copy, len, len, n, r_out, st) : cuda_fail(e, "cudaMemcpy2DAsync", __FILE__, __LINE__);
        cudaFreeAsync(copy, st);
        return rc;
    }
    int dev = 0, ncta = 148;
    AE_CUDA(cudaGetDevice(&dev));
    AE_CUDA(cudaDeviceGetAttribute(&ncta, cudaDevAttrMultiProcessorCount, dev));
    int64_t rows = (len + ncta - 1) / ncta;
    rows = (rows + kTsqrB - 1) / kTsqrB * kTsqrB;
    ncta = (int)((len + rows - 1) / rows);
    double *parts = nullptr;
    const size_t tri = (size_t)n * n;
    AE_CUDA(cudaMallocAsync((void **)&parts, ((size_t)ncta + (ncta + kTsqrFanIn - 1) / kTsqrFanIn) * tri * sizeof(double), st));
    int status = tsqr_launch(snap, ld, len, n, false, rows, ncta, ncta == 1 ? r_out : parts, st);
    if (status == AE_OK) status = tsqr_merge_tree(parts, parts + (size_t)ncta * tri, ncta, n, r_out, st);
    cudaFreeAsync(parts, st);
    return status;
}

extern "C" int ae_tsqr_merge(const double *r_stack, int32_t count, int32_t n, double *r_out, void *stream)
{
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    if (!r_stack || !r_out || count < 1 || n < 1) { set_error("ae_tsqr_merge: bad argument"); return AE_BAD_ARG; }
    if (n > kGN) {                                      // wide window: QR of the stacked triangles through the library
        const int64_t rows = (int64_t)count * n;
        double *cm = nullptr;
        AE_CUDA(cudaMallocAsync((void **)&cm, sizeof(double) * (size_t)rows * n, st));
        transpose_rows_kernel<<<256, 256, 0, st>>>(r_stack, rows, n, cm);
        cudaError_t e = cudaGetLastError();
        int rc = e == cudaSuccess ? geqrf_r(cm, rows, rows, n, r_out, st) : cuda_fail(e, "transpose_rows_kernel", __FILE__, __LINE__);
        cudaFreeAsync(cm, st);
        return rc;
    }
    return tsqr_launch(r_stack, n, (int64_t)count * n, n, true, (int64_t)count * n, 1, r_out, st);
}

// ---- the small dense part as ONE block per window (ae_dmd_small.cuh) ----------------------------------------------
// A thread block seen as groups of L lanes (L = 32: warps).  A group plays the "warp" of ae_dmd_small.cuh: it owns one
// column pair of a Jacobi round, so with L = 8 one warp instruction works on four pairs (the FP64 pipe issues a warp
// instruction in two cycles whatever the number of active lanes, and the rotation scalars are computed by every lane).
template <int L>
struct GroupTeam {
    __device__ int lane() const { return threadIdx.x & (L - 1); }
    __device__ int lanes() const { return L; }
    __device__ int warp() const { return threadIdx.x / L; }
    __device__ int warps() const { return blockDim.x / L; }
    __device__ double sum(double x) const
    {
        const unsigned mask = L == 32 ? 0xffffffffu : ((1u << L) - 1u) << (threadIdx.x & 31 & ~(L - 1));
#pragma unroll
        for (int o = L / 2; o; o >>= 1) x += __shfl_xor_sync(mask, x, o);       // butterfly: every lane gets the same bits
        return x;
    }
    __device__ void count(int *c) const { atomicAdd(c, 1); }
    __device__ void sync_warp() const { __syncwarp(); }
    __device__ void sync_block() const { __syncthreads(); }
};

constexpr int kSmallThreads = 512;

__global__ void __launch_bounds__(kSmallThreads, 1)
dmd_small_kernel(const double *__restrict__ r_stack, int K, double dt, double rank_tol, double *__restrict__ at_scratch,
                 double *__restrict__ alpha_re, double *__restrict__ alpha_im, double *__restrict__ sigma,
                 int *__restrict__ rank, int *__restrict__ info)
{
    extern __shared__ double small_smem[];
    const int n = K + 1;
    const size_t b = blockIdx.x;
    double *A = small_smem, *B = A + (size_t)n * K, *sig = B + (size_t)n * K, *eigw = sig + K;
    int *ord = (int *)(eigw + 3 * K), *counter = ord + K;
    GroupTeam<8> t;
    GroupTeam<32> te;
    const int st = ae_dmd_small::dmd_window(t, te, r_stack + b * n * n, K, dt, rank_tol, A, B, sig, eigw, ord, counter,
                                            at_scratch + b * n * n, alpha_re + b * K, alpha_im + b * K, sigma + b * K, rank + b);
    if (threadIdx.x == 0) info[b] = st;
}

static thread_local int t_last_sweeps = 0;           // Jacobi sweeps of the first window of this thread's last call (probe)
extern "C" int ae_dmd_last_jacobi_sweeps(void) { return t_last_sweeps; }

static bool small_fits(int K) { return K >= 1 && ae_dmd_small::dmd_window_smem_bytes(K) <= 227 * 1024; }

// `count` windows, r_stack [count][K+1][K+1] (device); host outputs [count][K] / [count]
static int dmd_small_batched(const double *r_stack, int count, int K, double dt, double rank_tol, double *alpha_re,
                             double *alpha_im, double *sigma_out, int32_t *rank, int32_t *info, cudaStream_t st)
{
    int status = AE_OK;
    const int n = K + 1;
    const size_t smem = ae_dmd_small::dmd_window_smem_bytes(K);
    double *dAt = nullptr, *dOut = nullptr;
    int *dInt = nullptr;
    const size_t per = (size_t)count * K;
    double *hOut = (double *)malloc(sizeof(double) * 3 * per);
    int *hInt = (int *)malloc(sizeof(int) * 2 * count);
    AE_CUDA_G(cudaFuncSetAttribute((const void *)dmd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AE_CUDA_G(cudaMallocAsync((void **)&dAt, sizeof(double) * (size_t)count * n * n, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dOut, sizeof(double) * 3 * per, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dInt, sizeof(int) * 2 * count, st));
    AE_CUDA_G(cudaMemsetAsync(dOut, 0, sizeof(double) * 3 * per, st));
    dmd_small_kernel<<<count, kSmallThreads, smem, st>>>(r_stack, K, dt, rank_tol, dAt, dOut, dOut + per, dOut + 2 * per, dInt, dInt + count);
    AE_CUDA_G(cudaGetLastError());
    AE_CUDA_G(cudaMemcpyAsync(hOut, dOut, sizeof(double) * 3 * per, cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaMemcpyAsync(hInt, dInt, sizeof(int) * 2 * count, cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaStreamSynchronize(st));
    memcpy(alpha_re, hOut, sizeof(double) * per);
    memcpy(alpha_im, hOut + per, sizeof(double) * per);
    if (sigma_out) memcpy(sigma_out, hOut + 2 * per, sizeof(double) * per);
    for (int m = 0; m < count; ++m) { rank[m] = hInt[m]; info[m] = hInt[count + m] & 0xff; }
    t_last_sweeps = hInt[count] >> 8;
done:
    cudaFreeAsync(dAt, st); cudaFreeAsync(dOut, st); cudaFreeAsync(dInt, st);
    free(hOut); free(hInt);
    return status;
}

static int small_info_to_status(int info, const char *who)
{
    if (info == ae_dmd_small::kOk) return AE_OK;
    if (info == ae_dmd_small::kNoRank) { set_error("%s: no singular value survives the rank rule", who); return AE_BAD_ARG; }
    set_error("%s: %s did not converge", who, info == ae_dmd_small::kSvdNotConverged ? "Jacobi SVD" : "QR eigenvalue iteration");
    return AE_CUDA_ERROR;
}

static int dmd_from_r_cusolver(const double *r_dev, int32_t K, double dt, double rank_tol, double *alpha_re, double *alpha_im,
                               double *sigma_out, int32_t *rank, cudaStream_t st);

// Windows that fit one block's shared memory (K <= 118) take the block kernel; AE_DMD_CUSOLVER=1 keeps them on the library (A/B).
static bool use_small_kernel(int K)
{
    const char *env = getenv("AE_DMD_CUSOLVER");              // read per call: the tests flip it
    return small_fits(K) && !(env && env[0] == '1');
}

extern "C" int ae_dmd_from_r(const double *r_dev, int32_t K, double dt, double rank_tol, double *alpha_re,
                             double *alpha_im, double *sigma_out, int32_t *rank, void *stream)
{
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    if (!r_dev || K < 1 || !alpha_re || !alpha_im || !rank) { set_error("ae_dmd_from_r: bad argument"); return AE_BAD_ARG; }
    if (!use_small_kernel(K)) return dmd_from_r_cusolver(r_dev, K, dt, rank_tol, alpha_re, alpha_im, sigma_out, rank, st);
    int32_t info = 0;
    AE_TRY(dmd_small_batched(r_dev, 1, K, dt, rank_tol, alpha_re, alpha_im, sigma_out, rank, &info, st));
    return small_info_to_status(info, "ae_dmd_from_r");
}

extern "C" int ae_dmd_from_r_batched(const double *r_stack, int32_t count, int32_t K, double dt, double rank_tol,
                                     double *alpha_re, double *alpha_im, double *sigma_out, int32_t *rank, int32_t *status,
                                     void *stream)
{
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    if (!r_stack || count < 1 || K < 1 || !alpha_re || !alpha_im || !rank || !status) { set_error("ae_dmd_from_r_batched: bad argument"); return AE_BAD_ARG; }
    if (use_small_kernel(K)) {
        AE_TRY(dmd_small_batched(r_stack, count, K, dt, rank_tol, alpha_re, alpha_im, sigma_out, rank, status, st));
        for (int m = 0; m < count; ++m) status[m] = small_info_to_status(status[m], "ae_dmd_from_r_batched");
        return AE_OK;
    }
    const size_t n = (size_t)K + 1;
    for (int m = 0; m < count; ++m)
        status[m] = dmd_from_r_cusolver(r_stack + m * n * n, K, dt, rank_tol, alpha_re + (size_t)m * K, alpha_im + (size_t)m * K,
                                        sigma_out ? sigma_out + (size_t)m * K : nullptr, rank + m, st);
    return AE_OK;
}

static int dmd_from_r_cusolver(const double *r_dev, int32_t K, double dt, double rank_tol, double *alpha_re,
                               double *alpha_im, double *sigma_out, int32_t *rank, cudaStream_t st)
{
    int status = AE_OK;
    const int n = K + 1;
    cusolverDnHandle_t h = nullptr;
    cusolverDnParams_t params = nullptr;
    double *dAx = nullptr, *dRy = nullptr, *dS = nullptr, *dU = nullptr, *dVT = nullptr, *dwork = nullptr, *dT = nullptr,
           *dAt = nullptr, *dInv = nullptr, *dEig = nullptr;
    void *dbuf = nullptr, *hbuf = nullptr;
    int *dinfo = nullptr;
    int lwork = 0, info = 0, r = 0;
    size_t dbytes = 0, hbytes = 0;
    double *sig = (double *)malloc(sizeof(double) * K), *inv = (double *)malloc(sizeof(double) * K),
           *W = (double *)malloc(sizeof(double) * 2 * K);

    if ((status = solver_ctx(st, &h, &params)) != AE_OK) goto done;
    AE_CUDA_G(cudaMallocAsync((void **)&dAx, sizeof(double) * n * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dRy, sizeof(double) * n * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dS, sizeof(double) * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dU, sizeof(double) * n * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dVT, sizeof(double) * K * K, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dinfo, sizeof(int), st));
    split_r_kernel<<<32, 256, 0, st>>>(r_dev, n, dAx, dRy);
    AE_CUDA_G(cudaGetLastError());
    // thin SVD of the (K+1) x K triangular factor: same singular values and V as X (time_unit.py:72)
    AE_CUSOLVER(cusolverDnDgesvd_bufferSize(h, n, K, &lwork));
    AE_CUDA_G(cudaMallocAsync((void **)&dwork, sizeof(double) * (size_t)lwork, st));
    AE_CUSOLVER(cusolverDnDgesvd(h, 'S', 'S', n, K, dAx, n, dS, dU, n, dVT, K, dwork, lwork, nullptr, dinfo));
    AE_CUDA_G(cudaMemcpyAsync(sig, dS, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaStreamSynchronize(st));
    if (info != 0) { set_error("gesvd info=%d", info); status = AE_CUDA_ERROR; goto done; }
    {
        double total = 0.0, run = 0.0;                  // rank rule, time_unit.py:75-76
        for (int j = 0; j < K; ++j) total += sig[j];
        for (int j = 0; j < K; ++j) {
            run += sig[j];
            if (1.0 - run / total > rank_tol && sig[j] > 0.0) r = j + 1; else break;
        }
        if (r < 1) { set_error("ae_dmd_from_r: no singular value survives the rank rule"); status = AE_BAD_ARG; goto done; }
        for (int j = 0; j < r; ++j) inv[j] = 1.0 / sig[j];
        if (sigma_out) memcpy(sigma_out, sig, sizeof(double) * K);
    }
    AE_CUDA_G(cudaMallocAsync((void **)&dT, sizeof(double) * n * r, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dAt, sizeof(double) * r * r, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dInv, sizeof(double) * r, st));
    AE_CUDA_G(cudaMallocAsync((void **)&dEig, sizeof(double) * 2 * r, st));
    AE_CUDA_G(cudaMemcpyAsync(dInv, inv, sizeof(double) * r, cudaMemcpyHostToDevice, st));
    atilde_from_svd_kernel<<<1, 1024, 0, st>>>(dU, n, dVT, K, dRy, n, dInv, r, dT, dAt);
    AE_CUDA_G(cudaGetLastError());
    AE_CUSOLVER(cusolverDnXgeev_bufferSize(h, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, r, CUDA_R_64F,
                                           dAt, r, CUDA_C_64F, dEig, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1,
                                           CUDA_R_64F, &dbytes, &hbytes));
    AE_CUDA_G(cudaMallocAsync(&dbuf, dbytes ? dbytes : 8, st));
    hbuf = malloc(hbytes ? hbytes : 8);
    AE_CUSOLVER(cusolverDnXgeev(h, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, r, CUDA_R_64F, dAt, r,
                                CUDA_C_64F, dEig, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dbuf,
                                dbytes, hbuf, hbytes, dinfo));
    AE_CUDA_G(cudaMemcpyAsync(W, dEig, sizeof(double) * 2 * r, cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    AE_CUDA_G(cudaStreamSynchronize(st));
    if (info != 0) { set_error("geev info=%d", info); status = AE_CUDA_ERROR; goto done; }
    for (int j = 0; j < r; ++j) {                       // alpha = (1 - 1/lambda)/dt  (time_unit.py:91)
        const double re = W[2 * j], im = W[2 * j + 1], d = re * re + im * im;
        alpha_re[j] = (1.0 - re / d) / dt;
        alpha_im[j] = (im / d) / dt;
    }
    *rank = r;
done:
    cudaFreeAsync(dAx, st); cudaFreeAsync(dRy, st); cudaFreeAsync(dS, st); cudaFreeAsync(dU, st); cudaFreeAsync(dVT, st); cudaFreeAsync(dwork, st); cudaFreeAsync(dT, st);
    cudaFreeAsync(dAt, st); cudaFreeAsync(dInv, st); cudaFreeAsync(dEig, st); cudaFreeAsync(dbuf, st); cudaFreeAsync(dinfo, st);
    free(hbuf); free(sig); free(inv); free(W);
    return status;
}
