// Host instantiation of alpha_eigen_b200/csrc/ae_dmd_small.cuh with a one-thread team: the same SVD / rank / Atilde /
// eigenvalue code the GPU block runs, callable from the CPU tests (tests/test_dmd_small_cpu.py builds it with g++).
#include <stdlib.h>
#include "../alpha_eigen_b200/csrc/ae_dmd_small.cuh"

struct OneThread {
    int lane() const { return 0; }
    int lanes() const { return 1; }
    int warp() const { return 0; }
    int warps() const { return 1; }
    double sum(double x) const { return x; }
    void count(int *c) const { *c += 1; }
    void sync_warp() const {}
    void sync_block() const {}
};

extern "C" int dmd_small_host(const double *R, int K, double dt, double rank_tol, double *alpha_re, double *alpha_im,
                              double *sigma, int *rank)
{
    const int n = K + 1;
    double *A = (double *)malloc(sizeof(double) * n * K), *B = (double *)malloc(sizeof(double) * n * K);
    double *sig = (double *)malloc(sizeof(double) * K), *eigw = (double *)malloc(sizeof(double) * 3 * K);
    double *at = (double *)malloc(sizeof(double) * n * n);
    int *ord = (int *)malloc(sizeof(int) * K), counter = 0;
    OneThread t;
    const int st = ae_dmd_small::dmd_window(t, t, R, K, dt, rank_tol, A, B, sig, eigw, ord, &counter, at, alpha_re, alpha_im, sigma, rank);
    free(A); free(B); free(sig); free(eigw); free(at); free(ord);
    return st & 0xff;
}

// eigenvalues of a general real matrix (row-major n x n) by the same balance / hessenberg / hqr chain
extern "C" int eig_small_host(const double *a_in, int n, double *wr, double *wi)
{
    const int ld = n | 1;
    double *a = (double *)calloc((size_t)n * ld, sizeof(double)), *v = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i * ld + j] = a_in[i * n + j];
    OneThread t;
    ae_dmd_small::balance(t, a, n, ld);
    ae_dmd_small::hessenberg(t, a, n, ld, v);
    const int st = ae_dmd_small::hqr(t, a, n, ld, wr, wi);
    free(a); free(v);
    return st;
}
