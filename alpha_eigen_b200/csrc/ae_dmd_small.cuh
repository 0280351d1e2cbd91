// The small dense part of a DMD (time_unit.py:72-91) for one snapshot window, written for ONE thread block:
//   thin SVD of A = R[:, 0:K] by one-sided Jacobi rotations (Hestenes), the rank rule, Atilde = U_r^T Y V_r S_r^-1
//   (the same rotations carried along on Y = R[:, 1:K+1] give Y V directly), and the eigenvalues of Atilde by
//   balancing, Householder reduction to Hessenberg form and the Francis double-shift QR iteration (EISPACK hqr).
// cuSOLVER's gesvd + Xgeev do the same in ~10^4 tiny launches (13 ms per window, host bound); here a window is one
// block of one launch, so an ensemble's windows run side by side.
//
// The algorithms are templates over a "team" that says how many warps / lanes execute them and how they synchronise:
// the CUDA team (a thread block) is defined in ae_dmd.cu; tests/dmd_small_host.cpp instantiates the SAME code with a
// one-thread team on the host and checks it against LAPACK there (no GPU needed).  Every lane executes the scalar
// control flow redundantly from the same shared values, only the vector loops are split.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define AE_SD __device__
#define AE_SHD __host__ __device__
#else
#define AE_SD
#define AE_SHD
#endif

namespace ae_dmd_small {

enum { kOk = 0, kSvdNotConverged = 1, kEigNotConverged = 2, kNoRank = 3 };
constexpr int kMaxSweeps = 60;

#if defined(__CUDA_ARCH__)
#define AE_RSQRT(x) rsqrt(x)
#else
#define AE_RSQRT(x) (1.0 / sqrt(x))
#endif

// ---- Householder QR with column pivoting of A (n x K, n > K, column-major, ld = lda), the reflectors and the column
// swaps applied to B (same shape) as well: on return the upper triangle of A's top K rows is R1 (A P = Q1 R1, diagonal
// non-increasing in magnitude) and B is H_K .. H_1 B P.  This is the preconditioner of Drmac & Veselic for the Jacobi SVD
// below: on a triangular factor of a window whose singular values span 16 decades the plain method needs 25 sweeps, on
// R1^T eight.  cn: K doubles, v: n doubles of scratch; returns |A|_F^2 through *fro2.
template <class Team>
AE_SD void qr_column_pivoting(Team &t, double *A, double *B, int n, int K, int lda, double *cn, double *v, double *fro2)
{
    const int tid = t.warp() * t.lanes() + t.lane(), nthreads = t.warps() * t.lanes();
    for (int k = 0; k < K; ++k) {
        for (int j = k + t.warp(); j < K; j += t.warps()) {            // remaining column norms (recomputed: no downdating error)
            const double *a = A + (size_t)j * lda;
            double s = 0.0;
            for (int i = k + t.lane(); i < n; i += t.lanes()) s = fma(a[i], a[i], s);
            s = t.sum(s);
            if (t.lane() == 0) cn[j] = s;
        }
        t.sync_block();
        int pv = k;
        double best = cn[k], total = 0.0;
        for (int j = k; j < K; ++j) {                                  // every thread: same values, same answer
            total += cn[j];
            if (cn[j] > best) { best = cn[j]; pv = j; }
        }
        if (k == 0 && tid == 0) *fro2 = total;
        if (pv != k)
            for (int i = tid; i < n; i += nthreads) {
                double x = A[(size_t)k * lda + i]; A[(size_t)k * lda + i] = A[(size_t)pv * lda + i]; A[(size_t)pv * lda + i] = x;
                x = B[(size_t)k * lda + i]; B[(size_t)k * lda + i] = B[(size_t)pv * lda + i]; B[(size_t)pv * lda + i] = x;
            }
        t.sync_block();
        // reflector of column k, rows k .. n-1:  v = x - alpha e_1,  H = I - beta v v^T
        const double *x = A + (size_t)k * lda;
        const double x0 = x[k], norm = sqrt(best);
        if (norm == 0.0) { t.sync_block(); continue; }                 // nothing left: the remaining columns are zero
        const double alpha = -copysign(norm, x0);
        const double beta = 1.0 / (best - x0 * alpha);                 // 2 / v^T v
        for (int i = k + tid; i < n; i += nthreads) v[i] = i == k ? x0 - alpha : x[i];
        t.sync_block();
        for (int c = t.warp(); c < 2 * K - k - 1; c += t.warps()) {    // trailing columns of A, all columns of B
            double *col = c < K - k - 1 ? A + (size_t)(k + 1 + c) * lda : B + (size_t)(c - (K - k - 1)) * lda;
            double s = 0.0;
            for (int i = k + t.lane(); i < n; i += t.lanes()) s = fma(v[i], col[i], s);
            s = t.sum(s) * beta;
            for (int i = k + t.lane(); i < n; i += t.lanes()) col[i] = fma(-s, v[i], col[i]);
        }
        for (int i = k + tid; i < n; i += nthreads) A[(size_t)k * lda + i] = i == k ? alpha : 0.0;
        t.sync_block();
    }
}

// In-place transpose of the top K x K block of a column-major n x K array (ld = lda); with `lower_only` the strict upper
// triangle of the result is zeroed (the transposed triangular factor); rows K .. n-1 are zeroed.
template <class Team>
AE_SD void transpose_top(Team &t, double *A, int n, int K, int lda, bool lower_only)
{
    const int tid = t.warp() * t.lanes() + t.lane(), nthreads = t.warps() * t.lanes();
    for (int idx = tid; idx < K * K; idx += nthreads) {
        const int i = idx % K, j = idx / K;                            // element (i, j), i < j: swap with (j, i)
        if (i < j) {
            const double up = A[(size_t)j * lda + i], lo = A[(size_t)i * lda + j];
            A[(size_t)i * lda + j] = up;
            A[(size_t)j * lda + i] = lower_only ? 0.0 : lo;
        }
    }
    for (int idx = tid; idx < (n - K) * K; idx += nthreads) A[(size_t)(idx / (n - K)) * lda + K + idx % (n - K)] = 0.0;
    t.sync_block();
}

// ---- one-sided Jacobi: rotate the columns of A (n x K, column-major, ld = lda) until they are mutually orthogonal; B
// (same shape) receives the same rotations.  On return column j of A is sigma_j u_j (unsorted).  Pairs of columns that
// are both below `floor2` (squared norm: rounding noise of the factorisation that produced A) are left alone.
// `counter` is one shared int.
template <class Team>
AE_SD int jacobi_columns(Team &t, double *A, double *B, int n, int K, int lda, double floor2, int *counter, int *sweeps_out)
{
    const int Kp = K + (K & 1);                        // round-robin over an even number of players
    const int rounds = Kp - 1, pairs = Kp / 2;
    const double tol = 1.5e-16 * sqrt((double)n);      // LAPACK dgesvj's threshold: sqrt(m) eps
    const double tol2 = tol * tol;
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        if (t.warp() == 0 && t.lane() == 0) *counter = 0;
        t.sync_block();
        for (int rd = 0; rd < rounds; ++rd) {
            for (int i = t.warp(); i < pairs; i += t.warps()) {
                int p, q;
                if (i == 0) { p = Kp - 1; q = rd; }
                else { p = (rd + i) % rounds; q = (rd + rounds - i) % rounds; }
                if (p >= K || q >= K) continue;        // the dummy player of an odd K
                if (p > q) { const int s = p; p = q; q = s; }
                double *ap = A + (size_t)p * lda, *aq = A + (size_t)q * lda;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int k = t.lane(); k < n; k += t.lanes()) {
                    const double x = ap[k], y = aq[k];
                    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                }
                al = t.sum(al); be = t.sum(be); ga = t.sum(ga);
                if (al == 0.0 || be == 0.0 || (al <= floor2 && be <= floor2) || !(ga * ga > tol2 * al * be)) continue;
                // tan of the rotation angle: the smaller root of t^2 + 2 zeta t - 1 = 0, zeta = (be - al) / (2 ga)
                const double d = be - al;
                const double num = (d < 0.0) != (ga < 0.0) ? -2.0 * fabs(ga) : 2.0 * fabs(ga);      // sign(zeta) 2|ga|
                const double tn = num / (fabs(d) + sqrt(fma(d, d, 4.0 * ga * ga)));
                const double c = AE_RSQRT(fma(tn, tn, 1.0)), s = c * tn;
                double *bp = B + (size_t)p * lda, *bq = B + (size_t)q * lda;
                for (int k = t.lane(); k < n; k += t.lanes()) {
                    const double x = ap[k], y = aq[k];
                    ap[k] = c * x - s * y; aq[k] = s * x + c * y;
                    const double u = bp[k], v = bq[k];
                    bp[k] = c * u - s * v; bq[k] = s * u + c * v;
                }
                if (t.lane() == 0) t.count(counter);
            }
            t.sync_block();
        }
        const int rotated = *counter;
        t.sync_block();
        if (rotated == 0) { *sweeps_out = sweep + 1; return kOk; }
    }
    *sweeps_out = kMaxSweeps;
    return kSvdNotConverged;
}

// ---- balancing (Parlett & Reinsch; powers of two only, so no rounding): a is n x n row-major with leading dimension ld.
// One warp.
template <class Team>
AE_SD void balance(Team &t, double *a, int n, int ld)
{
    const double radix = 2.0, sqrdx = 4.0;
    for (int pass = 0; pass < 64; ++pass) {
        bool done = true;
        for (int i = 0; i < n; ++i) {
            double c = 0.0, r = 0.0;
            for (int j = t.lane(); j < n; j += t.lanes())
                if (j != i) { c += fabs(a[j * ld + i]); r += fabs(a[i * ld + j]); }
            c = t.sum(c); r = t.sum(r);
            if (c != 0.0 && r != 0.0 && isfinite(c) && isfinite(r)) {
                double g = r / radix, f = 1.0;
                const double s = c + r;
                while (c < g) { f *= radix; c *= sqrdx; }
                g = r * radix;
                while (c > g) { f /= radix; c /= sqrdx; }
                if ((c + r) / f < 0.95 * s) {
                    done = false;
                    g = 1.0 / f;
                    for (int j = t.lane(); j < n; j += t.lanes()) a[i * ld + j] *= g;
                    t.sync_warp();
                    for (int j = t.lane(); j < n; j += t.lanes()) a[j * ld + i] *= f;
                }
            }
            t.sync_warp();
        }
        if (done) break;
    }
}

// ---- Householder reduction to upper Hessenberg form (similarity), `v` is n doubles of scratch.  One warp.
template <class Team>
AE_SD void hessenberg(Team &t, double *a, int n, int ld, double *v)
{
    for (int m = 1; m < n - 1; ++m) {
        double scale = 0.0;
        for (int i = m + t.lane(); i < n; i += t.lanes()) scale += fabs(a[i * ld + m - 1]);
        scale = t.sum(scale);
        if (scale == 0.0) continue;
        double h = 0.0;
        for (int i = m + t.lane(); i < n; i += t.lanes()) {
            const double x = a[i * ld + m - 1] / scale;
            v[i] = x;
            h = fma(x, x, h);
        }
        h = t.sum(h);
        t.sync_warp();
        const double x0 = v[m];
        const double g = -copysign(sqrt(h), x0);
        const double beta = 1.0 / (h - x0 * g);        // 2 / v^T v with v_0 = x_0 - g
        t.sync_warp();
        if (t.lane() == 0) v[m] = x0 - g;
        t.sync_warp();
        for (int j = m + t.lane(); j < n; j += t.lanes()) {            // H a: columns m .. n-1 (column m-1 is known)
            double s = 0.0;
            for (int i = m; i < n; ++i) s = fma(v[i], a[i * ld + j], s);
            s *= beta;
            for (int i = m; i < n; ++i) a[i * ld + j] = fma(-s, v[i], a[i * ld + j]);
        }
        t.sync_warp();
        for (int i = t.lane(); i < n; i += t.lanes()) {                // (H a) H: all rows
            double s = 0.0;
            for (int j = m; j < n; ++j) s = fma(a[i * ld + j], v[j], s);
            s *= beta;
            for (int j = m; j < n; ++j) a[i * ld + j] = fma(-s, v[j], a[i * ld + j]);
        }
        t.sync_warp();
        for (int i = m + t.lane(); i < n; i += t.lanes()) a[i * ld + m - 1] = i == m ? scale * g : 0.0;
        t.sync_warp();
    }
}

// ---- eigenvalues of an upper Hessenberg matrix (destroyed): Francis double-shift QR as in EISPACK hqr.  One warp.
template <class Team>
AE_SD int hqr(Team &t, double *a, int n, int ld, double *wr, double *wi)
{
#define AE_A(i, j) a[(i) * ld + (j)]
    double anorm = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = (i > 0 ? i - 1 : 0) + t.lane(); j < n; j += t.lanes()) anorm += fabs(AE_A(i, j));
    anorm = t.sum(anorm);
    int nn = n - 1;
    double tshift = 0.0;
    double p = 0.0, q = 0.0, r = 0.0;
    while (nn >= 0) {
        int its = 0;
        for (;;) {
            int l;
            for (l = nn; l >= 1; --l) {
                double s = fabs(AE_A(l - 1, l - 1)) + fabs(AE_A(l, l));
                if (s == 0.0) s = anorm;
                if (fabs(AE_A(l, l - 1)) + s == s) {
                    t.sync_warp();
                    if (t.lane() == 0) AE_A(l, l - 1) = 0.0;
                    t.sync_warp();
                    break;
                }
            }
            double x = AE_A(nn, nn);
            if (l == nn) {                              // one root
                if (t.lane() == 0) { wr[nn] = x + tshift; wi[nn] = 0.0; }
                nn -= 1;
                break;
            }
            double y = AE_A(nn - 1, nn - 1), w = AE_A(nn, nn - 1) * AE_A(nn - 1, nn);
            if (l == nn - 1) {                          // two roots
                p = 0.5 * (y - x);
                q = p * p + w;
                double z = sqrt(fabs(q));
                x += tshift;
                if (t.lane() == 0) {
                    if (q >= 0.0) {
                        z = p + copysign(z, p);
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = z; wi[nn] = -z;
                    }
                }
                nn -= 2;
                break;
            }
            if (its == 60) return kEigNotConverged;
            if (its == 10 || its == 20 || its == 40) {  // exceptional shift
                tshift += x;
                t.sync_warp();
                for (int i = t.lane(); i <= nn; i += t.lanes()) AE_A(i, i) -= x;
                t.sync_warp();
                const double s = fabs(AE_A(nn, nn - 1)) + fabs(AE_A(nn - 1, nn - 2));
                y = x = 0.75 * s;
                w = -0.4375 * s * s;
            }
            ++its;
            int m;
            double z;
            for (m = nn - 2; m >= l; --m) {             // look for two consecutive small subdiagonal elements
                z = AE_A(m, m);
                r = x - z;
                double s = y - z;
                p = (r * s - w) / AE_A(m + 1, m) + AE_A(m, m + 1);
                q = AE_A(m + 1, m + 1) - z - r - s;
                r = AE_A(m + 2, m + 1);
                s = fabs(p) + fabs(q) + fabs(r);
                p /= s; q /= s; r /= s;
                if (m == l) break;
                const double u = fabs(AE_A(m, m - 1)) * (fabs(q) + fabs(r));
                const double v = fabs(p) * (fabs(AE_A(m - 1, m - 1)) + fabs(z) + fabs(AE_A(m + 1, m + 1)));
                if (u + v == v) break;
            }
            t.sync_warp();
            for (int i = m + 2 + t.lane(); i <= nn; i += t.lanes()) {
                AE_A(i, i - 2) = 0.0;
                if (i != m + 2) AE_A(i, i - 3) = 0.0;
            }
            t.sync_warp();
            for (int k = m; k <= nn - 1; ++k) {         // double QR step on rows l .. nn, columns m .. nn
                if (k != m) {
                    p = AE_A(k, k - 1);
                    q = AE_A(k + 1, k - 1);
                    r = k != nn - 1 ? AE_A(k + 2, k - 1) : 0.0;
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x != 0.0) { p /= x; q /= x; r /= x; }
                }
                const double s = copysign(sqrt(p * p + q * q + r * r), p);
                if (s == 0.0) continue;
                t.sync_warp();                          // everybody has read the column k-1 entries
                if (t.lane() == 0) {
                    if (k == m) { if (l != m) AE_A(k, k - 1) = -AE_A(k, k - 1); }
                    else AE_A(k, k - 1) = -s * x;
                }
                p += s;
                x = p / s; y = q / s; z = r / s;
                q /= p; r /= p;
                for (int j = k + t.lane(); j <= nn; j += t.lanes()) {          // row modification
                    double pp = AE_A(k, j) + q * AE_A(k + 1, j);
                    if (k != nn - 1) { pp += r * AE_A(k + 2, j); AE_A(k + 2, j) -= pp * z; }
                    AE_A(k + 1, j) -= pp * y;
                    AE_A(k, j) -= pp * x;
                }
                t.sync_warp();
                const int mmin = nn < k + 3 ? nn : k + 3;
                for (int i = l + t.lane(); i <= mmin; i += t.lanes()) {        // column modification
                    double pp = x * AE_A(i, k) + y * AE_A(i, k + 1);
                    if (k != nn - 1) { pp += z * AE_A(i, k + 2); AE_A(i, k + 2) -= pp * r; }
                    AE_A(i, k + 1) -= pp * q;
                    AE_A(i, k) -= pp;
                }
                t.sync_warp();
            }
        }
    }
    t.sync_warp();
    return kOk;
#undef AE_A
}

// ---- the rank rule of time_unit.py:75-76 on singular values sorted in descending order (one thread)
AE_SHD inline int dmd_rank(const double *sig, int K, double rank_tol)
{
    double total = 0.0, run = 0.0;
    for (int j = 0; j < K; ++j) total += sig[j];
    int r = 0;
    for (int j = 0; j < K; ++j) {
        run += sig[j];
        if (1.0 - run / total > rank_tol && sig[j] > 0.0) r = j + 1; else break;
    }
    return r;
}

// ---- the whole window.  `t` is the team of the SVD stage (its "warps" take column pairs: on the GPU groups of 8 lanes, so
// that a warp instruction carries four pairs), `te` the team of the eigenvalue stage (one full warp).  R: (K+1) x (K+1) row-major triangular factor (global).  Work areas (shared memory on the GPU):
//   A, B    (K+1) x K doubles each, column-major; after the SVD the space of A holds Atilde (r x (r|1))
//   sig     K doubles, eigw 3 K doubles, ord K ints, counter 1 int
//   at      (K+1)*(K+1) doubles of GLOBAL scratch (Atilde is assembled there while A and B are still needed)
// Results: alpha_re / alpha_im [r], sigma_out [K] (descending), *rank_out; return value: status | (Jacobi sweeps << 8).
template <class Team, class EigTeam>
AE_SD int dmd_window(Team &t, EigTeam &te, const double *R, int K, double dt, double rank_tol, double *A, double *B, double *sig,
                     double *eigw, int *ord, int *counter, double *at, double *alpha_re, double *alpha_im, double *sigma_out, int *rank_out)
{
    const int n = K + 1;
    const int tid = t.warp() * t.lanes() + t.lane(), nthreads = t.warps() * t.lanes();
    for (int idx = tid; idx < n * K; idx += nthreads) {
        const int i = idx % n, j = idx / n;
        A[idx] = R[i * n + j];
        B[idx] = R[i * n + j + 1];
    }
    t.sync_block();
    // A P = Q1 R1;  L = R1^T;  Jacobi: L W = Ut S  =>  A = (Q1 W) S (P Ut)^T.  With Y2 = Q1^T Y P carried as B = Y2^T:
    // Atilde = U_r^T Y V_r S_r^-1 = W_r^T Y2 Ut_r S_r^-1, i.e. Atilde[i][j] = (b'_i . a'_j) / sigma_j^2   (time_unit.py:85-89)
    double *fro2 = eigw + n;
    qr_column_pivoting(t, A, B, n, K, n, sig, eigw, fro2);
    transpose_top(t, A, n, K, n, true);
    transpose_top(t, B, n, K, n, false);
    const double floor2 = 4.84e-30 * *fro2;            // (10 eps |A|_F)^2: below the rounding noise of the TSQR factor
    int sweeps = 0;
    const int st = jacobi_columns(t, A, B, n, K, n, floor2, counter, &sweeps);
    if (st != kOk) return st;
    for (int j = t.warp(); j < K; j += t.warps()) {                    // sigma_j = |column j|
        double s = 0.0;
        for (int k = t.lane(); k < n; k += t.lanes()) s = fma(A[(size_t)j * n + k], A[(size_t)j * n + k], s);
        s = t.sum(s);
        if (t.lane() == 0) sig[j] = sqrt(s);
    }
    t.sync_block();
    if (tid == 0) {                                                    // descending order (insertion sort on indices)
        for (int j = 0; j < K; ++j) ord[j] = j;
        for (int j = 1; j < K; ++j) {
            const int o = ord[j];
            const double s = sig[o];
            int k = j - 1;
            while (k >= 0 && sig[ord[k]] < s) { ord[k + 1] = ord[k]; --k; }
            ord[k + 1] = o;
        }
        for (int j = 0; j < K; ++j) sigma_out[j] = sig[ord[j]];
        *rank_out = dmd_rank(sigma_out, K, rank_tol);
    }
    t.sync_block();
    const int r = *rank_out;
    if (r < 1) return kNoRank;
    const int ld = r | 1;                                              // odd: column walks hit distinct banks
    for (int idx = t.warp(); idx < r * r; idx += t.warps()) {
        const int i = idx / r, j = idx % r;
        const double *b = B + (size_t)ord[i] * n, *w = A + (size_t)ord[j] * n;
        double s = 0.0;
        for (int k = t.lane(); k < n; k += t.lanes()) s = fma(b[k], w[k], s);
        s = t.sum(s);
        if (t.lane() == 0) at[i * ld + j] = s / (sig[ord[j]] * sig[ord[j]]);
    }
    t.sync_block();
    double *a = A, *v = eigw, *wr = eigw + K, *wi = eigw + 2 * K;
    for (int idx = tid; idx < r * ld; idx += nthreads) a[idx] = (idx % ld) < r ? at[idx] : 0.0;
    t.sync_block();
    int est = kOk;
    if (te.warp() == 0) {
        balance(te, a, r, ld);
        hessenberg(te, a, r, ld, v);
        est = hqr(te, a, r, ld, wr, wi);
        if (est == kOk)
            for (int j = te.lane(); j < r; j += te.lanes()) {          // alpha = (1 - 1/lambda)/dt  (time_unit.py:91)
                const double re = wr[j], im = wi[j], d = re * re + im * im;
                alpha_re[j] = (1.0 - re / d) / dt;
                alpha_im[j] = (im / d) / dt;
            }
        if (te.lane() == 0) *counter = est;
    }
    t.sync_block();
    return *counter | (sweeps << 8);                                   // status, and the Jacobi sweep count above bit 8
}

// work-area sizes in doubles / bytes for a window of K columns
AE_SHD inline size_t dmd_window_smem_bytes(int K)
{
    return (size_t)2 * (K + 1) * K * sizeof(double) + (size_t)4 * K * sizeof(double) + (size_t)(K + 2) * sizeof(int) + 16;
}

}  // namespace ae_dmd_small
