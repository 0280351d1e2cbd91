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

// ---- one-sided Jacobi: rotate the columns of A (n x K, column-major, ld = lda) until they are mutually orthogonal; B
// (same shape) receives the same rotations.  On return column j of A is sigma_j u_j (unsorted).  `counter` is one shared int.
template <class Team>
AE_SD int jacobi_columns(Team &t, double *A, double *B, int n, int K, int lda, int *counter)
{
    const int Kp = K + (K & 1);                        // round-robin over an even number of players
    const int rounds = Kp - 1, pairs = Kp / 2;
    const double tol = 1.5e-16 * sqrt((double)n);      // LAPACK dgesvj's threshold: sqrt(m) eps
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
                if (al == 0.0 || be == 0.0 || !(fabs(ga) > tol * sqrt(al) * sqrt(be))) continue;
                const double zeta = (be - al) / (2.0 * ga);
                const double tn = fabs(zeta) > 1e150 ? 0.5 / zeta : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tn * tn), s = c * tn;
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
        if (rotated == 0) return kOk;
    }
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

// ---- the whole window.  R: (K+1) x (K+1) row-major triangular factor (global).  Work areas (shared memory on the GPU):
//   A, B    (K+1) x K doubles each, column-major; after the SVD the space of A holds Atilde (r x (r|1))
//   sig     K doubles, eigw 3 K doubles, ord K ints, counter 1 int
//   at      (K+1)*(K+1) doubles of GLOBAL scratch (Atilde is assembled there while A and B are still needed)
// Results: alpha_re / alpha_im [r], sigma_out [K] (descending), *rank_out; return value is the status.
template <class Team>
AE_SD int dmd_window(Team &t, const double *R, int K, double dt, double rank_tol, double *A, double *B, double *sig,
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
    const int st = jacobi_columns(t, A, B, n, K, n, counter);
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
    // Atilde[i][j] = (w_i . b_j) / (sigma_i sigma_j): U_r^T (Y V_r) S_r^-1 with u_i = w_i / sigma_i   (time_unit.py:85-89)
    const int ld = r | 1;                                              // odd: column walks hit distinct banks
    for (int idx = t.warp(); idx < r * r; idx += t.warps()) {
        const int i = idx / r, j = idx % r;
        const double *w = A + (size_t)ord[i] * n, *b = B + (size_t)ord[j] * n;
        double s = 0.0;
        for (int k = t.lane(); k < n; k += t.lanes()) s = fma(w[k], b[k], s);
        s = t.sum(s);
        if (t.lane() == 0) at[i * ld + j] = s / (sig[ord[i]] * sig[ord[j]]);
    }
    t.sync_block();
    double *a = A, *v = eigw, *wr = eigw + K, *wi = eigw + 2 * K;
    for (int idx = tid; idx < r * ld; idx += nthreads) a[idx] = (idx % ld) < r ? at[idx] : 0.0;
    t.sync_block();
    int est = kOk;
    if (t.warp() == 0) {
        balance(t, a, r, ld);
        hessenberg(t, a, r, ld, v);
        est = hqr(t, a, r, ld, wr, wi);
        if (est == kOk)
            for (int j = t.lane(); j < r; j += t.lanes()) {            // alpha = (1 - 1/lambda)/dt  (time_unit.py:91)
                const double re = wr[j], im = wi[j], d = re * re + im * im;
                alpha_re[j] = (1.0 - re / d) / dt;
                alpha_im[j] = (im / d) / dt;
            }
        if (t.lane() == 0) *counter = est;
    }
    t.sync_block();
    return *counter;
}

// work-area sizes in doubles / bytes for a window of K columns
AE_SHD inline size_t dmd_window_smem_bytes(int K)
{
    return (size_t)2 * (K + 1) * K * sizeof(double) + (size_t)4 * K * sizeof(double) + (size_t)(K + 2) * sizeof(int) + 16;
}

}  // namespace ae_dmd_small
