// Fused source iteration for small slabs (sm_100a, cooperative launch).
//
// The Kornreich-Parsons problems of the reference (450-1800 cells, 4-196 angles, one group) and the members
// of a parameter ensemble are far too small to fill a B200: run through the streaming kernels they cost two
// launches and a host poll per source iteration, ~20 us for a few microseconds of arithmetic.  Here the WHOLE
// loop of OneGroup.source_iteration_for_scalar_flux (group.py:62-81) / TimeUnit.time_dependent_source_iteration
// (time_unit.py:44-59) runs in one kernel: one CTA per row block (8 or 16 angles of one group and direction),
// every warp sweeping its angle tile by tile with the same tile arithmetic as the streaming kernel
// (ae_tile.cuh), a grid-wide barrier, the flux update (assembly, norm, scatter source) spread over all
// threads, another grid barrier, the convergence decision taken redundantly by every CTA from the same
// per-block sums in a fixed order.  The host reads the control block once, at the end.
// Inputs stay in L2 (a few MB), so plain loads suffice; results are bit-compatible with the streaming path
// (same maps, same angle order in the flux sum), which the parity tests exercise through both.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "ae_common.cuh"
#include "ae_tile.cuh"

namespace cg = cooperative_groups;

namespace ae {

struct SmallArgs {
    const double *hs, *cdt, *mu_ihx, *wgt, *psi_in, *base, *scat;
    const int32_t *mat;
    double *q0, *q1, *phi0, *phi1, *partial, *scratch;
    ae_ctrl *ctrl;
    int64_t Z, Zp;
    int A, n_neg, G, nb_neg, nbg;
    int max_iterations;
    double cold, tol;
};

// sum of the per-block (num, den) pairs in block order; every CTA computes the same value
__device__ __forceinline__ void all_blocks_sum2(const double *scratch, int nblocks, double &a, double &b, double *sm)
{
    a = 0.0; b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        a += __ldcg(scratch + 2 * i);
        b += __ldcg(scratch + 2 * i + 1);
    }
    block_sum2(a, b, sm);
    __syncthreads();
    if (threadIdx.x == 0) { sm[0] = a; sm[1] = b; }
    __syncthreads();
    a = sm[0]; b = sm[1];
    __syncthreads();
}

template <int AC, bool READ_PSI>
__global__ void __launch_bounds__(AC * 32) small_source_iteration_kernel(const SmallArgs p)
{
    extern __shared__ double red[];                     // [2][AC][kTile]
    __shared__ double sm[64];
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nt = p.Zp / kTile;
    const int64_t n = (int64_t)p.G * p.Zp;

    // this CTA's row block
    const int rb = blockIdx.x;
    const int g = rb / p.nbg, blk = rb % p.nbg;
    const bool neg = blk < p.nb_neg;
    const int a0 = neg ? blk * AC : p.n_neg + (blk - p.nb_neg) * AC;
    const int nblk = min(AC, (neg ? p.n_neg : p.A) - a0);
    const int a = a0 + (warp < nblk ? warp : 0);
    const double m = fabs(p.mu_ihx[a]), two_m = 2.0 * m;
    const double hw = warp < nblk ? 0.5 * p.wgt[a] : 0.0;
    const int64_t goff = (int64_t)g * p.Zp + lane * 2;
    const int64_t prow = ((int64_t)g * p.A + a) * p.Zp + lane * 2;
    double *part = p.partial + ((int64_t)blk * p.G + g) * p.Zp;

    double *q[2] = {p.q0, p.q1};
    double *phi[2] = {p.phi0, p.phi1};
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + tid;

    // flux update over the cells owned by this thread; it == 0 is the cold start (group.py:66 / time_unit.py:45)
    auto flux_update = [&](int it, double &num, double &den) {
        const double *phi_in = phi[(it + 1) & 1];
        double *phi_out = phi[it & 1], *q_next = q[it & 1];
        num = 0.0; den = 0.0;
        for (int64_t s = gtid; s < p.Zp; s += gthreads) {
            const bool valid = logical_cell(s) < p.Z;
            for (int gg = 0; gg < p.G; ++gg) {
                const int64_t i = (int64_t)gg * p.Zp + s;
                double x;
                if (it == 0) {
                    x = valid ? p.cold : 0.0;
                } else {
                    x = p.partial[i];
                    for (int pp = 1; pp < p.nbg; ++pp) x += p.partial[(int64_t)pp * n + i];    // block order (group.py:77)
                    const double d = x - phi_in[i];                                         // group.py:78
                    num = fma(d, d, num);
                    den = fma(x, x, den);
                }
                phi_out[i] = x;                                                             // group.py:80
            }
            const double *sc = p.scat + (size_t)p.mat[s] * p.G * p.G;
            for (int gg = 0; gg < p.G; ++gg) {
                double acc = 0.0;
                for (int gp = 0; gp < p.G; ++gp) acc = fma(phi_out[(int64_t)gp * p.Zp + s], __ldg(sc + gp * p.G + gg), acc);
                const int64_t i = (int64_t)gg * p.Zp + s;
                q_next[i] = p.base[i] + 0.5 * acc;                                          // group.py:73 / time_unit.py:53
            }
        }
    };

    double num, den;
    flux_update(0, num, den);
    grid.sync();

    int it = 0, converged = 0;
    double change = 0.0;
    while (it < p.max_iterations - 1) {                 // `assert iteration < max_iterations`, group.py:72
        ++it;
        const double *qc = q[(it - 1) & 1] + goff;
        // ---- sweep phase: this CTA's row block over the whole slab ------------------------------------
        double carry = 0.0;                             // vacuum boundary: group.py:48,54
        for (int64_t ti = 0; ti < nt; ++ti) {
            const int64_t off = (neg ? nt - 1 - ti : ti) * kTile;
            double ca[kL], cb[kL], f[kL];
            tile_maps<READ_PSI>(p.hs + goff + off, qc + off, READ_PSI ? p.cdt + goff + off : nullptr,
                                READ_PSI ? p.psi_in + prow + off : nullptr, m, two_m, ca, cb);
            if (neg) tile_solve<true>(ca, cb, carry, f, lane);
            else tile_solve<false>(ca, cb, carry, f, lane);
            if (off + kTile > p.Z) {
#pragma unroll
                for (int jj = 0; jj < kL; ++jj)
                    if (off + lane * kL + jj >= p.Z) f[jj] = 0.0;
            }
            double *rw = red + (ti & 1) * (AC * kTile) + warp * kTile + lane * 2;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                *reinterpret_cast<double2 *>(rw + k * 64) = make_double2(hw * f[2 * k], hw * f[2 * k + 1]);
            __syncthreads();
            if (tid < kTile) {
                const double *rr = red + (ti & 1) * (AC * kTile) + tid;
                double sum = rr[0];
#pragma unroll
                for (int w = 1; w < AC; ++w) sum += rr[w * kTile];      // warps = angles in order (group.py:77)
                part[off + tid] = sum;
            }
        }
        grid.sync();
        // ---- flux update + norm --------------------------------------------------------------------------
        flux_update(it, num, den);
        block_sum2(num, den, sm);
        if (tid == 0) { p.scratch[2 * blockIdx.x] = num; p.scratch[2 * blockIdx.x + 1] = den; }
        grid.sync();
        double sa, sb;
        all_blocks_sum2(p.scratch, gridDim.x, sa, sb, sm);
        change = sqrt(sa) / sqrt(sb);                   // group.py:78
        if (change < p.tol) { converged = 1; break; }   // group.py:79
    }
    if (blockIdx.x == 0 && tid == 0) {
        p.ctrl->change = change;
        p.ctrl->inner_iters = it;
        p.ctrl->inner_done = converged;
        p.ctrl->ticket = 0;
    }
}

// Can the fused kernel run this problem?  One CTA per row block, all co-resident; small enough to be latency bound.
template <int AC, bool RD>
static int try_launch(const SmallArgs &p, int rows, cudaStream_t st, bool &launched)
{
    const size_t smem = 2 * AC * kTile * sizeof(double);
    static int capacity = 0;
    if (!capacity) {
        if (smem > 48 * 1024)
            AE_CUDA(cudaFuncSetAttribute((const void *)small_source_iteration_kernel<AC, RD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, dev = 0, sms = 0;
        AE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_source_iteration_kernel<AC, RD>, AC * 32, smem));
        AE_CUDA(cudaGetDevice(&dev));
        AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        capacity = per_sm * sms > 0 ? per_sm * sms : -1;
    }
    launched = false;
    if (capacity < rows) return AE_OK;
    void *args[] = {(void *)&p};
    AE_CUDA(cudaLaunchCooperativeKernel((const void *)small_source_iteration_kernel<AC, RD>, dim3(rows), dim3(AC * 32), args,
                                        smem, st));
    launched = true;
    return AE_OK;
}

int block_angles(int A, int n_neg);

// Returns AE_OK with *launched = 1 when the fused kernel was queued for this source iteration.
int launch_small_source_iteration(const ae_slab *s, ae_work *w, const double *psi_in, double cold, double tol,
                                  int max_iterations, int *launched, cudaStream_t st)
{
    *launched = 0;
    static const char *off = getenv("AE_NO_FUSED_SMALL");
    if (off && off[0] == '1') return AE_OK;
    if (w->comm) return AE_OK;                                          // multi-GPU goes through the exchange path
    const int64_t updates = s->Zp * (int64_t)s->A * s->G;
    if (updates > (1 << 23) || s->Zp / kTile > 256) return AE_OK;       // big enough for the streaming kernels
    const int ac = block_angles(s->A, s->n_neg);
    SmallArgs p;
    p.hs = s->hs; p.cdt = s->cdt; p.mu_ihx = s->mu_ihx; p.wgt = s->wgt; p.psi_in = psi_in; p.base = w->base;
    p.scat = s->scat; p.mat = s->mat; p.q0 = w->q0; p.q1 = w->q1; p.phi0 = w->phi; p.phi1 = w->phi_alt;
    p.partial = w->partial; p.scratch = w->norm_scratch; p.ctrl = w->ctrl; p.Z = s->Z; p.Zp = s->Zp; p.A = s->A;
    p.n_neg = s->n_neg; p.G = s->G;
    p.nb_neg = (s->n_neg + ac - 1) / ac;
    p.nbg = p.nb_neg + (s->A - s->n_neg + ac - 1) / ac;
    p.max_iterations = max_iterations; p.cold = cold; p.tol = tol;
    const int rows = s->G * p.nbg;
    if (2 * rows > 4 * (int)ae_norm_blocks(s->Zp) + 2 * s->G) return AE_OK;   // per-block sums must fit norm_scratch
    bool ok = false;
    if (ac == 16) {
        if (psi_in) AE_TRY((try_launch<16, true>(p, rows, st, ok)));
        else AE_TRY((try_launch<16, false>(p, rows, st, ok)));
    } else {
        if (psi_in) AE_TRY((try_launch<8, true>(p, rows, st, ok)));
        else AE_TRY((try_launch<8, false>(p, rows, st, ok)));
    }
    *launched = ok ? 1 : 0;
    return AE_OK;
}

}  // namespace ae
