// Source-iteration kernels: scalar-flux assembly, convergence norm, scatter and fission sources.
//
// Replaces the NumPy statements around the sweeps in the reference:
//   group.py:73      source = self.source + phi_old*scatter_xs*0.5
//   group.py:77-80   phi accumulation, change = ||phi-phi_old||/||phi||, phi_old = phi
//   group.py:104-105 fission source 0.5*chi*nu_sigmaf*phi.old
//   group.py:108-115 k = ||nu_sigmaf*phi.new||/||nu_sigmaf*phi.old||, phi.old = phi.new/k
//   time_unit.py:37-41 per-step fission source and outer change
// generalised to G groups with the formulas of zone.py:38-51 (scatter matrix sigma_s(g'->g),
// fission chi_g * sum_g' nu_sigma_f_g' phi_g').  All per-cell arrays are [G][Zp] in the
// lane-blocked storage order; none of these kernels couples neighbouring cells, so the order
// is irrelevant here and every access is stride-1 across the threads of a warp.
#include "ae_common.cuh"

namespace ae {

constexpr int kCB = 128;        // cells (threads) per block in the per-cell kernels

__device__ __forceinline__ bool cell_valid(int64_t s, int64_t Z, int64_t Zp) { return s < Zp && logical_cell(s) < Z; }

// Deterministic grid-wide reduction of (num, den): per-block partials to scratch, last block sums them
// in a fixed order and calls `fin(num, den)` from thread 0.
template <class Fin>
__device__ __forceinline__ void grid_sum2(double num, double den, double *scratch, ae_ctrl *ctrl, double *sm, Fin fin)
{
    __shared__ bool is_last;
    block_sum2(num, den, sm);
    if (threadIdx.x == 0) {
        scratch[2 * blockIdx.x] = num;
        scratch[2 * blockIdx.x + 1] = den;
        __threadfence();
        const unsigned t = atomicAdd(&ctrl->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            a += __ldcg(scratch + 2 * i);
            b += __ldcg(scratch + 2 * i + 1);
        }
        block_sum2(a, b, sm);
        if (threadIdx.x == 0) {
            ctrl->ticket = 0;
            fin(a, b);
        }
    }
}

struct FinArgs {
    const double *partial;      // [P][G][Zp] (P may be 1 when phi was already reduced / allreduced)
    const double *base;         // [G][Zp]
    const int32_t *mat;         // [Zp]
    const double *scat;         // [M][G][G]
    double *phi;                // [G][Zp] in: previous iterate, out: new iterate
    double *q_next;             // [G][Zp]
    double *scratch;
    ae_ctrl *ctrl;
    int64_t Z, Zp;
    int P, G, M;
    int init;                   // 1: phi := cold (no norm), 0: phi := sum of partials
    int gate;
    int tab_in_smem;
    int iteration;
    double cold, tol;
};

// One thread per cell.  smem: tile[G][kCB] (this cell's new flux for every group) + optional tables.
__global__ void __launch_bounds__(kCB) finalize_kernel(const FinArgs p)
{
    extern __shared__ double sm[];
    __shared__ double red[64];
    if (p.gate && p.ctrl->inner_done) return;
    const int tid = threadIdx.x;
    double *tile = sm;                                   // [G][kCB]
    const double *tab = p.scat;
    if (p.tab_in_smem) {
        double *st = sm + (size_t)p.G * kCB;
        const int n = p.M * p.G * p.G;
        for (int i = tid; i < n; i += kCB) st[i] = p.scat[i];
        tab = st;
    }
    if (p.init && blockIdx.x == 0 && tid == 0) {
        p.ctrl->inner_done = 0;
        p.ctrl->inner_iters = 0;
        p.ctrl->ticket = 0;
    }
    const int64_t s = (int64_t)blockIdx.x * kCB + tid;
    const bool valid = cell_valid(s, p.Z, p.Zp);
    const int64_t n = (int64_t)p.G * p.Zp;
    double num = 0.0, den = 0.0;
    if (s < p.Zp) {
        for (int g = 0; g < p.G; ++g) {
            const int64_t i = (int64_t)g * p.Zp + s;
            double v;
            if (p.init) {
                v = p.cold;
            } else {
                v = p.partial[i];
                for (int pp = 1; pp < p.P; ++pp) v += p.partial[(int64_t)pp * n + i];
            }
            if (!valid) v = 0.0;
            if (!p.init) {
                const double d = v - p.phi[i];          // group.py:78
                num = fma(d, d, num);
                den = fma(v, v, den);
            }
            p.phi[i] = v;                               // group.py:80 phi_old = phi.copy()
            tile[g * kCB + tid] = v;
        }
    }
    __syncthreads();
    if (s < p.Zp) {
        const double *sc = tab + (size_t)p.mat[s] * p.G * p.G;
        for (int g = 0; g < p.G; ++g) {
            double acc = 0.0;
            for (int gp = 0; gp < p.G; ++gp) acc = fma(tile[gp * kCB + tid], sc[gp * p.G + g], acc);
            const int64_t i = (int64_t)g * p.Zp + s;
            p.q_next[i] = p.base[i] + 0.5 * acc;        // group.py:73 / time_unit.py:53
        }
    }
    if (!p.init) {
        const int iteration = p.iteration;
        const double tol = p.tol;
        ae_ctrl *ctrl = p.ctrl;
        grid_sum2(num, den, p.scratch, ctrl, red, [=](double a, double b) {
            const double change = sqrt(a) / sqrt(b);    // group.py:78
            ctrl->change = change;
            ctrl->inner_iters = iteration;
            if (change < tol) ctrl->inner_done = 1;     // group.py:79
        });
    }
}

// base[g][s] = s_ext[g][s] + 0.5*chi[m][g] * sum_g' nusf[m][g'] * phi[g'][s]
// (group.py:104-105, time_unit.py:37-39, zone.py:38-42)
__global__ void __launch_bounds__(kCB) fission_base_kernel(const double *__restrict__ phi, const double *__restrict__ s_ext,
                                                           const int32_t *__restrict__ mat, const double *__restrict__ chi,
                                                           const double *__restrict__ nusf, int64_t Zp, int G,
                                                           double *__restrict__ base)
{
    const int64_t s = (int64_t)blockIdx.x * kCB + threadIdx.x;
    if (s >= Zp) return;
    const int m = mat[s];
    double acc = 0.0;
    for (int gp = 0; gp < G; ++gp) acc = fma(nusf[m * G + gp], phi[(int64_t)gp * Zp + s], acc);
    for (int g = 0; g < G; ++g) {
        const int64_t i = (int64_t)g * Zp + s;
        const double f = 0.5 * chi[m * G + g] * acc;
        base[i] = s_ext ? s_ext[i] + f : f;
    }
}

// k.new = ||nusf*phi_new|| / ||nusf*phi_old||  (group.py:108), convergence flag (group.py:109)
__global__ void __launch_bounds__(kCB) k_update_kernel(const double *__restrict__ phi_new, const double *__restrict__ phi_old,
                                                       const int32_t *__restrict__ mat, const double *__restrict__ nusf,
                                                       int64_t Z, int64_t Zp, int G, double tol, int iteration,
                                                       double *scratch, ae_ctrl *ctrl)
{
    __shared__ double red[64];
    const int64_t s = (int64_t)blockIdx.x * kCB + threadIdx.x;
    double num = 0.0, den = 0.0;
    if (cell_valid(s, Z, Zp)) {
        const int m = mat[s];
        for (int g = 0; g < G; ++g) {
            const double f = nusf[m * G + g];
            const double a = f * phi_new[(int64_t)g * Zp + s], b = f * phi_old[(int64_t)g * Zp + s];
            num = fma(a, a, num);
            den = fma(b, b, den);
        }
    }
    grid_sum2(num, den, scratch, ctrl, red, [=](double a, double b) {
        const double k_new = sqrt(a) / sqrt(b);
        const double dk = fabs(k_new - ctrl->k_old);
        ctrl->k_new = k_new;
        ctrl->k_change = dk;
        ctrl->power_iters = iteration;
        ctrl->power_done = dk < tol ? 1 : 0;            // group.py:109
        ctrl->k_old = k_new;                            // group.py:114
    });
}

// phi.old = phi.new / k.new  (group.py:115)
__global__ void scale_by_k_kernel(const double *__restrict__ phi_new, const ae_ctrl *ctrl, int64_t n, double *__restrict__ phi_old)
{
    const double k = ctrl->k_new;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        phi_old[i] = phi_new[i] / k;
}

// time_unit.py:41 outer change; also phi_outer := phi for the next outer iteration (time_unit.py:36)
__global__ void __launch_bounds__(kCB) outer_check_kernel(const double *__restrict__ phi, double *__restrict__ phi_outer,
                                                          int64_t Zp, int G, double *scratch, ae_ctrl *ctrl)
{
    __shared__ double red[64];
    const int64_t s = (int64_t)blockIdx.x * kCB + threadIdx.x;
    double num = 0.0, den = 0.0;
    if (s < Zp) {
        for (int g = 0; g < G; ++g) {
            const int64_t i = (int64_t)g * Zp + s;
            const double v = phi[i], d = v - phi_outer[i];      // padding cells are 0 in both
            num = fma(d, d, num);
            den = fma(v, v, den);
            phi_outer[i] = v;
        }
    }
    grid_sum2(num, den, scratch, ctrl, red, [=](double a, double b) { ctrl->outer_change = sqrt(a) / sqrt(b); });
}

__global__ void ctrl_init_kernel(ae_ctrl *ctrl)
{
    ctrl->change = 0.0; ctrl->outer_change = 0.0;
    ctrl->k_new = 1.0; ctrl->k_old = 1.1;               // eigenvalue.py:3-4
    ctrl->k_change = 0.0;
    ctrl->inner_done = 0; ctrl->inner_iters = 0; ctrl->power_done = 0; ctrl->power_iters = 0;
    ctrl->ticket = 0; ctrl->reserved = 0;
}

static int per_cell_blocks(int64_t Zp) { return (int)((Zp + kCB - 1) / kCB); }

static int smem_optin(const void *fn, size_t bytes)
{
    if (bytes > 48 * 1024) AE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return AE_OK;
}

int launch_finalize(const ae_slab *s, ae_work *w, const double *partial, int P, double *q_next, int init,
                    double cold, double tol, int iteration, int gate, cudaStream_t st)
{
    FinArgs p;
    p.partial = partial; p.base = w->base; p.mat = s->mat; p.scat = s->scat; p.phi = w->phi; p.q_next = q_next;
    p.scratch = w->norm_scratch; p.ctrl = w->ctrl; p.Z = s->Z; p.Zp = s->Zp; p.P = P; p.G = s->G; p.M = s->M;
    p.init = init; p.gate = gate; p.iteration = iteration; p.cold = cold; p.tol = tol;
    const size_t tile_bytes = (size_t)s->G * kCB * sizeof(double);
    const size_t tab_bytes = (size_t)s->M * s->G * s->G * sizeof(double);
    p.tab_in_smem = (tile_bytes + tab_bytes <= 200 * 1024) ? 1 : 0;
    const size_t smem = tile_bytes + (p.tab_in_smem ? tab_bytes : 0);
    if (smem > 227 * 1024) { set_error("finalize: G=%d needs %zu B of shared memory", s->G, smem); return AE_BAD_ARG; }
    AE_TRY(smem_optin((const void *)finalize_kernel, smem));
    finalize_kernel<<<per_cell_blocks(s->Zp), kCB, smem, st>>>(p);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_fission_base(const ae_slab *s, const double *phi, double *base, cudaStream_t st)
{
    fission_base_kernel<<<per_cell_blocks(s->Zp), kCB, 0, st>>>(phi, s->s_ext, s->mat, s->chi, s->nusf, s->Zp, s->G, base);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_k_update(const ae_slab *s, ae_work *w, double tol, int iteration, cudaStream_t st)
{
    k_update_kernel<<<per_cell_blocks(s->Zp), kCB, 0, st>>>(w->phi, w->phi_outer, s->mat, s->nusf, s->Z, s->Zp, s->G, tol,
                                                            iteration, w->norm_scratch, w->ctrl);
    AE_CUDA(cudaGetLastError());
    const int64_t n = (int64_t)s->G * s->Zp;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    scale_by_k_kernel<<<blocks, 256, 0, st>>>(w->phi, w->ctrl, n, w->phi_outer);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_outer_check(const ae_slab *s, ae_work *w, cudaStream_t st)
{
    outer_check_kernel<<<per_cell_blocks(s->Zp), kCB, 0, st>>>(w->phi, w->phi_outer, s->Zp, s->G, w->norm_scratch, w->ctrl);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_ctrl_init(ae_work *w, cudaStream_t st)
{
    ctrl_init_kernel<<<1, 1, 0, st>>>(w->ctrl);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

}  // namespace ae

extern "C" int64_t ae_norm_blocks(int64_t Zp) { return ae::per_cell_blocks(Zp); }
