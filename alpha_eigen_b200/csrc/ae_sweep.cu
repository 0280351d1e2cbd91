// Batched diamond-difference slab sweep as an affine-recurrence scan (sm_100a, FP64).
//
// Replaces the per-cell Python loop of OneGroup.sweep1D (reference group.py:41-60) and the
// angle loop + quadrature accumulate around it (group.py:74-77, time_unit.py:52-55).
//
// One cell update  psi_out = (q + (m - s/2) psi_in) / (m + s/2),  m = |mu|/hx,  s = sigT
// is the affine map psi_out = a psi_in + b with a = (m - s/2)/(m + s/2) = 2m/(m + s/2) - 1 and
// b = q/(m + s/2).  Maps compose associatively, (a2,b2)o(a1,b1) = (a2 a1, a2 b1 + b2), so a
// row (one angle of one group) is solved by: each lane folds its 8 consecutive cells serially
// (the reference's operation order inside a lane), a 5-step warp-shuffle scan composes the 32
// lane maps, and a single running carry links consecutive 256-cell tiles.  One warp owns one
// (group, angle) row for the whole slab, so there is no inter-CTA dependency; the warps of a
// CTA are angles of the same group and direction and share hs/q/cdt loads through L1, and
// their weighted midpoint fluxes are summed through shared memory once per tile into a
// per-angle-block partial scalar flux (deterministic order).
//
// Memory: cells are stored in the lane-blocked tile order of alpha_eigen_b200.h, so every
// global access below is a fully coalesced 16-byte-per-lane vector access.
// Algorithmic bytes per cell-angle-group update: 8 (psi read) + 8 (psi write) + 24/A
// (hs, q, cdt amortised over the angles) + 8/AC (partial write).
#include "ae_common.cuh"

namespace ae {

struct SweepArgs {
    const double *hs, *cdt, *q, *mu_ihx, *wgt, *psi_in;
    double *psi_out, *partial;
    int64_t Z, Zp;
    int A, n_neg, G, nb_neg;
};

// 1/d to ~1 ulp: MUFU.RCP64H seed (about 20 bits) + one cubic and one quadratic Newton step.
__device__ __forceinline__ double rcp_nr(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    return x;
}

// psi is touched exactly once per sweep: keep it out of L1.
__device__ __forceinline__ double2 ld_stream(const double *p)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

template <int AC, bool READ_PSI, bool WRITE_PSI, bool NEG>
__device__ __forceinline__ void sweep_rows(const SweepArgs &p, int a0, int a_end, int blk, double *red)
{
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    const int a = a0 + warp;
    const bool active = a < a_end;
    const int nactive = min(AC, a_end - a0);
    const double m = active ? fabs(p.mu_ihx[a]) : 1.0;
    const double two_m = 2.0 * m;
    const double hw = active ? 0.5 * p.wgt[a] : 0.0;
    const int64_t nt = p.Zp / kTile;
    const int64_t goff = (int64_t)g * p.Zp + lane * 2;
    const double *hs = p.hs + goff;
    const double *q = p.q + goff;
    const double *cdt = READ_PSI ? p.cdt + goff : nullptr;
    const int64_t row = ((int64_t)g * p.A + (active ? a : a0)) * p.Zp + lane * 2;
    const double *pin = READ_PSI ? p.psi_in + row : nullptr;
    double *pout = WRITE_PSI ? p.psi_out + row : nullptr;
    double *part = p.partial + ((int64_t)blk * p.G + g) * p.Zp;
    double carry = 0.0;     // psi entering the tile (vacuum boundary: group.py:48,54)

    for (int64_t it = 0; it < nt; ++it) {
        const int64_t off = (NEG ? nt - 1 - it : it) * kTile;
        double *redbuf = red + (it & 1) * (AC * kTile);
        if (active) {
            double2 vh[4], vq[4], vc[4], vp[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                vh[k] = __ldg(reinterpret_cast<const double2 *>(hs + off + k * 64));
                vq[k] = __ldg(reinterpret_cast<const double2 *>(q + off + k * 64));
                if (READ_PSI) {
                    vc[k] = __ldg(reinterpret_cast<const double2 *>(cdt + off + k * 64));
                    vp[k] = ld_stream(pin + off + k * 64);
                }
            }
            double ca[kL], cb[kL];
#pragma unroll
            for (int j = 0; j < kL; ++j) {
                const double h = (j & 1) ? vh[j >> 1].y : vh[j >> 1].x;
                double src = (j & 1) ? vq[j >> 1].y : vq[j >> 1].x;
                if (READ_PSI) {
                    const double c = (j & 1) ? vc[j >> 1].y : vc[j >> 1].x;
                    const double pp = (j & 1) ? vp[j >> 1].y : vp[j >> 1].x;
                    src = fma(c, pp, src);              // time_unit.py:23: S + psi*inv_speed/dt
                }
                const double r = rcp_nr(m + h);         // 1/(0.5*sigT + mu/hx), group.py:50
                ca[j] = fma(two_m, r, -1.0);            // (mu/hx - 0.5*sigT)/(0.5*sigT + mu/hx)
                cb[j] = src * r;
            }
            // fold this lane's 8 cells in sweep order
            double Ai, Bi;
            if (!NEG) {
                Ai = ca[0]; Bi = cb[0];
#pragma unroll
                for (int j = 1; j < kL; ++j) { Bi = fma(ca[j], Bi, cb[j]); Ai *= ca[j]; }
            } else {
                Ai = ca[kL - 1]; Bi = cb[kL - 1];
#pragma unroll
                for (int j = kL - 2; j >= 0; --j) { Bi = fma(ca[j], Bi, cb[j]); Ai *= ca[j]; }
            }
            // inclusive scan of the lane maps along the sweep direction
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double Ao = NEG ? __shfl_down_sync(full, Ai, d) : __shfl_up_sync(full, Ai, d);
                const double Bo = NEG ? __shfl_down_sync(full, Bi, d) : __shfl_up_sync(full, Bi, d);
                const bool take = NEG ? (lane + d < 32) : (lane >= d);
                if (take) { Bi = fma(Ai, Bo, Bi); Ai *= Ao; }
            }
            double Ae = NEG ? __shfl_down_sync(full, Ai, 1) : __shfl_up_sync(full, Ai, 1);
            double Be = NEG ? __shfl_down_sync(full, Bi, 1) : __shfl_up_sync(full, Bi, 1);
            if (NEG ? (lane == 31) : (lane == 0)) { Ae = 1.0; Be = 0.0; }
            double in = fma(Ae, carry, Be);
            const double At = __shfl_sync(full, Ai, NEG ? 0 : 31);
            const double Bt = __shfl_sync(full, Bi, NEG ? 0 : 31);
            carry = fma(At, carry, Bt);

            double f[kL];
#pragma unroll
            for (int jj = 0; jj < kL; ++jj) {
                const int j = NEG ? kL - 1 - jj : jj;
                const double out = fma(ca[j], in, cb[j]);
                const double s2 = in + out;             // psi_mid = 0.5*(in+out), group.py:51,57
                f[j] = s2;
                in = out;
            }
            double *rw = redbuf + warp * kTile + lane * 2;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                *reinterpret_cast<double2 *>(rw + k * 64) = make_double2(hw * f[2 * k], hw * f[2 * k + 1]);
            if (WRITE_PSI) {
                const bool ragged = off + kTile > p.Z;  // only the last tile holds padding cells
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double2 v = make_double2(0.5 * f[2 * k], 0.5 * f[2 * k + 1]);
                    if (ragged) {
                        const int64_t z = off + lane * kL + 2 * k;
                        if (z >= p.Z) v.x = 0.0;
                        if (z + 1 >= p.Z) v.y = 0.0;
                    }
                    st_stream(pout + off + k * 64, v);
                }
            }
        }
        __syncthreads();
        // partial scalar flux of this angle block: sum over its warps in angle order (group.py:77)
        for (int c = threadIdx.x; c < kTile; c += AC * 32) {
            double sum = redbuf[c];
            for (int w = 1; w < nactive; ++w) sum += redbuf[w * kTile + c];
            part[off + c] = sum;
        }
        // no second barrier: the next tile writes the other half of `red`
    }
}

template <int AC, bool READ_PSI, bool WRITE_PSI>
__global__ void __launch_bounds__(AC * 32) sweep_kernel(const SweepArgs p, const ae_ctrl *ctrl, int gate)
{
    extern __shared__ double red[];                     // [2][AC][kTile]
    if (gate && ctrl->inner_done) return;               // batch launched past convergence
    const int blk = blockIdx.x;
    if (blk < p.nb_neg) {
        sweep_rows<AC, READ_PSI, WRITE_PSI, true>(p, blk * AC, p.n_neg, blk, red);
    } else {
        sweep_rows<AC, READ_PSI, WRITE_PSI, false>(p, p.n_neg + (blk - p.nb_neg) * AC, p.A, blk, red);
    }
}

// phi[g][s] = sum_p partial[p][g][s]; padding cells -> 0
__global__ void reduce_partials_kernel(const double *partial, int P, int64_t Z, int64_t Zp, int G,
                                       double *phi)   // phi may alias partial[0]
{
    const int64_t n = (int64_t)G * Zp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = partial[i];
        for (int p = 1; p < P; ++p) s += partial[(int64_t)p * n + i];
        phi[i] = logical_cell(i % Zp) < Z ? s : 0.0;
    }
}

int angles_per_cta(int A, int n_neg)
{
    const int side = max(n_neg, A - n_neg);
    int ac = 1;
    while (ac < 8 && ac < side) ac <<= 1;
    return ac;
}

template <int AC>
static int launch_ac(const SweepArgs &p, const ae_ctrl *ctrl, int gate, cudaStream_t st)
{
    const int nb_pos = (p.A - p.n_neg + AC - 1) / AC;
    dim3 grid(p.nb_neg + nb_pos, p.G);
    const size_t smem = 2 * AC * kTile * sizeof(double);
    const bool rd = p.psi_in != nullptr, wr = p.psi_out != nullptr;
    if (rd && wr) sweep_kernel<AC, true, true><<<grid, AC * 32, smem, st>>>(p, ctrl, gate);
    else if (rd) sweep_kernel<AC, true, false><<<grid, AC * 32, smem, st>>>(p, ctrl, gate);
    else if (wr) sweep_kernel<AC, false, true><<<grid, AC * 32, smem, st>>>(p, ctrl, gate);
    else sweep_kernel<AC, false, false><<<grid, AC * 32, smem, st>>>(p, ctrl, gate);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

// internal launcher shared with the solver loops (gate: skip when ctrl->inner_done is set)
int launch_sweep(const ae_slab *s, const double *q, const double *psi_in, double *psi_out, double *partial,
                 const ae_ctrl *ctrl, int gate, cudaStream_t st)
{
    if (!s || !q || !partial || s->Zp % kTile || s->A <= 0 || s->G <= 0) {
        set_error("ae_sweep_set: bad argument");
        return AE_BAD_ARG;
    }
    if (psi_in && !s->cdt) { set_error("ae_sweep_set: psi_in given but slab has no cdt"); return AE_BAD_ARG; }
    SweepArgs p;
    p.hs = s->hs; p.cdt = s->cdt; p.q = q; p.mu_ihx = s->mu_ihx; p.wgt = s->wgt;
    p.psi_in = psi_in; p.psi_out = psi_out; p.partial = partial;
    p.Z = s->Z; p.Zp = s->Zp; p.A = s->A; p.n_neg = s->n_neg; p.G = s->G;
    const int ac = angles_per_cta(s->A, s->n_neg);
    p.nb_neg = (s->n_neg + ac - 1) / ac;
    switch (ac) {
        case 1: return launch_ac<1>(p, ctrl, gate, st);
        case 2: return launch_ac<2>(p, ctrl, gate, st);
        case 4: return launch_ac<4>(p, ctrl, gate, st);
        default: return launch_ac<8>(p, ctrl, gate, st);
    }
}

}  // namespace ae

extern "C" int32_t ae_partial_count(int32_t A, int32_t n_neg)
{
    const int ac = ae::angles_per_cta(A, n_neg);
    return (n_neg + ac - 1) / ac + (A - n_neg + ac - 1) / ac;
}

extern "C" int ae_sweep_set(const ae_slab *s, const double *q, const double *psi_in, double *psi_out,
                            double *partial, void *stream)
{
    return ae::launch_sweep(s, q, psi_in, psi_out, partial, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int ae_reduce_partials(const ae_slab *s, const double *partial, int32_t P, double *phi, void *stream)
{
    if (!s || !partial || !phi || P < 1) { ae::set_error("ae_reduce_partials: bad argument"); return AE_BAD_ARG; }
    const int64_t n = (int64_t)s->G * s->Zp;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    ae::reduce_partials_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(partial, P, s->Z, s->Zp, s->G, phi);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}
