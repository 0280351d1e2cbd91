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
//
// Multi-GPU: every kernel works on a CELL RANGE [c0, c0+cn) of every group row.  With one GPU the
// range is the whole row.  With N ranks each rank owns Zp/N cells: the partial fluxes are
// reduce-scattered by rows so that rank r receives the summed flux of its cells only, does 1/N of
// the update work, and the new sources are all-gathered; the two sums of a norm travel as a
// (num, den) pair in the same NCCL group and are combined in rank order on every rank.
// With enough groups (G >= 2 N) the ranks are sharded by GROUPS instead: a rank sweeps its groups with all angles, so
// their flux is complete on it -- no reduction; it assembles those rows over all cells, the rows are all-gathered (one
// collective per iteration instead of two) and it builds the scatter / fission sources of its own groups from the
// full flux.
#include <stdlib.h>
#include "ae_common.cuh"

namespace ae {

constexpr int kCB = 128;        // cells (threads) per block in the per-cell kernels

enum NormMode { kNormInner = 0, kNormK = 1, kNormOuter = 2 };

// What a finished norm does to the control block (thread 0 of the last block, or norm_finalize_kernel).
__device__ __forceinline__ void apply_norm(int mode, double a, double b, ae_ctrl *ctrl, double tol, int iteration)
{
    if (mode == kNormInner) {
        const double change = sqrt(a) / sqrt(b);        // group.py:78
        ctrl->change = change;
        ctrl->inner_iters = iteration;
        if (change < tol) ctrl->inner_done = 1;         // group.py:79
    } else if (mode == kNormK) {
        const double k_new = sqrt(a) / sqrt(b);         // group.py:108
        const double dk = fabs(k_new - ctrl->k_old);
        ctrl->k_new = k_new;
        ctrl->k_change = dk;
        ctrl->power_iters = iteration;
        ctrl->power_done = dk < tol ? 1 : 0;            // group.py:109
        ctrl->k_old = k_new;                            // group.py:114
    } else {
        ctrl->outer_change = sqrt(a) / sqrt(b);         // time_unit.py:41
    }
}

struct NormOut {                // where a norm goes
    double *scratch;            // per-block partials
    ae_ctrl *ctrl;
    double *pair_local;         // != NULL: sharded, publish this rank's (num, den) here instead of finishing
    double tol;
    int mode, iteration;
};

// Deterministic grid-wide reduction of (num, den): per-block partials to scratch, last block sums them
// in a fixed order.
__device__ __forceinline__ void grid_sum2(double num, double den, const NormOut &o, double *sm)
{
    __shared__ bool is_last;
    const unsigned nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    block_sum2(num, den, sm);
    if (threadIdx.x == 0) {
        o.scratch[2 * bid] = num;
        o.scratch[2 * bid + 1] = den;
        __threadfence();
        const unsigned t = atomicAdd(&o.ctrl->ticket, 1u);
        is_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < (int)nblocks; i += blockDim.x) {
            a += __ldcg(o.scratch + 2 * i);
            b += __ldcg(o.scratch + 2 * i + 1);
        }
        block_sum2(a, b, sm);
        if (threadIdx.x == 0) {
            o.ctrl->ticket = 0;
            if (o.pair_local) { o.pair_local[0] = a; o.pair_local[1] = b; }
            else apply_norm(o.mode, a, b, o.ctrl, o.tol, o.iteration);
        }
    }
}

// Sharded norms: combine the ranks' pairs in rank order (identical on every rank).
constexpr int kPairSlots = 4;       // (num, den) pairs per rank: one per group chunk (ae_comm.cu)
__global__ void norm_finalize_kernel(const double *pairs, int nranks, int nslots, int mode, ae_ctrl *ctrl, double tol,
                                     int iteration, int gate)
{
    if (gate && ctrl->inner_done) return;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < nranks; ++r)
        for (int c = 0; c < nslots; ++c) { a += pairs[2 * (r * kPairSlots + c)]; b += pairs[2 * (r * kPairSlots + c) + 1]; }
    apply_norm(mode, a, b, ctrl, tol, iteration);
}

struct FinArgs {
    const double *partial;      // [P][G][Zp] (P may be 1 when phi was already reduced over angle blocks / ranks)
    const double *base;         // [G][Zp]
    const double *fixed;        // [G][Zp] added to the sum of the partials (ae_work.phi_fixed) or NULL
    const int32_t *mat;         // [Zp]
    const double *scat;         // [M][G][G]
    const double *phi_in;       // [G][Zp] previous iterate
    double *phi;                // [G][Zp] new iterate.  A different buffer: on B200 reading and then writing the
                                // same addresses in one streaming kernel runs at 2.4 TB/s instead of 6.8 (tools/stride_probe.cu)
    double *q_next;             // [G][Zp]
    NormOut norm;
    int64_t Z, Zp;
    int64_t c0, cn;             // cell range of this rank (multiples of 256)
    int g0, gl;                 // group rows this rank updates: all G, or its own groups when group-sharded
                                // (then partial holds only those gl rows per slab)
    int P, G, M;
    int init;                   // 1: phi := cold (no norm), 0: phi := sum of partials
    int gate;
    double cold;
};

constexpr int kGSmall = 8;      // up to this many groups one thread owns a whole cell

// G <= kGSmall: one thread per cell does everything (flux assembly, norm, scatter source).
// This is the latency-critical path of the one-group Kornreich-Parsons problems.
__global__ void __launch_bounds__(kCB) finalize_small_kernel(const FinArgs p)
{
    __shared__ double red[64];
    if (p.gate && p.norm.ctrl->inner_done) return;
    const int tid = threadIdx.x;
    if (p.init && blockIdx.x == 0 && tid == 0) {
        p.norm.ctrl->inner_done = 0;
        p.norm.ctrl->inner_iters = 0;
        p.norm.ctrl->ticket = 0;
    }
    const int64_t s = p.c0 + (int64_t)blockIdx.x * kCB + tid;
    const bool inside = s < p.c0 + p.cn;
    const bool valid = inside && logical_cell(s) < p.Z;
    const int64_t n = (int64_t)p.G * p.Zp;
    double num = 0.0, den = 0.0;
    if (inside) {
        double v[kGSmall];
#pragma unroll
        for (int g = 0; g < kGSmall; ++g) {
            v[g] = 0.0;
            if (g < p.G) {
                const int64_t i = (int64_t)g * p.Zp + s;
                double x;
                if (p.init) {
                    x = p.cold;
                } else {
                    x = p.partial[i];
                    for (int pp = 1; pp < p.P; ++pp) x += p.partial[(int64_t)pp * n + i];
                    if (p.fixed) x += p.fixed[i];
                }
                if (!valid) x = 0.0;
                if (!p.init) {
                    const double d = x - p.phi_in[i];   // group.py:78
                    num = fma(d, d, num);
                    den = fma(x, x, den);
                }
                p.phi[i] = x;                           // group.py:80 phi_old = phi.copy()
                v[g] = x;
            }
        }
        const double *sc = p.scat + (size_t)p.mat[s] * p.G * p.G;
#pragma unroll
        for (int g = 0; g < kGSmall; ++g) {
            if (g < p.G) {
                double acc = 0.0;
#pragma unroll
                for (int gp = 0; gp < kGSmall; ++gp)
                    if (gp < p.G) acc = fma(v[gp], __ldg(sc + gp * p.G + g), acc);
                const int64_t i = (int64_t)g * p.Zp + s;
                p.q_next[i] = p.base[i] + 0.5 * acc;    // group.py:73 / time_unit.py:53
            }
        }
    }
    if (!p.init) grid_sum2(num, den, p.norm, red);
}

// G > kGSmall, step 1: phi := sum of partials, norm.  Grid (x over cell pairs of the range, y = group), one
// double2 per thread and iteration, 40 registers so 48 warps per SM with ~4 loads each are in flight: a
// P-buffer sum needs ~100 KB per SM in flight to stream at 6.5-7 TB/s on B200 (tools/stride_probe.cu).
// Padding cells need no masking here: the sweep writes exact zeros there.
__global__ void __launch_bounds__(256, 6) flux_assemble_kernel(const FinArgs p)
{
    __shared__ double red[64];
    if (p.gate && p.norm.ctrl->inner_done) return;
    const int64_t n = (int64_t)p.gl * p.Zp;                      // slab stride of the partials (this rank's rows)
    const int64_t row = (int64_t)(p.g0 + blockIdx.y) * p.Zp + p.c0;
    const double *part = p.partial + (int64_t)blockIdx.y * p.Zp + p.c0;
    const double *fix = p.fixed ? p.fixed + row : nullptr;
    const double *old = p.phi_in + row;
    double *out = p.phi + row;
    double num = 0.0, den = 0.0;
    for (int64_t i = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); i < p.cn; i += 2 * (int64_t)gridDim.x * blockDim.x) {
        double2 x = *reinterpret_cast<const double2 *>(part + i);
#pragma unroll 4
        for (int pp = 1; pp < p.P; ++pp) {                       // block order (group.py:77 over angle blocks)
            const double2 y = *reinterpret_cast<const double2 *>(part + (int64_t)pp * n + i);
            x.x += y.x; x.y += y.y;
        }
        if (fix) {                                               // the part of the flux psi^n drives (once per step)
            const double2 y = *reinterpret_cast<const double2 *>(fix + i);
            x.x += y.x; x.y += y.y;
        }
        const double2 o = *reinterpret_cast<const double2 *>(old + i);
        const double dx = x.x - o.x, dy = x.y - o.y;             // group.py:78
        num = fma(dx, dx, fma(dy, dy, num));
        den = fma(x.x, x.x, fma(x.y, x.y, den));
        *reinterpret_cast<double2 *>(out + i) = x;               // group.py:80
    }
    grid_sum2(num, den, p.norm, red);
}

// Cold start of the inner iteration: phi := cold on real cells, 0 on padding (group.py:66 / time_unit.py:45).
__global__ void flux_init_kernel(const FinArgs p)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.norm.ctrl->inner_done = 0;
        p.norm.ctrl->inner_iters = 0;
        p.norm.ctrl->ticket = 0;
    }
    const int64_t n = (int64_t)p.G * p.Zp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p.phi[i] = logical_cell(i % p.Zp) < p.Z ? p.cold : 0.0;
}

// G > kGSmall, step 2: q_next[g][s] = base[g][s] + 0.5 * sum_g' phi[g'][s] * sigma_s[m(s)][g'][g]
// (group.py:73 / time_unit.py:53 with the scatter matrix of zone.py:48-51).
// CTA = 128 consecutive cells x all groups; warp y owns the block of kGB = 8 output groups y and every lane
// 4 cells (lane, +32, +64, +96), i.e. 32 accumulators per thread: per source group g' a thread issues 4
// coalesced phi loads (shared by the warps of the CTA through L1) and 4 broadcast LDS.128 of the scatter
// row, for 32 FMAs -- FP64-pipe bound instead of load-issue bound.  The scatter matrix of the tile's
// material is staged once per CTA in shared memory (row stride padded to a multiple of 8); the rare
// tile that straddles a material interface takes the per-cell global-memory path.
// (Round 2 tried this product on the FP64 tensor path -- mma.sync m8n8k4 with S^T and the flux tile in shared memory,
// zero blocks of S^T masked out, one and two CTAs per SM, fused with the flux assembly or not: 217-340 us per C3
// iteration against 193 us for this kernel.  The product is not what limits it: 3 x G separate 1 KB runs per tile is a
// poor DRAM access pattern, see DESIGN.md.)
constexpr int kGB = 8;
constexpr int kSC = 128;                // cells per CTA
constexpr int kSWarps = 10;             // warps per CTA (each loops over its group blocks)
__global__ void __launch_bounds__(32 * kSWarps, 2) scatter_source_kernel(const FinArgs p)
{
    extern __shared__ double stab[];                            // [G][Gp] scatter matrix, then [G][kSC] phi tile
    __shared__ int s_uniform;
    if (p.gate && p.norm.ctrl->inner_done) return;
    const int lane = threadIdx.x, gb = threadIdx.y, tid = gb * 32 + lane, nthr = blockDim.y * 32;
    const int Gp = (p.G + kGB - 1) / kGB * kGB;
    const int gblocks = Gp / kGB;
    double *sphi = stab + p.G * Gp;                             // [G][kSC]
    const double *phi = sphi + lane;
    int staged = -1;                                            // material whose matrix is in stab
    // persistent CTA: the scatter matrix is restaged only when the tile's material changes (5 regions)
    for (int64_t tile = blockIdx.x; tile < p.cn / kSC; tile += gridDim.x) {
        const int64_t s0 = p.c0 + tile * kSC;
        const int m0 = p.mat[s0];
        __syncthreads();                                        // previous tile fully consumed
        if (tid == 0) s_uniform = 1;
        __syncthreads();
        for (int c = tid; c < kSC; c += nthr)
            if (p.mat[s0 + c] != m0) s_uniform = 0;             // benign race: everyone writes 0
        if (m0 != staged) {
            const double *gtab = p.scat + (size_t)m0 * p.G * p.G;
            for (int i = tid; i < p.G * Gp; i += nthr) {
                const int gp = i / Gp, g = i % Gp;
                stab[i] = g < p.G ? gtab[gp * p.G + g] : 0.0;
            }
            staged = m0;
        }
        for (int i = tid; i < p.G * (kSC / 2); i += nthr) {     // coalesced double2 rows of 128 cells
            const int gp = i / (kSC / 2), c2 = i % (kSC / 2);
            *reinterpret_cast<double2 *>(sphi + gp * kSC + 2 * c2) =
                *reinterpret_cast<const double2 *>(p.phi + (int64_t)gp * p.Zp + s0 + 2 * c2);
        }
        __syncthreads();
        // output group blocks that overlap this rank's groups (all of them unless group-sharded)
        const int b_lo = p.g0 / kGB, nb = (p.g0 + p.gl - 1) / kGB - b_lo + 1;
        if (p.gl >= p.G) {
            // a block of 8 output groups x 128 cells (4 per lane) per warp and step
            for (int bi = gb; bi < nb; bi += blockDim.y) {
                const int g0 = (b_lo + bi) * kGB;
                const int ng = min(kGB, p.G - g0);
                double acc[4][kGB];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int j = 0; j < kGB; ++j) acc[c][j] = 0.0;
                if (s_uniform) {
#pragma unroll 2
                    for (int gp = 0; gp < p.G; ++gp) {
                        double x[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) x[c] = phi[gp * kSC + 32 * c];
                        const double2 *row = reinterpret_cast<const double2 *>(stab + gp * Gp + g0);
#pragma unroll
                        for (int j2 = 0; j2 < kGB / 2; ++j2) {
                            const double2 t = row[j2];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                acc[c][2 * j2] = fma(x[c], t.x, acc[c][2 * j2]);
                                acc[c][2 * j2 + 1] = fma(x[c], t.y, acc[c][2 * j2 + 1]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double *sc = p.scat + (size_t)p.mat[s0 + lane + 32 * c] * p.G * p.G + g0;
                        for (int gp = 0; gp < p.G; ++gp) {
                            const double x = phi[gp * kSC + 32 * c];
#pragma unroll
                            for (int j = 0; j < kGB; ++j)
                                if (j < ng) acc[c][j] = fma(x, __ldg(sc + gp * p.G + j), acc[c][j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kGB; ++j)
                    if (j < ng && g0 + j >= p.g0 && g0 + j < p.g0 + p.gl) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int64_t i = (int64_t)(g0 + j) * p.Zp + s0 + lane + 32 * c;
                            p.q_next[i] = p.base[i] + 0.5 * acc[c][j];
                        }
                    }
            }
        } else {
            // group-sharded rank: a warp takes (block, 32-cell column) items so that every warp of the CTA has work
            // (whole blocks would leave most warps idle: 2 blocks for 10 warps at 8 ranks)
            for (int it = gb; it < nb * 4; it += blockDim.y) {
                const int g0 = (b_lo + it / 4) * kGB, c = it % 4;
                const int ng = min(kGB, p.G - g0);
                double acc[kGB];
#pragma unroll
                for (int j = 0; j < kGB; ++j) acc[j] = 0.0;
                if (s_uniform) {
#pragma unroll 4
                    for (int gp = 0; gp < p.G; ++gp) {
                        const double x = phi[gp * kSC + 32 * c];
                        const double2 *row = reinterpret_cast<const double2 *>(stab + gp * Gp + g0);
#pragma unroll
                        for (int j2 = 0; j2 < kGB / 2; ++j2) {
                            const double2 t = row[j2];
                            acc[2 * j2] = fma(x, t.x, acc[2 * j2]);
                            acc[2 * j2 + 1] = fma(x, t.y, acc[2 * j2 + 1]);
                        }
                    }
                } else {
                    const double *sc = p.scat + (size_t)p.mat[s0 + lane + 32 * c] * p.G * p.G + g0;
                    for (int gp = 0; gp < p.G; ++gp) {
                        const double x = phi[gp * kSC + 32 * c];
#pragma unroll
                        for (int j = 0; j < kGB; ++j)
                            if (j < ng) acc[j] = fma(x, __ldg(sc + gp * p.G + j), acc[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < kGB; ++j)
                    if (j < ng && g0 + j >= p.g0 && g0 + j < p.g0 + p.gl) {
                        const int64_t i = (int64_t)(g0 + j) * p.Zp + s0 + lane + 32 * c;
                        p.q_next[i] = p.base[i] + 0.5 * acc[j];
                    }
            }
        }
    }
}

// base[g][s] = s_ext[g][s] + 0.5*chi[m][g] * sum_g' nusf[m][g'] * phi[g'][s]
// (group.py:104-105, time_unit.py:37-39, zone.py:38-42)
__global__ void __launch_bounds__(kCB) fission_base_kernel(const double *__restrict__ phi, const double *__restrict__ s_ext,
                                                           const int32_t *__restrict__ mat, const double *__restrict__ chi,
                                                           const double *__restrict__ nusf, int64_t Zp, int64_t c0, int64_t cn,
                                                           int G, int g0, int gl, double *__restrict__ base)
{
    const int64_t s = c0 + (int64_t)blockIdx.x * kCB + threadIdx.x;
    if (s >= c0 + cn) return;
    const int m = mat[s];
    double acc = 0.0;
    for (int gp = 0; gp < G; ++gp) acc = fma(nusf[m * G + gp], phi[(int64_t)gp * Zp + s], acc);
    for (int g = g0; g < g0 + gl; ++g) {                        // the rows this rank sweeps (all, unless group-sharded)
        const int64_t i = (int64_t)g * Zp + s;
        const double f = 0.5 * chi[m * G + g] * acc;
        base[i] = s_ext ? s_ext[i] + f : f;
    }
}

// k.new = ||nusf*phi_new|| / ||nusf*phi_old||  (group.py:108), convergence flag (group.py:109)
__global__ void __launch_bounds__(kCB) k_update_kernel(const double *__restrict__ phi_new, const double *__restrict__ phi_old,
                                                       const int32_t *__restrict__ mat, const double *__restrict__ nusf,
                                                       int64_t Z, int64_t Zp, int64_t c0, int64_t cn, int G, int g0, int gl,
                                                       NormOut norm)
{
    __shared__ double red[64];
    const int64_t s = c0 + (int64_t)blockIdx.x * kCB + threadIdx.x;
    double num = 0.0, den = 0.0;
    if (s < c0 + cn && logical_cell(s) < Z) {
        const int m = mat[s];
        for (int g = g0; g < g0 + gl; ++g) {                    // this rank's share of the two sums
            const double f = nusf[m * G + g];
            const double a = f * phi_new[(int64_t)g * Zp + s], b = f * phi_old[(int64_t)g * Zp + s];
            num = fma(a, a, num);
            den = fma(b, b, den);
        }
    }
    grid_sum2(num, den, norm, red);
}

// phi.old = phi.new / k.new  (group.py:115)
__global__ void scale_by_k_kernel(const double *__restrict__ phi_new, const ae_ctrl *ctrl, int64_t n, double *__restrict__ phi_old)
{
    const double k = ctrl->k_new;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        phi_old[i] = phi_new[i] / k;
}

// time_unit.py:41 outer change; also phi_outer := phi for the next outer iteration (time_unit.py:36).  Flat over
// the (group, cell pair) entries of the rank's cell range: grid (x over cell pairs, y = group), many loads in flight.
__global__ void __launch_bounds__(256) outer_check_kernel(const double *__restrict__ phi, double *__restrict__ phi_outer,
                                                          int64_t Zp, int64_t c0, int64_t cn, int g0, int gl, NormOut norm)
{
    __shared__ double red[64];
    const int64_t row = (int64_t)blockIdx.y * Zp + c0;
    const bool mine = (int)blockIdx.y >= g0 && (int)blockIdx.y < g0 + gl;   // group-sharded: rows of other ranks are only copied
    double num = 0.0, den = 0.0;
    for (int64_t i = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); i < cn; i += 2 * (int64_t)gridDim.x * blockDim.x) {
        const double2 v = *reinterpret_cast<const double2 *>(phi + row + i);
        const double2 o = *reinterpret_cast<const double2 *>(phi_outer + row + i);     // padding cells are 0 in both
        const double dx = v.x - o.x, dy = v.y - o.y;
        if (mine) {
            num = fma(dx, dx, fma(dy, dy, num));
            den = fma(v.x, v.x, fma(v.y, v.y, den));
        }
        *reinterpret_cast<double2 *>(phi_outer + row + i) = v;
    }
    grid_sum2(num, den, norm, red);
}

__global__ void ctrl_init_kernel(ae_ctrl *ctrl)
{
    ctrl->change = 0.0; ctrl->outer_change = 0.0;
    ctrl->k_new = 1.0; ctrl->k_old = 1.1;               // eigenvalue.py:3-4
    ctrl->k_change = 0.0;
    ctrl->inner_done = 0; ctrl->inner_iters = 0; ctrl->power_done = 0; ctrl->power_iters = 0;
    ctrl->ticket = 0; ctrl->reserved = 0;
}

// ---- host side -----------------------------------------------------------------------------------
int comm_rank(const void *comm);
int comm_size(const void *comm);
double *comm_pair_local(void *comm);
double *comm_pairs(void *comm);
int comm_all_gather_rows(void *comm, double *rows, int G, int64_t Zp, int with_pairs, cudaStream_t st);

static int per_cell_blocks(int64_t cells) { return (int)((cells + kCB - 1) / kCB); }

struct Shard { int64_t c0, cn; int g0, gl; };
static Shard shard_of(const ae_slab *s, const ae_work *w)
{
    if (s->gl > 0) return Shard{0, s->Zp, s->g0, s->gl};         // group-sharded: own groups, all cells
    const int n = comm_size(w->comm), r = comm_rank(w->comm);
    const int64_t cn = s->Zp / n;
    return Shard{r * cn, cn, 0, s->G};                           // angle-sharded sweeps: own cells, all groups
}
int comm_all_gather_groups(void *comm, double *rows, int G, int64_t Zp, int with_pairs, cudaStream_t st);

static NormOut norm_out(ae_work *w, int mode, double tol, int iteration)
{
    return NormOut{w->norm_scratch, w->ctrl, w->comm ? comm_pair_local(w->comm) : nullptr, tol, mode, iteration};
}

// after a sharded norm kernel: all-gather the pairs (and, for the inner iteration, the new source rows in the
// same NCCL group), then let every rank apply the combined norm to its control block
static int finish_sharded_norm(const ae_slab *s, ae_work *w, double *rows, int mode, double tol, int iteration, int gate,
                               cudaStream_t st)
{
    if (s->gl > 0) AE_TRY(comm_all_gather_groups(w->comm, rows, s->G, s->Zp, 1, st));
    else AE_TRY(comm_all_gather_rows(w->comm, rows, s->G, s->Zp, 1, st));
    norm_finalize_kernel<<<1, 1, 0, st>>>(comm_pairs(w->comm), comm_size(w->comm), 1, mode, w->ctrl, tol, iteration, gate);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_finalize(const ae_slab *s, ae_work *w, const double *partial, int P, const double *fixed, const double *phi_in,
                    double *phi_out, double *q_next, int init, double cold, double tol, int iteration, int gate, cudaStream_t st)
{
    const Shard sh = shard_of(s, w);
    if (s->gl <= 0 && s->Zp % (kTile * comm_size(w->comm))) { set_error("Zp must be a multiple of 256 * ranks"); return AE_BAD_ARG; }
    if (s->gl > 0 && s->G <= kGSmall) { set_error("group sharding needs more than %d groups", kGSmall); return AE_BAD_ARG; }
    FinArgs p;
    p.partial = partial; p.base = w->base; p.fixed = fixed; p.mat = s->mat; p.scat = s->scat; p.phi_in = phi_in; p.phi = phi_out;
    p.q_next = q_next; p.norm = norm_out(w, kNormInner, tol, iteration);
    p.Z = s->Z; p.Zp = s->Zp; p.c0 = sh.c0; p.cn = sh.cn; p.g0 = sh.g0; p.gl = sh.gl; p.P = P; p.G = s->G; p.M = s->M;
    p.init = init; p.gate = gate; p.cold = cold;
    if (s->G <= kGSmall) {
        finalize_small_kernel<<<per_cell_blocks(sh.cn), kCB, 0, st>>>(p);
        AE_CUDA(cudaGetLastError());
    } else {
        if (init) {
            const int64_t n = (int64_t)s->G * s->Zp;
            int64_t bi = (n + 255) / 256;
            if (bi > 148 * 32) bi = 148 * 32;
            flux_init_kernel<<<(unsigned)bi, 256, 0, st>>>(p);
        } else {
            // norm_scratch holds 2*per_cell_blocks(Zp) + G (num, den) pairs
            int64_t bx = (sh.cn / 2 + 255) / 256;
            const int64_t cap = (2 * (int64_t)per_cell_blocks(s->Zp)) / s->G;
            if (bx > cap) bx = cap;
            if (bx > (148 * 48 + s->G - 1) / s->G) bx = (148 * 48 + s->G - 1) / s->G;
            if (bx < 1) bx = 1;
            flux_assemble_kernel<<<dim3((unsigned)bx, sh.gl), 256, 0, st>>>(p);
        }
        AE_CUDA(cudaGetLastError());
        // group-sharded: the scatter source needs every group's new flux -- the one exchange of the iteration
        if (w->comm && s->gl > 0 && !init) AE_TRY(finish_sharded_norm(s, w, phi_out, kNormInner, tol, iteration, gate, st));
        const int gblocks = (s->G + kGB - 1) / kGB;
        const size_t smem = (size_t)s->G * (gblocks * kGB + kSC) * sizeof(double);
        if (smem > 220 * 1024) { set_error("flux update: scatter matrix of G=%d does not fit in shared memory", s->G); return AE_BAD_ARG; }
        if (smem > 48 * 1024)
            AE_CUDA(cudaFuncSetAttribute((const void *)scatter_source_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t sgrid = sh.cn / kSC;                         // the range is a multiple of 256 cells
        if (sgrid > 148 * 2) sgrid = 148 * 2;
        scatter_source_kernel<<<(unsigned)sgrid, dim3(32, gblocks < kSWarps ? gblocks : kSWarps), smem, st>>>(p);
        AE_CUDA(cudaGetLastError());
    }
    if (w->comm && s->gl <= 0) {
        if (init) AE_TRY(comm_all_gather_rows(w->comm, q_next, s->G, s->Zp, 0, st));
        else AE_TRY(finish_sharded_norm(s, w, q_next, kNormInner, tol, iteration, gate, st));
    }
    return AE_OK;
}

// ---- pieces of the overlapped group-sharded step (ae_sweep_and_update) --------------------------------------------
// flux assembly of group rows [g0, g0+gl) from `partial` ([P][gl][Zp]); the (num, den) pair goes to slot `slot`
int launch_assemble_rows(const ae_slab *s, ae_work *w, const double *partial, int P, const double *fixed, const double *phi_in,
                         double *phi_out, int g0, int gl, int slot, cudaStream_t st)
{
    FinArgs p;
    p.partial = partial; p.base = w->base; p.fixed = fixed; p.mat = s->mat; p.scat = s->scat; p.phi_in = phi_in; p.phi = phi_out;
    p.q_next = nullptr; p.norm = norm_out(w, kNormInner, 0.0, 0);
    p.norm.pair_local = comm_pair_local(w->comm) + 2 * slot;
    p.Z = s->Z; p.Zp = s->Zp; p.c0 = 0; p.cn = s->Zp; p.g0 = g0; p.gl = gl; p.P = P; p.G = s->G; p.M = s->M;
    p.init = 0; p.gate = 0; p.cold = 0.0;
    int64_t bx = (s->Zp / 2 + 255) / 256;
    const int64_t cap = (2 * (int64_t)per_cell_blocks(s->Zp)) / s->G;
    if (bx > cap) bx = cap;
    if (bx > (148 * 48 + gl - 1) / gl) bx = (148 * 48 + gl - 1) / gl;
    if (bx < 1) bx = 1;
    flux_assemble_kernel<<<dim3((unsigned)bx, gl), 256, 0, st>>>(p);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

// after the chunks' pairs have been all-gathered: the norm, then the scatter source of this rank's groups
int launch_norm_and_scatter(const ae_slab *s, ae_work *w, const double *phi, double *q_next, int nslots, double tol, int iteration,
                            cudaStream_t st)
{
    norm_finalize_kernel<<<1, 1, 0, st>>>(comm_pairs(w->comm), comm_size(w->comm), nslots, kNormInner, w->ctrl, tol, iteration, 0);
    AE_CUDA(cudaGetLastError());
    FinArgs p;
    p.partial = nullptr; p.base = w->base; p.fixed = nullptr; p.mat = s->mat; p.scat = s->scat; p.phi_in = nullptr;
    p.phi = const_cast<double *>(phi); p.q_next = q_next; p.norm = norm_out(w, kNormInner, tol, iteration);
    p.Z = s->Z; p.Zp = s->Zp; p.c0 = 0; p.cn = s->Zp; p.g0 = s->g0; p.gl = s->gl; p.P = 0; p.G = s->G; p.M = s->M;
    p.init = 0; p.gate = 0; p.cold = 0.0;
    const int gblocks = (s->G + kGB - 1) / kGB;
    const size_t smem = (size_t)s->G * (gblocks * kGB + kSC) * sizeof(double);
    if (smem > 220 * 1024) { set_error("flux update: scatter matrix of G=%d does not fit in shared memory", s->G); return AE_BAD_ARG; }
    if (smem > 48 * 1024)
        AE_CUDA(cudaFuncSetAttribute((const void *)scatter_source_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t sgrid = s->Zp / kSC;
    if (sgrid > 148 * 2) sgrid = 148 * 2;
    scatter_source_kernel<<<(unsigned)sgrid, dim3(32, gblocks < kSWarps ? gblocks : kSWarps), smem, st>>>(p);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_fission_base(const ae_slab *s, ae_work *w, const double *phi, double *base, int with_ext, cudaStream_t st)
{
    const Shard sh = shard_of(s, w);
    fission_base_kernel<<<per_cell_blocks(sh.cn), kCB, 0, st>>>(phi, with_ext ? s->s_ext : nullptr, s->mat, s->chi, s->nusf,
                                                                 s->Zp, sh.c0, sh.cn, s->G, sh.g0, sh.gl, base);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_k_update(const ae_slab *s, ae_work *w, double tol, int iteration, cudaStream_t st)
{
    const Shard sh = shard_of(s, w);
    k_update_kernel<<<per_cell_blocks(sh.cn), kCB, 0, st>>>(w->phi, w->phi_outer, s->mat, s->nusf, s->Z, s->Zp, sh.c0, sh.cn,
                                                             s->G, sh.g0, sh.gl, norm_out(w, kNormK, tol, iteration));
    AE_CUDA(cudaGetLastError());
    if (w->comm) AE_TRY(finish_sharded_norm(s, w, nullptr, kNormK, tol, iteration, 0, st));
    const int64_t n = (int64_t)s->G * s->Zp;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    scale_by_k_kernel<<<blocks, 256, 0, st>>>(w->phi, w->ctrl, n, w->phi_outer);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

int launch_outer_check(const ae_slab *s, ae_work *w, cudaStream_t st)
{
    const Shard sh = shard_of(s, w);
    // norm_scratch holds 2 * (per_cell_blocks(Zp) + 512) (num, den) pairs
    int64_t bx = (sh.cn / 2 + 255) / 256;
    const int64_t cap = (2 * ((int64_t)per_cell_blocks(s->Zp) + 512)) / s->G;
    if (bx > cap) bx = cap;
    if (bx > (148 * 16 + s->G - 1) / s->G) bx = (148 * 16 + s->G - 1) / s->G;
    if (bx < 1) bx = 1;
    outer_check_kernel<<<dim3((unsigned)bx, s->G), 256, 0, st>>>(w->phi, w->phi_outer, s->Zp, sh.c0, sh.cn, sh.g0, sh.gl,
                                                                norm_out(w, kNormOuter, 0.0, 0));
    AE_CUDA(cudaGetLastError());
    if (w->comm) AE_TRY(finish_sharded_norm(s, w, nullptr, kNormOuter, 0.0, 0, 0, st));
    return AE_OK;
}

int launch_ctrl_init(ae_work *w, cudaStream_t st)
{
    ctrl_init_kernel<<<1, 1, 0, st>>>(w->ctrl);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

}  // namespace ae

// +512: the fused small-slab kernels keep two (num, den) pairs per CTA of a grid of up to ~600 CTAs there
extern "C" int64_t ae_norm_blocks(int64_t Zp) { return ae::per_cell_blocks(Zp) + 512; }
