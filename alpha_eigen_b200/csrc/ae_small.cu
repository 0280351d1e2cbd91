// Fused solvers for small slabs (sm_100a, cooperative launch).
//
// The Kornreich-Parsons problems of the reference (450-1800 cells, 4-196 angles, one group) and the members
// of a parameter ensemble are far too small to fill a B200: run through the streaming kernels they cost two
// launches and a host poll per source iteration, ~20 us for a few microseconds of arithmetic.  Here whole
// loops run in ONE kernel, one CTA per row block (8 or 16 angles of one group and direction):
//   small_source_iteration_kernel  OneGroup.source_iteration_for_scalar_flux (group.py:62-81) /
//                                  TimeUnit.time_dependent_source_iteration (time_unit.py:44-59)
//   small_time_march_kernel        TimeUnit.calculate_time_dependent_flux (time_unit.py:17-42, repairs R2-R5):
//                                  every step, its outer (fission) and inner (scatter) loops, the committing
//                                  sweep and the snapshot writes
// Every warp sweeps its angle tile by tile with the same tile arithmetic as the streaming kernel (ae_tile.cuh);
// per-cell work (flux assembly, norms, scatter and fission sources) is spread over all threads with a FIXED
// thread-to-cell map, so consecutive per-cell phases need no barrier; grid-wide barriers separate sweeps from
// per-cell phases; convergence decisions are taken redundantly by every CTA from the same per-block sums in a
// fixed order.  The host reads the control block once, at the end.  Inputs stay in L2 (a few MB), so plain
// loads suffice; results are bit-compatible with the streaming path (same maps, same angle order in the flux
// sum), which the parity tests exercise through both.
//
// Ensembles (BASELINE configs[4]: 1024 parameter-varied 12-group slabs).  One member is a couple of dozen CTAs,
// so ensemble_time_march_kernel runs MANY members in one cooperative launch: the grid is cut into teams of
// `rows` CTAs, a team marches one member at a time exactly as small_time_march_kernel would (same code, the
// grid barrier replaced by a barrier over the team's CTAs only), and pulls the next member from a device-side
// queue when it is done.  Members keep their own convergence histories, so results equal stand-alone runs.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "ae_common.cuh"
#include "ae_tile.cuh"

namespace cg = cooperative_groups;

// Groups of a cell whose loads are in flight together in the flux update.  Measured on the 12-group ensemble march
// (72 members, one B200): 4 -> 1.31 s, 8 -> 2.73 s, 16 -> 1.73 s (wider chunks spill and waste predicated work); an
// L1 prefetch of the next tile during the sweeps cost another 9 %, and a cp.async double-buffered shared-memory ring
// for the tiles changed nothing on the one-group problems and cost 6 % on the ensemble
// (profiles/r01_small_kernel_cp_async_staging_ab.log): tile-load latency is not what limits these sweeps, so the
// tiles are simply loaded when needed.  (One-group problems would prefer a width of 1 by 5-10 %.)
#ifndef AE_SMALL_KU
#define AE_SMALL_KU 4
#endif

namespace ae {

struct SmallArgs {
    const double *hs, *cdt, *mu_ihx, *wgt, *scat, *chi, *nusf, *s_ext, *inflow;
    const int32_t *mat;
    double *base, *q0, *q1, *phi0, *phi1, *phi_outer, *partial, *scratch;
    ae_ctrl *ctrl;
    int64_t Z, Zp;
    int A, n_neg, G, nb_neg, nbg;
    // source iteration
    const double *psi_in;
    int max_iterations;
    double cold, tol;
    // time march
    double *psi, *psi_snap, *phi_snap;
    int32_t *outer_hist, *inner_hist;
    int num_steps;
};

// Barrier over the CTAs of one team (all co-resident: cooperative launch).  `gen` counts the barriers passed.
struct alignas(128) TeamBar {
    unsigned count, gen;
    int member;                                         // member the team is working on (written by its first CTA)
};
struct alignas(128) EnsQueue {
    unsigned next;                                      // next member to hand out
};
struct EnsResult {
    int32_t fail, steps_done, inner_iters, pad;         // fail: 0 ok, 1 inner / 2 outer loop did not converge
    double change, outer_change;
};

__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Arrival is ONE acq_rel atomic (it publishes this CTA's writes -- bar.sync made them thread 0's -- and, for the last
// arriver, acquires everybody else's); waiters poll with RELAXED loads and fence once when the generation has
// moved.  Polling with ld.acquire costs an L1 invalidation (CCTL.IVALL) per poll -- 570 per wait in the first
// version, which also emptied the L1 of the other teams' CTAs on the SM (march 7 % slower).
__device__ __forceinline__ void team_barrier(TeamBar *bar, unsigned &gen, int nblocks)
{
    ++gen;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atom_add_acq_rel(&bar->count, 1u) == (unsigned)nblocks - 1) {
            bar->count = 0;                             // everybody has arrived; nobody re-arrives before gen moves
            st_release(&bar->gen, gen);
        } else {
            while (ld_relaxed(&bar->gen) != gen) __nanosleep(32);
            __threadfence();                            // acquire side of the release store above
        }
    }
    __syncthreads();
}

template <int AC>
struct Small {
    const SmallArgs &p;
    cg::grid_group grid;
    TeamBar *bar;                                       // NULL: the team is the whole grid (grid.sync())
    unsigned bar_gen;
    int bid, nblocks;                                   // this CTA's row block, CTAs working on the problem
    double *red, *sm;
    int tid, warp, lane;
    int64_t nt, n, gthreads, gtid;
    // this CTA's row block
    int g, blk, nblk;
    bool neg;
    double m, two_m, hw, bc;                            // bc: flux entering this warp's row (0 = vacuum)
    int64_t goff, prow;
    double *part;

    __device__ Small(const SmallArgs &p_, double *red_, double *sm_, int bid_, int nblocks_, TeamBar *bar_ = nullptr,
                     unsigned gen_ = 0)
        : p(p_), grid(cg::this_grid()), bar(bar_), bar_gen(gen_), bid(bid_), nblocks(nblocks_), red(red_), sm(sm_)
    {
        tid = threadIdx.x; warp = tid >> 5; lane = tid & 31;
        nt = p.Zp / kTile; n = (int64_t)p.G * p.Zp;
        gthreads = (int64_t)nblocks * blockDim.x; gtid = (int64_t)bid * blockDim.x + tid;
        const int rb = bid;
        g = rb / p.nbg; blk = rb % p.nbg;
        neg = blk < p.nb_neg;
        const int a0 = neg ? blk * AC : p.n_neg + (blk - p.nb_neg) * AC;
        nblk = min(AC, (neg ? p.n_neg : p.A) - a0);
        const int a = a0 + (warp < nblk ? warp : 0);
        m = fabs(p.mu_ihx[a]); two_m = 2.0 * m;
        hw = warp < nblk ? 0.5 * p.wgt[a] : 0.0;
        bc = (p.inflow && warp < nblk) ? p.inflow[(int64_t)g * p.A + a] : 0.0;     // one_direction.py:12
        goff = (int64_t)g * p.Zp + lane * 2;
        prow = ((int64_t)g * p.A + a) * p.Zp + lane * 2;
        part = p.partial + ((int64_t)blk * p.G + g) * p.Zp;
    }

    __device__ void sync()
    {
        if (bar) team_barrier(bar, bar_gen, nblocks);
        else grid.sync();
    }

    // One sweep of this CTA's row block over the whole slab (group.py:41-60 for each of its angles).
    template <bool RD, bool WR>
    __device__ void sweep(const double *q, const double *psi_in, double *psi_out)
    {
        const double *qc = q + goff;
        double carry = bc;                              // vacuum boundary (group.py:48,54) unless a boundary_condition is set
        for (int64_t ti = 0; ti < nt; ++ti) {
            const int64_t off = (neg ? nt - 1 - ti : ti) * kTile;
            double ca[kL], cb[kL], f[kL];
            tile_maps<RD>(p.hs + goff + off, qc + off, RD ? p.cdt + goff + off : nullptr, RD ? psi_in + prow + off : nullptr, m,
                          two_m, ca, cb);
            if (neg) tile_solve<true>(ca, cb, carry, f, lane);
            else tile_solve<false>(ca, cb, carry, f, lane);
            if (off + kTile > p.Z) {
#pragma unroll
                for (int jj = 0; jj < kL; ++jj)
                    if (off + lane * kL + jj >= p.Z) f[jj] = 0.0;
            }
            if (WR && warp < nblk) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *reinterpret_cast<double2 *>(psi_out + prow + off + k * 64) = make_double2(0.5 * f[2 * k], 0.5 * f[2 * k + 1]);
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
        __syncthreads();                                // red is reused by the next sweep
    }

    // Flux update over the cells owned by this thread; it == 0 is the cold start (group.py:66 / time_unit.py:45).
    // Groups are handled kU at a time: the loads of a chunk are issued together and only then consumed (every
    // operand sits in L2, ~1 us away, and one group after the other would expose that latency G times); sums
    // keep the reference's order (partials in block order, norms in group order, scatter over g' ascending).
    __device__ void flux_update(int it, double cold, double *const (&q)[2], double *const (&phi)[2], double &num, double &den)
    {
        const double *phi_in = phi[(it + 1) & 1];
        double *phi_out = phi[it & 1], *q_next = q[it & 1];
        num = 0.0; den = 0.0;
        constexpr int kLocalG = 16;                     // a cell's fluxes stay in registers / L1-resident local memory
        constexpr int kU = AE_SMALL_KU;                 // groups in flight together
        const int G = p.G;
        for (int64_t s = gtid; s < p.Zp; s += gthreads) {
            const bool valid = logical_cell(s) < p.Z;
            double xv[kLocalG];
            for (int g0 = 0; g0 < G; g0 += kU) {
                double x[kU], old[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    x[u] = 0.0; old[u] = 0.0;
                    if (g0 + u < G && it != 0) {
                        const int64_t i = (int64_t)(g0 + u) * p.Zp + s;
                        x[u] = p.partial[i];
                        old[u] = phi_in[i];
                    }
                }
                for (int pp = 1; pp < p.nbg; ++pp) {                                        // block order (group.py:77)
#pragma unroll
                    for (int u = 0; u < kU; ++u)
                        if (g0 + u < G && it != 0) x[u] += p.partial[(int64_t)pp * n + (int64_t)(g0 + u) * p.Zp + s];
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    if (g0 + u >= G) continue;
                    if (it == 0) {
                        x[u] = valid ? cold : 0.0;
                    } else {
                        const double d = x[u] - old[u];                                     // group.py:78
                        num = fma(d, d, num);
                        den = fma(x[u], x[u], den);
                    }
                    phi_out[(int64_t)(g0 + u) * p.Zp + s] = x[u];                           // group.py:80
                    if (g0 + u < kLocalG) xv[g0 + u] = x[u];
                }
            }
            const double *sc = p.scat + (size_t)p.mat[s] * G * G;
            for (int g0 = 0; g0 < G; g0 += kU) {
                double acc[kU], b[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    acc[u] = 0.0;
                    b[u] = g0 + u < G ? p.base[(int64_t)(g0 + u) * p.Zp + s] : 0.0;         // in flight during the sums
                }
                for (int gp = 0; gp < G; ++gp) {
                    const double xg = G <= kLocalG ? xv[gp] : phi_out[(int64_t)gp * p.Zp + s];
#pragma unroll
                    for (int u = 0; u < kU; ++u)
                        if (g0 + u < G) acc[u] = fma(xg, __ldg(sc + gp * G + g0 + u), acc[u]);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    if (g0 + u < G) q_next[(int64_t)(g0 + u) * p.Zp + s] = b[u] + 0.5 * acc[u];   // group.py:73 / time_unit.py:53
            }
        }
    }

    // Grid-wide sqrt(sum num)/sqrt(sum den): per-block sums to scratch, barrier, every CTA adds them in block order.
    // Two scratch halves alternate so that a fast CTA entering the next reduction cannot overwrite sums a slow
    // CTA is still reading (every reduction contains one grid barrier).
    int ratio_calls = 0;
    __device__ double grid_ratio(double num, double den)
    {
        double *scr = p.scratch + (ratio_calls++ & 1) * 2 * nblocks;
        block_sum2(num, den, sm);
        if (tid == 0) { scr[2 * bid] = num; scr[2 * bid + 1] = den; }
        sync();
        double a = 0.0, b = 0.0;
        for (int i = tid; i < nblocks; i += blockDim.x) {
            a += __ldcg(scr + 2 * i);
            b += __ldcg(scr + 2 * i + 1);
        }
        block_sum2(a, b, sm);
        __syncthreads();
        if (tid == 0) { sm[0] = a; sm[1] = b; }
        __syncthreads();
        a = sm[0]; b = sm[1];
        __syncthreads();
        return sqrt(a) / sqrt(b);
    }

    // The inner (scatter) iteration.  Iteration n leaves the flux in phi[n & 1] and was fed by q[(n-1) & 1].
    template <bool RD>
    __device__ bool source_iteration(const double *psi_in, double cold, double tol, int max_iterations, double *const (&q)[2],
                                     double *const (&phi)[2], int &it, double &change)
    {
        double num, den;
        flux_update(0, cold, q, phi, num, den);
        sync();
        it = 0; change = 0.0;
        while (it < max_iterations - 1) {               // `assert iteration < max_iterations`, group.py:72
            ++it;
            sweep<RD, false>(q[(it - 1) & 1], psi_in, nullptr);
            sync();
            flux_update(it, cold, q, phi, num, den);
            change = grid_ratio(num, den);              // group.py:78
            if (change < tol) return true;              // group.py:79
        }
        return false;
    }
};

template <int AC, bool READ_PSI>
__global__ void __launch_bounds__(AC * 32, AC == 8 ? 2 : 1) small_source_iteration_kernel(const SmallArgs p)
{
    extern __shared__ double red[];                     // [2][AC][kTile]
    __shared__ double sm[64];
    Small<AC> S(p, red, sm, blockIdx.x, gridDim.x);
    double *const q[2] = {p.q0, p.q1};
    double *const phi[2] = {p.phi0, p.phi1};
    int it;
    double change;
    const bool ok = S.template source_iteration<READ_PSI>(p.psi_in, p.cold, p.tol, p.max_iterations, q, phi, it, change);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.ctrl->change = change;
        p.ctrl->inner_iters = it;
        p.ctrl->inner_done = ok ? 1 : 0;
        p.ctrl->ticket = 0;
    }
}

// TimeUnit.calculate_time_dependent_flux for the problem S works on (all CTAs of its team call this together).
// ctrl->reserved on exit: 0 ok, 1 inner loop did not converge, 2 outer loop did not converge; power_iters = steps done.
template <int AC>
__device__ void time_march(Small<AC> &S, const SmallArgs &p, EnsResult *res)
{
    double *const q[2] = {p.q0, p.q1};
    double *const phi[2] = {p.phi0, p.phi1};
    const int64_t n_psi = S.n * p.A;
    int fail = 0, steps_done = 0, it = 0;
    double change = 0.0, outer_change = 0.0;
    for (int step = 0; step < p.num_steps && !fail; ++step) {
        const double *psi_prev = p.psi_snap ? (step == 0 ? p.psi : p.psi_snap + (int64_t)(step - 1) * n_psi) : p.psi;
        double *psi_next = p.psi_snap ? p.psi_snap + (int64_t)step * n_psi : p.psi;
        for (int64_t s = S.gtid; s < p.Zp; s += S.gthreads)                     // time_unit.py:30 phi = zeros
            for (int gg = 0; gg < p.G; ++gg) p.phi_outer[(int64_t)gg * p.Zp + s] = 0.0;
        int outer = 0, inner_total = 0;
        for (;;) {
            ++outer;
            if (!(outer < 100)) { fail = 2; break; }                            // time_unit.py:35
            for (int64_t s = S.gtid; s < p.Zp; s += S.gthreads) {               // time_unit.py:37-39 / zone.py:38-42
                const int mi = p.mat[s];
                double acc = 0.0;
                for (int gp = 0; gp < p.G; ++gp) acc = fma(p.nusf[mi * p.G + gp], p.phi_outer[(int64_t)gp * p.Zp + s], acc);
                for (int gg = 0; gg < p.G; ++gg) {
                    const int64_t i = (int64_t)gg * p.Zp + s;
                    const double f = 0.5 * p.chi[mi * p.G + gg] * acc;
                    p.base[i] = p.s_ext ? p.s_ext[i] + f : f;
                }
            }
            if (!S.template source_iteration<true>(psi_prev, 0.0, p.tol, 100, q, phi, it, change)) { fail = 1; break; }   // :40
            inner_total += it;
            const double *cur = phi[it & 1];
            double num = 0.0, den = 0.0;
            for (int64_t s = S.gtid; s < p.Zp; s += S.gthreads)                 // time_unit.py:41, :36 (+R4)
                for (int gg = 0; gg < p.G; ++gg) {
                    const int64_t i = (int64_t)gg * p.Zp + s;
                    const double v = cur[i], d = v - p.phi_outer[i];
                    num = fma(d, d, num);
                    den = fma(v, v, den);
                    p.phi_outer[i] = v;
                }
            outer_change = S.grid_ratio(num, den);
            if (outer_change < p.tol) break;                                    // time_unit.py:42
        }
        if (S.bid == 0 && threadIdx.x == 0) { p.outer_hist[step] = outer; p.inner_hist[step] = inner_total; }
        if (fail) break;
        // commit: replay the last sweep with its own source, now writing psi (time_unit.py:26-27)
        S.template sweep<true, true>(q[(it - 1) & 1], psi_prev, psi_next);
        if (p.phi_snap) {
            const double *cur = phi[it & 1];
            double *dst = p.phi_snap + (int64_t)step * S.n;                     // time_unit.py:25
            for (int64_t s = S.gtid; s < p.Zp; s += S.gthreads)
                for (int gg = 0; gg < p.G; ++gg) dst[(int64_t)gg * p.Zp + s] = cur[(int64_t)gg * p.Zp + s];
        }
        steps_done = step + 1;
        S.sync();                                       // q / phi / partial are rewritten by the next step
    }
    if (S.bid == 0 && threadIdx.x == 0) {
        p.ctrl->change = change;
        p.ctrl->outer_change = outer_change;
        p.ctrl->inner_iters = it;                       // the flux of the last step is in phi[it & 1]
        p.ctrl->inner_done = fail ? 0 : 1;
        p.ctrl->power_iters = steps_done;
        p.ctrl->reserved = fail;
        p.ctrl->ticket = 0;
        if (res) {
            res->fail = fail; res->steps_done = steps_done; res->inner_iters = it; res->pad = 0;
            res->change = change; res->outer_change = outer_change;
        }
    }
}

template <int AC>
__global__ void __launch_bounds__(AC * 32, AC == 8 ? 2 : 1) small_time_march_kernel(const SmallArgs p)
{
    extern __shared__ double red[];
    __shared__ double sm[64];
    Small<AC> S(p, red, sm, blockIdx.x, gridDim.x);
    time_march<AC>(S, p, nullptr);
}

// Many members in one launch: teams of `rows` CTAs pull members from a queue (see the file header).
// MINB = CTAs per SM the register allocation aims at: members in flight = MINB * SMs / rows.  The march of a small
// member is latency bound (serial tile chains, team barriers), so more teams in flight win until spills bite.
template <int AC, int MINB>
__global__ void __launch_bounds__(AC * 32, MINB)
    ensemble_time_march_kernel(const SmallArgs *members, int count, int rows, EnsQueue *queue, TeamBar *bars, EnsResult *results)
{
    extern __shared__ double red[];
    __shared__ double sm[64];
    __shared__ SmallArgs sp;
    const int team = blockIdx.x / rows, bid = blockIdx.x - team * rows;
    TeamBar *bar = bars + team;
    unsigned gen = 0;
    for (;;) {
        if (bid == 0 && threadIdx.x == 0) bar->member = (int)atomicAdd(&queue->next, 1u);
        team_barrier(bar, gen, rows);
        const int m = *(volatile int *)&bar->member;
        if (m >= count) break;
        if (threadIdx.x == 0) sp = members[m];
        __syncthreads();
        {
            Small<AC> S(sp, red, sm, bid, rows, bar, gen);
            time_march<AC>(S, sp, results + m);
            gen = S.bar_gen;
        }
        team_barrier(bar, gen, rows);                   // everybody has read bar->member and sp, and left the member's buffers
    }
}

// One CTA per row block, all co-resident (cooperative launch).
template <class Kernel>
static int coop_launch(Kernel kernel, int ac, const SmallArgs &p, int rows, cudaStream_t st, bool &launched)
{
    const size_t smem = 2 * (size_t)ac * kTile * sizeof(double);
    launched = false;
    if (smem > 48 * 1024) AE_CUDA(cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    AE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ac * 32, smem));
    AE_CUDA(cudaGetDevice(&dev));
    AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm * sms < rows) return AE_OK;
    void *args[] = {(void *)&p};
    AE_CUDA(cudaLaunchCooperativeKernel((const void *)kernel, dim3(rows), dim3(ac * 32), args, smem, st));
    launched = true;
    return AE_OK;
}

int block_angles(int A, int n_neg);

// Small enough to be launch-latency bound through the streaming kernels?  Fills the problem part of SmallArgs.
static bool small_problem(const ae_slab *s, ae_work *w, SmallArgs &p, int &ac, int &rows)
{
    static const char *off = getenv("AE_NO_FUSED_SMALL");
    if ((off && off[0] == '1') || (w->flags & AE_WORK_NO_FUSED)) return false;
    if (w->comm) return false;                                          // multi-GPU goes through the exchange path
    const int64_t updates = s->Zp * (int64_t)s->A * s->G;
    if (updates > (1 << 23) || s->Zp / kTile > 256) return false;       // big enough for the streaming kernels
    ac = block_angles(s->A, s->n_neg);
    p.hs = s->hs; p.cdt = s->cdt; p.mu_ihx = s->mu_ihx; p.wgt = s->wgt; p.scat = s->scat; p.chi = s->chi; p.nusf = s->nusf;
    p.s_ext = s->s_ext; p.inflow = s->inflow; p.mat = s->mat; p.base = w->base; p.q0 = w->q0; p.q1 = w->q1; p.phi0 = w->phi; p.phi1 = w->phi_alt;
    p.phi_outer = w->phi_outer; p.partial = w->partial; p.scratch = w->norm_scratch; p.ctrl = w->ctrl;
    p.Z = s->Z; p.Zp = s->Zp; p.A = s->A; p.n_neg = s->n_neg; p.G = s->G;
    p.nb_neg = (s->n_neg + ac - 1) / ac;
    p.nbg = p.nb_neg + (s->A - s->n_neg + ac - 1) / ac;
    p.psi_in = nullptr; p.max_iterations = 0; p.cold = 0.0; p.tol = 0.0;
    p.psi = nullptr; p.psi_snap = nullptr; p.phi_snap = nullptr; p.outer_hist = nullptr; p.inner_hist = nullptr; p.num_steps = 0;
    rows = s->G * p.nbg;
    return 4 * rows <= 4 * (int)ae_norm_blocks(s->Zp) + 2 * s->G;       // two halves of per-block sums must fit norm_scratch
}

// Returns AE_OK with *launched = 1 when the fused kernel was queued for this source iteration.
int launch_small_source_iteration(const ae_slab *s, ae_work *w, const double *psi_in, double cold, double tol,
                                  int max_iterations, int *launched, cudaStream_t st)
{
    *launched = 0;
    SmallArgs p;
    int ac = 0, rows = 0;
    if (!small_problem(s, w, p, ac, rows)) return AE_OK;
    p.psi_in = psi_in; p.max_iterations = max_iterations; p.cold = cold; p.tol = tol;
    bool ok = false;
    if (ac == 16) {
        if (psi_in) AE_TRY(coop_launch(small_source_iteration_kernel<16, true>, 16, p, rows, st, ok));
        else AE_TRY(coop_launch(small_source_iteration_kernel<16, false>, 16, p, rows, st, ok));
    } else {
        if (psi_in) AE_TRY(coop_launch(small_source_iteration_kernel<8, true>, 8, p, rows, st, ok));
        else AE_TRY(coop_launch(small_source_iteration_kernel<8, false>, 8, p, rows, st, ok));
    }
    *launched = ok ? 1 : 0;
    return AE_OK;
}

// Whole time march in one kernel.  outer_hist / inner_hist are DEVICE arrays of num_steps ints.
int launch_small_time_march(const ae_slab *s, ae_work *w, double *psi, double *psi_snap, double *phi_snap, int num_steps,
                            double tol, int32_t *outer_hist, int32_t *inner_hist, int *launched, cudaStream_t st)
{
    *launched = 0;
    SmallArgs p;
    int ac = 0, rows = 0;
    if (!small_problem(s, w, p, ac, rows) || !s->cdt) return AE_OK;
    p.tol = tol; p.psi = psi; p.psi_snap = psi_snap; p.phi_snap = phi_snap; p.outer_hist = outer_hist; p.inner_hist = inner_hist;
    p.num_steps = num_steps;
    bool ok = false;
    if (ac == 16) AE_TRY(coop_launch(small_time_march_kernel<16>, 16, p, rows, st, ok));
    else AE_TRY(coop_launch(small_time_march_kernel<8>, 8, p, rows, st, ok));
    *launched = ok ? 1 : 0;
    return AE_OK;
}

int check_work(const ae_slab *s, const ae_work *w, const char *who);

}  // namespace ae

// Time march of `count` same-shape small slabs in ONE cooperative launch (teams of CTAs, member queue).
extern "C" int ae_ensemble_time_march(const ae_member *members, int32_t count, int32_t num_steps, double tol,
                                      int32_t *status, int32_t *teams_out, void *stream)
{
    using namespace ae;
    cudaStream_t st = (cudaStream_t)stream;
    if (!members || count < 1 || num_steps < 1) { set_error("ae_ensemble_time_march: bad argument"); return AE_BAD_ARG; }
    SmallArgs *h_args = (SmallArgs *)malloc(sizeof(SmallArgs) * (size_t)count);
    EnsResult *h_res = (EnsResult *)malloc(sizeof(EnsResult) * (size_t)count);
    if (!h_args || !h_res) { free(h_args); free(h_res); set_error("ae_ensemble_time_march: out of host memory"); return AE_CUDA_ERROR; }
    struct Guard {
        SmallArgs *a; EnsResult *r; void *d; cudaStream_t st;
        ~Guard() { free(a); free(r); if (d) cudaFreeAsync(d, st); }
    } guard{h_args, h_res, nullptr, st};
    int ac0 = 0, rows0 = 0;
    const ae_slab *s0 = members[0].slab;
    for (int m = 0; m < count; ++m) {
        const ae_member &mem = members[m];
        AE_TRY(check_work(mem.slab, mem.work, "ae_ensemble_time_march"));
        int ac = 0, rows = 0;
        if (!mem.psi || !mem.hist || !mem.slab->cdt || mem.work->comm || !small_problem(mem.slab, mem.work, h_args[m], ac, rows)) {
            set_error("ae_ensemble_time_march: member %d is not a single-GPU small time-dependent slab", m);
            return AE_BAD_ARG;
        }
        if (m == 0) { ac0 = ac; rows0 = rows; }
        if (mem.slab->Z != s0->Z || mem.slab->Zp != s0->Zp || mem.slab->A != s0->A || mem.slab->n_neg != s0->n_neg ||
            mem.slab->G != s0->G) {
            set_error("ae_ensemble_time_march: member %d has a different shape (Z, A, n_neg, G must match)", m);
            return AE_BAD_ARG;
        }
        SmallArgs &p = h_args[m];
        p.tol = tol; p.psi = mem.psi; p.psi_snap = mem.psi_snap; p.phi_snap = mem.phi_snap;
        p.outer_hist = mem.hist; p.inner_hist = mem.hist + num_steps; p.num_steps = num_steps;
        AE_CUDA(cudaMemsetAsync(mem.hist, 0, sizeof(int32_t) * 2 * (size_t)num_steps, st));
    }
    // 8-warp CTAs: 2 (128 registers), 3 (80, spills) or 4 (64, more spills) CTAs per SM; measured equal within 5 % on
    // the 12-group ensemble (profiles/r01_c5_ensemble.jsonl), so the spill-free build is the default;
    // AE_ENSEMBLE_CTAS_PER_SM overrides
    static const char *force = getenv("AE_ENSEMBLE_CTAS_PER_SM");
    const int minb = force && atoi(force) >= 2 && atoi(force) <= 4 ? atoi(force) : 2;
    const void *kernel = ac0 == 16 ? (const void *)ensemble_time_march_kernel<16, 1>
                         : minb == 2 ? (const void *)ensemble_time_march_kernel<8, 2>
                         : minb == 3 ? (const void *)ensemble_time_march_kernel<8, 3>
                                     : (const void *)ensemble_time_march_kernel<8, 4>;
    const size_t smem = 2 * (size_t)ac0 * kTile * sizeof(double);
    if (smem > 48 * 1024) AE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    AE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ac0 * 32, smem));
    AE_CUDA(cudaGetDevice(&dev));
    AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int teams = per_sm * sms / rows0;
    if (teams > count) teams = count;
    if (teams < 1) { set_error("ae_ensemble_time_march: a member needs %d co-resident CTAs, the device holds %d", rows0, per_sm * sms); return AE_BAD_ARG; }
    if (teams_out) *teams_out = teams;
    // one device block: [SmallArgs x count][EnsResult x count][EnsQueue][TeamBar x teams]
    const size_t off_res = (sizeof(SmallArgs) * (size_t)count + 127) / 128 * 128;
    const size_t off_q = off_res + (sizeof(EnsResult) * (size_t)count + 127) / 128 * 128;
    const size_t off_bar = off_q + sizeof(EnsQueue);
    const size_t bytes = off_bar + sizeof(TeamBar) * (size_t)teams;
    char *d = nullptr;
    AE_CUDA(cudaMallocAsync((void **)&d, bytes, st));
    guard.d = d;
    AE_CUDA(cudaMemsetAsync(d + off_res, 0, bytes - off_res, st));
    AE_CUDA(cudaMemcpyAsync(d, h_args, sizeof(SmallArgs) * (size_t)count, cudaMemcpyHostToDevice, st));
    const SmallArgs *d_args = (const SmallArgs *)d;
    EnsResult *d_res = (EnsResult *)(d + off_res);
    EnsQueue *d_q = (EnsQueue *)(d + off_q);
    TeamBar *d_bar = (TeamBar *)(d + off_bar);
    int cnt = count, rows = rows0;
    void *args[] = {(void *)&d_args, (void *)&cnt, (void *)&rows, (void *)&d_q, (void *)&d_bar, (void *)&d_res};
    AE_CUDA(cudaLaunchCooperativeKernel(kernel, dim3((unsigned)(teams * rows0)), dim3(ac0 * 32), args, smem, st));
    AE_CUDA(cudaMemcpyAsync(h_res, d_res, sizeof(EnsResult) * (size_t)count, cudaMemcpyDeviceToHost, st));
    AE_CUDA(cudaStreamSynchronize(st));
    int any_fail = 0;
    const size_t n_psi = (size_t)s0->G * s0->Zp * s0->A;
    for (int m = 0; m < count; ++m) {
        ae_work *w = members[m].work;
        if (h_res[m].inner_iters & 1) { double *t = w->phi; w->phi = w->phi_alt; w->phi_alt = t; }   // newest flux -> w->phi
        const int failed = h_res[m].fail != 0 || h_res[m].steps_done != num_steps;
        any_fail |= failed;
        if (status) status[m] = failed ? AE_NOT_CONVERGED : AE_OK;
        if (!failed && members[m].psi_snap)
            AE_CUDA(cudaMemcpyAsync(members[m].psi, members[m].psi_snap + (size_t)(num_steps - 1) * n_psi, n_psi * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
    }
    AE_CUDA(cudaStreamSynchronize(st));
    return (any_fail && !status) ? AE_NOT_CONVERGED : AE_OK;
}
