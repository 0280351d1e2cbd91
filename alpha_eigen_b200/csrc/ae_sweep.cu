// Batched diamond-difference slab sweep as an affine-recurrence scan (sm_100a, FP64).
//
// Replaces the per-cell Python loop of OneGroup.sweep1D (reference group.py:41-60) and the
// angle loop + quadrature accumulate around it (group.py:74-77, time_unit.py:52-55).
//
// Arithmetic.  One cell update  psi_out = (q + (m - s/2) psi_in) / (m + s/2),  m = |mu|/hx,
// s = sigT, is the affine map psi_out = a psi_in + b with a = 2m/(m + s/2) - 1 and
// b = q/(m + s/2).  Maps compose associatively, (a2,b2)o(a1,b1) = (a2 a1, a2 b1 + b2), so a row
// (one angle of one group) is solved by: each lane folds its 8 consecutive cells serially (the
// reference's order inside a lane), a 5-step warp-shuffle scan composes the 32 lane maps, and a
// running carry links consecutive 256-cell tiles.
//
// Mapping.  A "row block" is up to AC (8 or 16) angles of one group and one direction; its rows are the
// AC compute warps of a CTA, which therefore share the per-(cell, group) inputs hs, q, cdt; a reducer
// warp sums their weighted midpoint fluxes once per tile (angle order, deterministic) into a
// per-angle-block partial scalar flux, and a producer warp feeds the shared-memory ring.  The
// (row block, tile) pairs of the whole sweep set form one sequence that is cut into equal contiguous
// ranges, one per persistent CTA (grid = resident CTAs), so every SM gets the same number of tiles
// whatever A, G and the rank count are.  A range that starts / ends inside a row block hands the
// running carries to its neighbour through a small global mailbox; the piece that PRODUCES a hand-off is
// processed first and the piece that CONSUMES one last, so the wait is satisfied long before it is
// reached whenever a CTA owns at least one row block's worth of tiles.
//
// Memory.  Cells are stored in the lane-blocked tile order of alpha_eigen_b200.h, so a tile of any row is
// one contiguous, 2 KB, 16-byte aligned run.  Tiles are staged by the TMA engine (cp.async.bulk +
// mbarrier complete_tx) into a 3- or 4-stage shared-memory ring: the lanes of the producer warp issue
// the (3 + AC) bulk copies of a tile (hs, q, cdt once per CTA, psi per compute warp) as soon as every
// compute warp has pulled the previous occupant of the stage into registers, so loads never sit on a
// compute warp's scoreboard.  psi is written back with 16-byte streaming stores (in place is safe: a
// tile is read before it is written and prefetches only run ahead).  Algorithmic bytes per
// cell-angle-group update: 8 (psi read) + 8 (psi write) + 24/A (hs, q, cdt over the rank's angles)
// + 8/AC (partial flux write).
#include <stdlib.h>
#include "ae_common.cuh"
#include "ae_tile.cuh"

#ifndef AE_SWEEP_REDUCERS
#define AE_SWEEP_REDUCERS 2      // reducer warps of the 16-angle CTA (1, 2 or 4)
#endif

namespace ae {

constexpr int kACsmall = 8, kACbig = 16; // compute warps (angles) per CTA: small / many-angle problems
constexpr int kTileBytes = kTile * 8;   // 2 KB

struct SweepArgs {
    const double *hs, *cdt, *q, *mu_ihx, *wgt, *psi_in;
    double *psi_out, *partial;
    SweepSync *sync;
    int64_t Z, Zp;
    int64_t total_tiles;                // units * nt (units = G * nbg / gs row-block gangs)
    int A, n_neg, G, nb_neg, nbg;       // nbg = angle blocks per group
    int gs, subs;                       // gang size (CTAs that walk the same cells together), nb_dir / gs
    int l2_hints;                       // keep hs/q/cdt tiles in L2 (evict_last), stream psi through it (evict_first)
};

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA engine) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint
// expires) instead of burning issue slots in a spin loop next to the compute warps.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// L2 eviction policies: tiles several CTAs read (hs, q, cdt) are kept, psi is used exactly once.
__device__ __forceinline__ uint64_t l2_policy(int kind)      // 0 normal, 1 evict_first, 2 evict_last
{
    uint64_t pol;
    if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void st_stream(double *p, double2 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol)
                 : "memory");
}
// Walks the tiles a CTA processes.  Its range [t0, t1) of the flattened (row block, tile) sequence
// is visited as: last (head) piece, whole row blocks, first (tail) piece -- see the file header.
struct Cursor {
    int nt, rb, ti;                      // tiles per row; current row block and tile (sweep order)
    int b0, b1, first_ti;                // first / last row block of the range, tile where the range starts
    int64_t j, len_last;                 // tiles done so far; length of the head piece
    __host__ __device__ void init(int64_t t0, int64_t t1, int nt_)
    {
        nt = nt_;
        b0 = (int)(t0 / nt); b1 = (int)((t1 - 1) / nt);
        first_ti = (int)(t0 - (int64_t)b0 * nt);
        j = 0;
        if (b0 == b1) { len_last = t1 - t0; rb = b0; ti = first_ti; }
        else { len_last = t1 - (int64_t)b1 * nt; rb = b1; ti = 0; }
    }
    __host__ __device__ void advance()
    {
        ++j; ++ti;
        if (b0 == b1) return;
        if (j == len_last) {                        // head piece done
            if (b1 - b0 > 1) { rb = b0 + 1; ti = 0; } else { rb = b0; ti = first_ti; }
        } else if (ti == nt && j > len_last) {      // a whole row block done
            ++rb; ti = 0;
            if (rb == b1) { rb = b0; ti = first_ti; }
        }
    }
};

struct RowCtx {                          // per (row block, warp) constants
    double m, two_m, hw;
    int64_t row;                         // element offset of the block's first psi row
    int64_t poff;                        // element offset of the partial row
    int g, nblk;                         // group, angles present in this block (<= AC)
    bool neg;
};

template <int AC>
__device__ __forceinline__ void row_setup(const SweepArgs &p, int rb, int warp, RowCtx &c)
{
    c.g = rb / p.nbg;
    const int blk = rb % p.nbg;
    c.neg = blk < p.nb_neg;
    const int a0 = c.neg ? blk * AC : p.n_neg + (blk - p.nb_neg) * AC;
    const int a_end = c.neg ? p.n_neg : p.A;
    c.nblk = min(AC, a_end - a0);
    const int a = a0 + (warp < c.nblk ? warp : 0);
    c.m = fabs(p.mu_ihx[a]);
    c.two_m = 2.0 * c.m;
    c.hw = warp < c.nblk ? 0.5 * p.wgt[a] : 0.0;
    c.row = ((int64_t)c.g * p.A + a0) * p.Zp;
    c.poff = ((int64_t)blk * p.G + c.g) * p.Zp;
}

// Work unit u of gang member mb -> row block.  Without gangs (gs == 1) a unit IS a row block.  With gangs a unit is
// gs angle blocks of one group and direction: its members walk the same cells at the same time, so hs/q/cdt
// come from HBM once per unit and from L2 for the other gs - 1 members.
__host__ __device__ __forceinline__ int unit_row(const SweepArgs &p, int u, int mb)
{
    if (p.gs == 1) return u;
    const int upg = 2 * p.subs;                     // units per group (both directions have nb_neg blocks)
    const int g = u / upg, r = u - g * upg;
    const int dir = r / p.subs, sub = r - dir * p.subs;
    return g * p.nbg + dir * p.nb_neg + sub * p.gs + mb;
}

template <int AC> struct StageCfg {
    static constexpr int stages = AC == kACbig ? 4 : 3;
    // Reducer warps, each summing its share of a tile's cells.  With one reducer the compute warps of the 16-angle
    // CTA spent 21 % of their samples waiting for the reduction array to be recycled (ncu, C3); a second one
    // halves that latency and still fits the register file (19 warps x 96 registers).  Measured (sweep kernel, ms,
    // profiles/r01_sweep_reducer_warps_ab.log): C3 1.414 / 1.284 / 1.228 and C4 62.0 / 54.3 / 56.8 with 1 / 2 / 4
    // reducers -- C4 runs at the board's power cap, where the extra warps of the 4-reducer build cost more than
    // they hide.  Two 8-angle CTAs per SM have no room for another warp.
    static constexpr int reducers = AC == kACbig ? AE_SWEEP_REDUCERS : 1;
    static constexpr int threads = (AC + reducers + 1) * 32;
};

// Warp-specialised persistent kernel: AC compute warps + reducer warp(s) + producer warp.
// Compute warp w owns one angle of the current row block and meets no CTA-wide barrier inside the tile
// loop; the producer warp feeds the TMA ring; the reducer warp folds the per-warp weighted fluxes into
// the partial scalar flux.  Hand-shakes are mbarriers:
//   full[s]      producer -> compute   tile staged (expect_tx bytes landed)
//   empty[s]     compute -> producer   all AC warps hold stage s in registers (released after phase A)
//   red_full[b]  compute -> reducer    all AC warps have written their flux contribution into red[b]
//   red_empty[b] reducer -> compute    red[b] has been summed and may be overwritten
template <int AC, bool READ_PSI, bool WRITE_PSI>
__global__ void __launch_bounds__(StageCfg<AC>::threads, AC == kACsmall ? 2 : 1) sweep_kernel(const SweepArgs p, const ae_ctrl *ctrl, int gate)
{
    constexpr int kStages = StageCfg<AC>::stages;
    constexpr int kRed = StageCfg<AC>::reducers, kPerRed = 4 / kRed;   // a reducer sums kPerRed of the 4 double2 columns
    constexpr int kSlots = (READ_PSI ? 3 + AC : 2);             // 2 KB slots per stage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);                    // [kStages][kSlots][kTile]
    double *red = ring + kStages * kSlots * kTile;                          // [2][AC][kTile]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * AC * kTile);    // [kStages]
    uint64_t *empty = full + kStages;                                       // [kStages]
    uint64_t *red_full = empty + kStages;                                   // [2]
    uint64_t *red_empty = red_full + 2;                                     // [2]
    __shared__ int s_cta;
    if (gate && ctrl->inner_done) return;                                   // batch launched past convergence

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    SweepSync *sy = p.sync;
    if (tid == 0) {
        s_cta = (int)atomicInc(&sy->ticket, gridDim.x - 1);                // logical id in START order; self-resetting
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, AC); }
        for (int b = 0; b < 2; ++b) { mbar_init(red_full + b, AC); mbar_init(red_empty + b, kRed); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int cta = s_cta;
    const int gang = cta / p.gs, mb = cta - gang * p.gs;                    // members of a gang start back to back
    const unsigned stamp = sy->epoch + 1;                                   // epoch only advances after every CTA has left
    const int64_t nt = p.Zp / kTile;
    const int64_t T = p.total_tiles, NC = gridDim.x / p.gs;
    const int64_t t0 = gang * T / NC, t1 = (gang + 1) * T / NC, W = t1 - t0;
    const bool publishes = (t1 % nt) != 0;                                  // range ends inside a row block
    Cursor cur;
    cur.init(t0, t1, (int)nt);
    RowCtx c;
    int c_rb = -1;

    if (warp == AC + kRed) {
        // ================= producer warp: TMA ring =====================================================
        const uint64_t pol = l2_policy(p.l2_hints ? (lane < 3 ? 2 : 1) : 0);   // lanes 0-2 fetch hs/q/cdt, the rest psi
        for (int64_t j = 0; j < W; ++j, cur.advance()) {
            if (cur.rb != c_rb) { row_setup<AC>(p, unit_row(p, cur.rb, mb), 0, c); c_rb = cur.rb; }
            const int64_t off = (int64_t)(c.neg ? nt - 1 - cur.ti : cur.ti) * kTile;
            const int st = (int)(j % kStages);
            double *dst = ring + (size_t)st * kSlots * kTile;
            const int64_t goff = (int64_t)c.g * p.Zp + off;
            const int nslots = READ_PSI ? 3 + c.nblk : 2;
            mbar_wait(empty + st, (uint32_t)(((j / kStages) & 1) ^ 1));     // consumers released the stage
            if (lane == 0) mbar_expect_tx(full + st, (uint32_t)(nslots * kTileBytes));
            __syncwarp();
            if (lane < nslots) {
                const double *src = lane == 0 ? p.hs + goff : lane == 1 ? p.q + goff : lane == 2 ? p.cdt + goff
                                                                                   : p.psi_in + c.row + (int64_t)(lane - 3) * p.Zp + off;
                bulk_g2s(dst + lane * kTile, src, kTileBytes, full + st, pol);
            }
        }
    } else if (warp >= AC) {
        // ================= reducer warp(s): partial scalar flux of the block ===========================
        const int k0 = (warp - AC) * kPerRed;                               // this warp's double2 columns of a tile
        double hwv[AC];                                                     // half quadrature weights of the block's angles
        for (int64_t j = 0; j < W; ++j, cur.advance()) {
            if (cur.rb != c_rb) {
                const int row = unit_row(p, cur.rb, mb);
                row_setup<AC>(p, row, 0, c);
                c_rb = cur.rb;
                const int a0 = c.neg ? (row % p.nbg) * AC : p.n_neg + (row % p.nbg - p.nb_neg) * AC;
#pragma unroll
                for (int w = 0; w < AC; ++w) hwv[w] = w < c.nblk ? 0.5 * p.wgt[a0 + w] : 0.0;
            }
            const int64_t off = (int64_t)(c.neg ? nt - 1 - cur.ti : cur.ti) * kTile;
            const int b = (int)(j & 1);
            const double *rb_ = red + b * (AC * kTile) + lane * 2;
            mbar_wait(red_full + b, (uint32_t)((j >> 1) & 1));
            double2 acc[kPerRed];
#pragma unroll
            for (int k = 0; k < kPerRed; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(rb_ + (k0 + k) * 64);
                acc[k] = make_double2(hwv[0] * v.x, hwv[0] * v.y);
            }
#pragma unroll
            for (int w = 1; w < AC; ++w)                                    // warps = angles in order (group.py:77)
#pragma unroll
                for (int k = 0; k < kPerRed; ++k) {
                    if (w >= c.nblk) continue;          // idle warps of a ragged block worked on stale data (may be NaN)
                    const double2 v = *reinterpret_cast<const double2 *>(rb_ + w * kTile + (k0 + k) * 64);
                    acc[k].x = fma(hwv[w], v.x, acc[k].x);                  // 0.5*w_a*(in+out) = w_a*psi_mid
                    acc[k].y = fma(hwv[w], v.y, acc[k].y);
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(red_empty + b);
            double *dst = p.partial + c.poff + off + lane * 2;
#pragma unroll
            for (int k = 0; k < kPerRed; ++k) *reinterpret_cast<double2 *>(dst + (k0 + k) * 64) = acc[k];
        }
    } else {
        // ================= compute warps ==============================================================
        double carry = 0.0;
        int64_t j = 0;
        const uint64_t pol = l2_policy(p.l2_hints ? 1 : 0);
        for (int piece = 0; piece < 2; ++piece) {
            const int64_t jend = piece == 0 ? (cur.len_last < W ? cur.len_last : W) : W;
            for (; j < jend; ++j, cur.advance()) {
                const int rb = cur.rb, ti = cur.ti;
                if (rb != c_rb || ti == 0) {
                    if (rb != c_rb) { row_setup<AC>(p, unit_row(p, rb, mb), warp, c); c_rb = rb; }
                    carry = 0.0;                                            // vacuum boundary: group.py:48,54
                    if (ti != 0) {                                          // continue a row another CTA started
                        if (lane == 0 && warp < c.nblk) {
                            while (ld_acquire(&sy->flag[cta - p.gs][warp]) != stamp) {}    // same member of the previous gang
                            carry = sy->carry[cta - p.gs][warp];
                        }
                        carry = __shfl_sync(0xffffffffu, carry, 0);
                    }
                }
                const int st = (int)(j % kStages), b = (int)(j & 1);
                const int64_t off = (int64_t)(c.neg ? nt - 1 - ti : ti) * kTile;
                const double *stage = ring + (size_t)st * kSlots * kTile + lane * 2;
                double ca[kL], cb[kL], f[kL];
                mbar_wait(full + st, (uint32_t)((j / kStages) & 1));
                tile_maps<READ_PSI>(stage, stage + kTile, stage + 2 * kTile, stage + (3 + warp) * kTile, c.m, c.two_m, ca, cb);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + st);                     // stage is in registers: release it
                if (c.neg) tile_solve<true>(ca, cb, carry, f, lane);
                else tile_solve<false>(ca, cb, carry, f, lane);
                if (off + kTile > p.Z) {                                    // only the last tile holds padding cells:
#pragma unroll
                    for (int jj = 0; jj < kL; ++jj)                         // exact zeros there, in psi and in the flux
                        if (off + lane * kL + jj >= p.Z) f[jj] = 0.0;
                }
                if (WRITE_PSI && warp < c.nblk) {
                    double *pout = p.psi_out + c.row + (int64_t)warp * p.Zp + off + lane * 2;
#pragma unroll
                    for (int k = 0; k < 4; ++k) st_stream(pout + k * 64, make_double2(0.5 * f[2 * k], 0.5 * f[2 * k + 1]), pol);
                }
                double *rw = red + b * (AC * kTile) + warp * kTile + lane * 2;
                mbar_wait(red_empty + b, (uint32_t)(((j >> 1) & 1) ^ 1));
#pragma unroll
                for (int k = 0; k < 4; ++k)                                 // in+out; the reducer applies the weights
                    *reinterpret_cast<double2 *>(rw + k * 64) = make_double2(f[2 * k], f[2 * k + 1]);
                __syncwarp();
                if (lane == 0) mbar_arrive(red_full + b);
            }
            // after the head piece: hand the carries to the CTA that continues this row
            if (piece == 0 && publishes && lane == 0 && warp < c.nblk) {
                sy->carry[cta][warp] = carry;
                st_release(&sy->flag[cta][warp], stamp);
            }
        }
    }
    // last CTA out advances the epoch (stamps of this launch can never match the next one's)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicInc(&sy->done, gridDim.x - 1) == gridDim.x - 1) atomicAdd(&sy->epoch, 1u);
    }
}

// phi[g][s] = sum_p partial[p][g][s] (padding cells are already exact zeros in every partial).  Flat, low
// register count for high occupancy (see flux_assemble_kernel).  phi may alias partial[0].
__global__ void __launch_bounds__(256, 8) reduce_partials_kernel(const double *partial, int P, int64_t n, double *phi)
{
    for (int64_t i = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 2 * (int64_t)gridDim.x * blockDim.x) {
        double2 x = *reinterpret_cast<const double2 *>(partial + i);
#pragma unroll 4
        for (int pp = 1; pp < P; ++pp) {
            const double2 y = *reinterpret_cast<const double2 *>(partial + (int64_t)pp * n + i);
            x.x += y.x; x.y += y.y;
        }
        *reinterpret_cast<double2 *>(phi + i) = x;
    }
}

// Row-block shape.  AC = compute warps (angles) per CTA.  16-warp CTAs halve the per-update
// traffic of the shared inputs (hs, q, cdt: 24/AC B) and of the partial flux (8/AC B) when a direction has
// the angles to fill them; AE_ANGLES_PER_CTA (8 | 16) overrides for tuning.
int block_angles(int A, int n_neg)
{
    static const char *force = getenv("AE_ANGLES_PER_CTA");
    const int side = n_neg > A - n_neg ? n_neg : A - n_neg;
    int ac = side >= kACbig ? kACbig : kACsmall;
    if (force && atoi(force) == kACsmall) ac = kACsmall;
    if (force && atoi(force) == kACbig) ac = kACbig;
    return ac;
}

// Gang size for a grid of `grid` persistent CTAs: the divisor gs of the angle blocks per direction that minimises
// (bytes per update with hs/q/cdt fetched nbg/gs times per group) / (fraction of the grid a whole number of gangs
// fills).  C4 on one GPU (8 + 8 blocks of 16 angles, 148 CTAs): gs = 4 -> 37 gangs, shared inputs read 4x instead
// of 16x per group.  AE_SWEEP_GANG=<n> forces a size (1 = every CTA walks its own cells, the pre-gang order).
static int gang_size(const SweepArgs &p, int64_t grid, int64_t full_grid)
{
    static const char *force = getenv("AE_SWEEP_GANG");
    const int nbd = p.nb_neg;
    if (p.nbg != 2 * nbd || nbd < 2 || grid < full_grid) return 1;     // ragged directions / small problems: no gangs
    if (force && atoi(force) >= 1) {
        const int f = atoi(force);
        return (nbd % f == 0 && grid / f >= 1) ? f : 1;
    }
    int best = 1;
    double best_cost = 1e300;
    for (int gs = 1; gs <= nbd; ++gs) {
        if (nbd % gs) continue;
        const int64_t ng = grid / gs;
        if (ng < 1 || (int64_t)p.G * p.nbg / gs < ng) continue;     // a gang owns at least one unit's worth of tiles, so
                                                                     // the hand-off it waits for was produced long before
        const double bytes = 16.0 + 8.0 * p.nbg / p.A + 24.0 * (p.nbg / gs) / p.A;
        const double cost = bytes * (double)grid / (double)(ng * gs);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = gs; }
    }
    return best;
}

// Grid size and gang shape for `resident` co-resident CTAs; fills p.gs / p.subs / p.total_tiles.
static int64_t plan_grid(SweepArgs &p, int64_t resident)
{
    int64_t grid = resident;
    if (grid > kMaxCtas) grid = kMaxCtas;
    const int64_t full_grid = grid;
    // small problems: at least one CTA per row block (rows are the only parallelism), else >= 4 tiles per CTA
    const int64_t rows = (int64_t)p.G * p.nbg, quarter = (p.total_tiles + 3) / 4;
    const int64_t want = rows > quarter ? rows : quarter;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    p.gs = gang_size(p, grid, full_grid);
    p.subs = p.nb_neg / p.gs;
    if (p.gs > 1) {
        grid = grid / p.gs * p.gs;                                  // whole gangs only
        p.total_tiles = (int64_t)p.G * (p.nbg / p.gs) * (p.Zp / kTile);
    }
    return grid;
}

template <int AC, bool RD, bool WR>
static int launch_variant(const SweepArgs &p_in, const ae_ctrl *ctrl, int gate, cudaStream_t st)
{
    SweepArgs p = p_in;
    constexpr int slots = RD ? 3 + AC : 2;
    constexpr int threads = StageCfg<AC>::threads;
    const size_t smem = (size_t)(StageCfg<AC>::stages * slots + 2 * AC) * kTileBytes + (2 * StageCfg<AC>::stages + 4) * sizeof(uint64_t);
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        AE_CUDA(cudaFuncSetAttribute((const void *)sweep_kernel<AC, RD, WR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 0;
        AE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sweep_kernel<AC, RD, WR>, threads, smem));
        if (n < 1) { set_error("sweep kernel does not fit on an SM"); return AE_CUDA_ERROR; }
        ctas_per_sm = n;
    }
    int dev = 0, sms = 0;
    AE_CUDA(cudaGetDevice(&dev));
    AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t grid = plan_grid(p, (int64_t)sms * ctas_per_sm);
    sweep_kernel<AC, RD, WR><<<(unsigned)grid, threads, smem, st>>>(p, ctrl, gate);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

template <int AC>
static int launch_shape(const SweepArgs &p, const ae_ctrl *ctrl, int gate, cudaStream_t st)
{
    const bool rd = p.psi_in != nullptr, wr = p.psi_out != nullptr;
    if (rd && wr) return launch_variant<AC, true, true>(p, ctrl, gate, st);
    if (rd) return launch_variant<AC, true, false>(p, ctrl, gate, st);
    if (wr) return launch_variant<AC, false, true>(p, ctrl, gate, st);
    return launch_variant<AC, false, false>(p, ctrl, gate, st);
}

// Row-block shape of a problem: fills the integer fields of SweepArgs, returns the angles per CTA.
static int shape_args(int A, int n_neg, int G, int64_t Zp, SweepArgs &p)
{
    const int blk = block_angles(A, n_neg);
    p.Zp = Zp; p.A = A; p.n_neg = n_neg; p.G = G;
    p.nb_neg = (n_neg + blk - 1) / blk;
    p.nbg = p.nb_neg + (A - n_neg + blk - 1) / blk;
    p.total_tiles = (int64_t)G * p.nbg * (Zp / kTile);
    p.gs = 1; p.subs = p.nb_neg;                        // gangs are chosen per launch (grid size)
    return blk;
}

// internal launcher shared with the solver loops (gate: skip when ctrl->inner_done is set)
int launch_sweep(const ae_slab *s, const double *q, const double *psi_in, double *psi_out, double *partial,
                 void *sync, const ae_ctrl *ctrl, int gate, cudaStream_t st)
{
    if (!s || !q || !partial || !sync || s->Zp % kTile || s->A <= 0 || s->G <= 0) {
        set_error("ae_sweep_set: bad argument");
        return AE_BAD_ARG;
    }
    if (psi_in && !s->cdt) { set_error("ae_sweep_set: psi_in given but slab has no cdt"); return AE_BAD_ARG; }
    SweepArgs p;
    const int ac = shape_args(s->A, s->n_neg, s->G, s->Zp, p);
    p.hs = s->hs; p.cdt = s->cdt; p.q = q; p.mu_ihx = s->mu_ihx; p.wgt = s->wgt;
    p.psi_in = psi_in; p.psi_out = psi_out; p.partial = partial; p.sync = (SweepSync *)sync;
    p.Z = s->Z;
    static const char *hints = getenv("AE_SWEEP_L2_HINTS");
    p.l2_hints = !(hints && hints[0] == '0');
    return ac == kACbig ? launch_shape<kACbig>(p, ctrl, gate, st) : launch_shape<kACsmall>(p, ctrl, gate, st);
}

}  // namespace ae

extern "C" int32_t ae_partial_count(int32_t A, int32_t n_neg)
{
    const int blk = ae::block_angles(A, n_neg);
    return (n_neg + blk - 1) / blk + (A - n_neg + blk - 1) / blk;
}

// Host-side replay of the kernel's work assignment (same Cursor / unit_row / plan_grid code): the (row block,
// tile index in sweep order) pairs logical CTA `cta` visits, in order.  Test hook: no GPU needed.
extern "C" int64_t ae_sweep_plan(int32_t A, int32_t n_neg, int32_t G, int64_t Zp, int32_t resident_ctas, int32_t cta,
                                 int32_t *grid_out, int32_t *gang_out, int32_t *row_blocks, int32_t *tiles, int64_t cap)
{
    if (A <= 0 || n_neg < 0 || n_neg > A || G <= 0 || Zp <= 0 || Zp % ae::kTile || resident_ctas < 1) return -1;
    ae::SweepArgs p;
    ae::shape_args(A, n_neg, G, Zp, p);
    const int64_t grid = ae::plan_grid(p, resident_ctas);
    if (grid_out) *grid_out = (int32_t)grid;
    if (gang_out) *gang_out = p.gs;
    if (cta < 0 || cta >= grid) return 0;
    const int64_t nt = Zp / ae::kTile, NC = grid / p.gs;
    const int gang = cta / p.gs, mb = cta - gang * p.gs;
    const int64_t t0 = gang * p.total_tiles / NC, t1 = (gang + 1) * p.total_tiles / NC, W = t1 - t0;
    ae::Cursor cur;
    cur.init(t0, t1, (int)nt);
    for (int64_t j = 0; j < W; ++j, cur.advance()) {
        if (j < cap) {
            if (row_blocks) row_blocks[j] = ae::unit_row(p, cur.rb, mb);
            if (tiles) tiles[j] = cur.ti;
        }
    }
    return W;
}

extern "C" int64_t ae_sweep_sync_bytes(void) { return (int64_t)sizeof(ae::SweepSync); }

extern "C" int ae_sweep_set(const ae_slab *s, const double *q, const double *psi_in, double *psi_out,
                            double *partial, void *sync, void *stream)
{
    return ae::launch_sweep(s, q, psi_in, psi_out, partial, sync, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int ae_reduce_partials(const ae_slab *s, const double *partial, int32_t P, double *phi, void *stream)
{
    if (!s || !partial || !phi || P < 1) { ae::set_error("ae_reduce_partials: bad argument"); return AE_BAD_ARG; }
    const int64_t n = (int64_t)s->G * s->Zp;
    int64_t bx = (n / 2 + 255) / 256;
    if (bx > 148 * 32) bx = 148 * 32;
    ae::reduce_partials_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(partial, P, n, phi);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}
