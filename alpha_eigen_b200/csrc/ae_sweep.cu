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
    const double *tinfo;                // [G][Zp/kTile][2] uniform-tile records (ae_tile_info) or NULL: general path only
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
// A CTA's range [t0, t1) of the flattened (row block, tile) sequence is visited as: last (head) piece, whole row
// blocks, first (tail) piece -- see the file header -- as a short list of segments (row block, first tile in sweep
// order, tiles): what the kernel's tile loops iterate over.  Segment 0 is the piece whose end hands carries on; the last one may consume a hand-off.
struct Segments {
    int nt, b0, b1, first_ti, len_last, count;
    __host__ __device__ void init(int64_t t0, int64_t t1, int nt_)
    {
        nt = nt_;
        b0 = (int)(t0 / nt); b1 = (int)((t1 - 1) / nt);
        first_ti = (int)(t0 - (int64_t)b0 * nt);
        if (t1 <= t0) { count = 0; len_last = 0; }
        else if (b0 == b1) { len_last = (int)(t1 - t0); count = 1; }
        else { len_last = (int)(t1 - (int64_t)b1 * nt); count = b1 - b0 + 1; }
    }
    __host__ __device__ void get(int k, int &rb, int &ti0, int &n) const
    {
        if (count == 1) { rb = b0; ti0 = first_ti; n = len_last; }
        else if (k == 0) { rb = b1; ti0 = 0; n = len_last; }
        else if (k == count - 1) { rb = b0; ti0 = first_ti; n = nt - first_ti; }
        else { rb = b0 + k; ti0 = 0; n = nt; }
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

#ifndef AE_SWEEP_PHI_STAGES
#define AE_SWEEP_PHI_STAGES 12
#endif
template <int AC, bool READ_PSI = true> struct StageCfg {
    // phi-only sweeps stage one 2 KB tile (q) per step, so a deep ring costs little and keeps enough tiles in flight
    // to cover the TMA latency at their much higher tile rate
    static constexpr int stages = READ_PSI ? (AC == kACbig ? 4 : 3) : AE_SWEEP_PHI_STAGES;
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
    constexpr int kStages = StageCfg<AC, READ_PSI>::stages;
    constexpr int kRed = StageCfg<AC>::reducers, kPerRed = 4 / kRed;   // a reducer sums kPerRed of the 4 double2 columns
    constexpr int kSlots = (READ_PSI ? 3 + AC : 2);             // 2 KB slots per stage
    constexpr int kStageLen = kSlots * kTile + 2;               // doubles per stage: the slots + the tile's record
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);                    // [kStages][kStageLen]
    double *red = ring + kStages * kStageLen;                               // [2][AC][kTile]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * AC * kTile);    // [kStages]
    uint64_t *empty = full + kStages;                                       // [kStages]
    uint64_t *red_full = empty + kStages;                                   // [2]
    uint64_t *red_empty = red_full + 2;                                     // [2]
    __shared__ int s_cta;
    __shared__ __align__(16) double s_ap[AC][kL];                           // uniform-tile constants, per compute warp
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
    const int nt = (int)(p.Zp / kTile);
    const int64_t T = p.total_tiles, NC = gridDim.x / p.gs;
    const int64_t t0 = gang * T / NC, t1 = (gang + 1) * T / NC;
    const bool publishes = (t1 % nt) != 0;                                  // range ends inside a row block
    Segments segs;
    segs.init(t0, t1, nt);
    const bool info = p.tinfo != nullptr;
    RowCtx c;
    // Every role walks the same segments (pieces of rows) in the same order; inside a segment the tile loop is
    // plain counting: ring stage `st` with phase `ph`, reduction buffer `b` with phase `bph`.
    int st = 0, b = 0;
    uint32_t ph = 0, bph = 0;

    if (warp == AC + kRed) {
        // ================= producer warp: TMA ring =====================================================
        const uint64_t pol = l2_policy(p.l2_hints ? (lane < 3 ? 2 : 1) : 0);   // lanes 0-2 fetch hs/q/cdt, the rest psi
        for (int k = 0; k < segs.count; ++k) {
            int rb, ti0, n;
            segs.get(k, rb, ti0, n);
            row_setup<AC>(p, unit_row(p, rb, mb), 0, c);
            const int dir = c.neg ? -1 : 1;
            int tile = c.neg ? nt - 1 - ti0 : ti0;                          // storage tile index, walks along the sweep
            // The record of a tile ({hs, cdt} when all its cells share them, NaN otherwise) tells which copies it
            // needs.  The records of a segment are contiguous: each lane fetches one, 32 tiles per refill and one
            // refill ahead, and they are handed round by shuffle (a load per tile in this loop would put its whole
            // L2 latency on every tile whenever the ring is not full).
            const double *recs = info ? p.tinfo + 2 * ((int64_t)c.g * nt + tile) : nullptr;
            double rec_cur = 0.0, rec_nxt = 0.0;
            if (info && lane < n) rec_nxt = __ldg(recs + 2 * dir * lane);
            const int nslots = READ_PSI ? 3 + c.nblk : 2;
            for (int i = 0; i < n; ++i, tile += dir) {
                if ((i & 31) == 0) {
                    rec_cur = rec_nxt;
                    if (info && i + 32 + lane < n) rec_nxt = __ldg(recs + 2 * dir * (i + 32 + lane));
                }
                const double rec_h = __shfl_sync(0xffffffffu, rec_cur, i & 31);
                const bool uniform = info && rec_h == rec_h;
                const int64_t off = (int64_t)tile * kTile;
                double *dst = ring + (size_t)st * kStageLen;
                const int64_t goff = (int64_t)c.g * p.Zp + off;
                // a uniform tile needs neither the hs nor the cdt tile
                const int ncopies = uniform ? nslots - (READ_PSI ? 2 : 1) : nslots;
                mbar_wait(empty + st, ph ^ 1);                              // consumers released the stage
                if (lane == 0) mbar_expect_tx(full + st, (uint32_t)(ncopies * kTileBytes + (info ? 16 : 0)));
                __syncwarp();
                if (lane < nslots && !(uniform && (lane == 0 || lane == 2))) {
                    const double *src = lane == 0 ? p.hs + goff : lane == 1 ? p.q + goff : lane == 2 ? p.cdt + goff
                                                                                       : p.psi_in + c.row + (int64_t)(lane - 3) * p.Zp + off;
                    bulk_g2s(dst + lane * kTile, src, kTileBytes, full + st, pol);
                }
                if (info && lane == 31) bulk_g2s(dst + kSlots * kTile, recs + 2 * dir * i, 16, full + st, pol);
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp >= AC) {
        // ================= reducer warp(s): partial scalar flux of the block ===========================
        const int k0 = (warp - AC) * kPerRed;                               // this warp's double2 columns of a tile
        double hwv[AC];                                                     // half quadrature weights of the block's angles
        for (int k = 0; k < segs.count; ++k) {
            int rb, ti0, n;
            segs.get(k, rb, ti0, n);
            const int row = unit_row(p, rb, mb);
            row_setup<AC>(p, row, 0, c);
            const int a0 = c.neg ? (row % p.nbg) * AC : p.n_neg + (row % p.nbg - p.nb_neg) * AC;
#pragma unroll
            for (int w = 0; w < AC; ++w) hwv[w] = w < c.nblk ? 0.5 * p.wgt[a0 + w] : 0.0;
            const int step = c.neg ? -kTile : kTile;
            double *dst = p.partial + c.poff + (int64_t)(c.neg ? nt - 1 - ti0 : ti0) * kTile + lane * 2 + k0 * 64;
            for (int i = 0; i < n; ++i, dst += step) {
                const double *rb_ = red + b * (AC * kTile) + lane * 2 + k0 * 64;
                mbar_wait(red_full + b, bph);
                // straight-line: every index is a compile-time constant.  Absent angles of a ragged block have
                // weight 0 and their warps store zeros, so they are simply summed.
                double2 acc[kPerRed];
#pragma unroll
                for (int kk = 0; kk < kPerRed; ++kk) {
                    const double2 v = *reinterpret_cast<const double2 *>(rb_ + kk * 64);
                    acc[kk] = make_double2(hwv[0] * v.x, hwv[0] * v.y);
                }
#pragma unroll
                for (int w = 1; w < AC; ++w)                                // warps = angles in order (group.py:77)
#pragma unroll
                    for (int kk = 0; kk < kPerRed; ++kk) {
                        const double2 v = *reinterpret_cast<const double2 *>(rb_ + w * kTile + kk * 64);
                        acc[kk].x = fma(hwv[w], v.x, acc[kk].x);            // 0.5*w_a*(in+out) = w_a*psi_mid
                        acc[kk].y = fma(hwv[w], v.y, acc[kk].y);
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(red_empty + b);
#pragma unroll
                for (int kk = 0; kk < kPerRed; ++kk) *reinterpret_cast<double2 *>(dst + kk * 64) = acc[kk];
                b ^= 1;
                bph ^= (b == 0);
            }
        }
    } else {
        // ================= compute warps ==============================================================
        const uint64_t pol = l2_policy(p.l2_hints ? 1 : 0);
        const double nan_v = __longlong_as_double(0x7ff8000000000000LL);
        UniCoef uc;                                                         // constants of the current (row, sigT) pair
        for (int k = 0; k < segs.count; ++k) {
            int rb, ti0, n;
            segs.get(k, rb, ti0, n);
            row_setup<AC>(p, unit_row(p, rb, mb), warp, c);
            double u_h = nan_v;                                             // the hs value uc was built for (NaN: none)
            double carry = 0.0;                                             // vacuum boundary: group.py:48,54
            if (ti0 != 0) {                                                 // continue a row another CTA started
                if (lane == 0 && warp < c.nblk) {
                    while (ld_acquire(&sy->flag[cta - p.gs][warp]) != stamp) {}    // same member of the previous gang
                    carry = sy->carry[cta - p.gs][warp];
                }
                carry = __shfl_sync(0xffffffffu, carry, 0);
            }
            const bool present = warp < c.nblk;
            const int step = c.neg ? -kTile : kTile;
            int64_t off = (int64_t)(c.neg ? nt - 1 - ti0 : ti0) * kTile;    // first cell of the tile
            double *pout = WRITE_PSI ? p.psi_out + c.row + (int64_t)warp * p.Zp + off + lane * 2 : nullptr;
            for (int i = 0; i < n; ++i, off += step) {
                const double *stage = ring + (size_t)st * kStageLen + lane * 2;
                double cb[kL], f[kL];
                mbar_wait(full + st, ph);
                double2 rec = make_double2(nan_v, 0.0);
                if (info) rec = *reinterpret_cast<const double2 *>(ring + (size_t)st * kStageLen + kSlots * kTile);
                if (rec.x == rec.x) {
                    // uniform tile: constant coefficients (rebuilt when the row or the region changes)
                    if (!(rec.x == u_h)) {
                        if (c.neg) uni_setup<true>(c.m, c.two_m, rec.x, lane, s_ap[warp], uc);
                        else uni_setup<false>(c.m, c.two_m, rec.x, lane, s_ap[warp], uc);
                        u_h = rec.x;
                    }
                    uni_maps<READ_PSI>(stage + kTile, stage + (3 + warp) * kTile, rec.y, uc.r, cb);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + st);                 // stage is in registers: release it
                    if (c.neg) uni_solve<true>(cb, uc, carry, f, lane);
                    else uni_solve<false>(cb, uc, carry, f, lane);
                } else {
                    double ca[kL];
                    tile_maps<READ_PSI>(stage, stage + kTile, stage + 2 * kTile, stage + (3 + warp) * kTile, c.m, c.two_m, ca, cb);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + st);                 // stage is in registers: release it
                    if (c.neg) tile_solve<true>(ca, cb, carry, f, lane);
                    else tile_solve<false>(ca, cb, carry, f, lane);
                    if (off + kTile > p.Z) {                                // only the last tile holds padding cells:
#pragma unroll
                        for (int jj = 0; jj < kL; ++jj)                     // exact zeros there, in psi and in the flux
                            if (off + lane * kL + jj >= p.Z) f[jj] = 0.0;
                    }
                }
                if (WRITE_PSI) {
                    if (present) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            st_stream(pout + kk * 64, make_double2(0.5 * f[2 * kk], 0.5 * f[2 * kk + 1]), pol);
                    }
                    pout += step;
                }
                if (!present) {                                             // angle slot a ragged block does not have:
#pragma unroll
                    for (int jj = 0; jj < kL; ++jj) f[jj] = 0.0;            // whatever came out of stale memory is dropped
                }
                double *rw = red + b * (AC * kTile) + warp * kTile + lane * 2;
                mbar_wait(red_empty + b, bph ^ 1);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)                              // in+out; the reducer applies the weights
                    *reinterpret_cast<double2 *>(rw + kk * 64) = make_double2(f[2 * kk], f[2 * kk + 1]);
                __syncwarp();
                if (lane == 0) mbar_arrive(red_full + b);
                if (++st == kStages) { st = 0; ph ^= 1; }
                b ^= 1;
                bph ^= (b == 0);
            }
            // after the head piece: hand the carries to the CTA that continues this row
            if (k == 0 && publishes && lane == 0 && present) {
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

// One record per (group, tile): {hs, cdt} of the tile when all its 256 cells are real (no padding) and share
// bitwise-identical values, {NaN, NaN} otherwise.  One warp per record.
__global__ void tile_info_kernel(const double *__restrict__ hs, const double *__restrict__ cdt, int64_t Z, int64_t Zp, int G,
                                 double *__restrict__ info)
{
    const int64_t nt = Zp / kTile;
    const int64_t rec = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rec >= (int64_t)G * nt) return;
    const int64_t g = rec / nt, t = rec % nt;
    const double *h = hs + g * Zp + t * kTile, *c = cdt ? cdt + g * Zp + t * kTile : nullptr;
    const double h0 = h[0], c0 = c ? c[0] : 0.0;
    bool same = (t + 1) * kTile <= Z;
    for (int i = lane; i < kTile; i += 32) {
        same = same && __double_as_longlong(h[i]) == __double_as_longlong(h0);
        if (c) same = same && __double_as_longlong(c[i]) == __double_as_longlong(c0);
    }
    same = __all_sync(0xffffffffu, same) && h0 == h0 && c0 == c0;
    if (lane == 0) {
        const double nan_v = __longlong_as_double(0x7ff8000000000000LL);
        info[2 * rec] = same ? h0 : nan_v;
        info[2 * rec + 1] = same ? c0 : nan_v;
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
    constexpr int stages = StageCfg<AC, RD>::stages;
    const size_t smem = (size_t)(stages * slots + 2 * AC) * kTileBytes + stages * 16 + (2 * stages + 4) * sizeof(uint64_t);
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
    static const char *nofast = getenv("AE_SWEEP_NO_FAST");                // A/B and parity tests: general path only
    p.tinfo = (nofast && nofast[0] == '1') ? nullptr : s->tile_info;
    return ac == kACbig ? launch_shape<kACbig>(p, ctrl, gate, st) : launch_shape<kACsmall>(p, ctrl, gate, st);
}

}  // namespace ae

extern "C" int32_t ae_partial_count(int32_t A, int32_t n_neg)
{
    const int blk = ae::block_angles(A, n_neg);
    return (n_neg + blk - 1) / blk + (A - n_neg + blk - 1) / blk;
}

// Host-side replay of the kernel's work assignment (same Segments / unit_row / plan_grid code): the (row block,
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
    ae::Segments segs;
    segs.init(t0, t1, (int)nt);
    int64_t j = 0;
    for (int k = 0; k < segs.count; ++k) {
        int rb, ti0, n;
        segs.get(k, rb, ti0, n);
        for (int i = 0; i < n; ++i, ++j) {
            if (j < cap) {
                if (row_blocks) row_blocks[j] = ae::unit_row(p, rb, mb);
                if (tiles) tiles[j] = ti0 + i;
            }
        }
    }
    return W;
}

extern "C" int64_t ae_sweep_sync_bytes(void) { return (int64_t)sizeof(ae::SweepSync); }

extern "C" int ae_tile_info(const double *hs, const double *cdt, int64_t Z, int64_t Zp, int32_t G, double *info, void *stream)
{
    if (!hs || !info || Z < 1 || Zp < Z || Zp % ae::kTile || G < 1) { ae::set_error("ae_tile_info: bad argument"); return AE_BAD_ARG; }
    const int64_t warps = (int64_t)G * (Zp / ae::kTile);
    ae::tile_info_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hs, cdt, Z, Zp, G, info);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}

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
