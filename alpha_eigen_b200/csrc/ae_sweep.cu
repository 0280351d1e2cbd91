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
    const double *inflow;               // [G][A] incident boundary flux per row or NULL (vacuum)
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
            // flux entering the row: vacuum (group.py:48,54) or the angle's boundary_condition (one_direction.py:12)
            double carry = (p.inflow && ti0 == 0 && warp < c.nblk) ? p.inflow[c.row / p.Zp + warp] : 0.0;
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
            // stores one tile's in+out fluxes: psi (when written) and the block's reduction buffer
            auto emit = [&](double (&f)[kL]) {
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
                b ^= 1;
                bph ^= (b == 0);
            };
            int i = 0;
            while (i < n) {
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
                emit(f);
                if (++st == kStages) { st = 0; ph ^= 1; }
                ++i;
                off += step;
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


// ---- phi-only sweeps: several rows per warp -----------------------------------------------------------------------
// The inner iterations of a time step and all steady-state iterations sweep the isotropic source only (no psi in or
// out).  Through sweep_kernel (one angle per warp) they hit two walls at about the same height (ncu, C3): the
// SHARED-MEMORY pipe (84 %: every angle-warp reads the block's 2 KB source tile and pushes 2 KB of fluxes through the
// reduction buffer) and, once that traffic is cut, the LATENCY of the fold / scan chain (27 dependent FP64 and shuffle
// steps per tile, ~800 cycles, with only 16 such chains in flight per SM; FP64 pipe 27 %).  Here a compute warp carries
// R angles of a block of 8R (w, w + 8, ...): one read of the source tile feeds R rows, their weighted fluxes are added
// in registers before they go to the reduction buffer, and the R chains advance in lock step, so 8R chains per SM are
// in flight with the registers of 10 warps.  Same segments, ring, hand-offs and uniform-tile records as sweep_kernel;
// a block's partial flux is sum_w (sum_r w_a f_a), and there are ceil(n/8R) blocks per direction -- the launcher
// reports how many partial slabs it wrote.
constexpr int kPhiCW = 8;                        // compute warps
#ifndef AE_PHI_ROWS
#define AE_PHI_ROWS 4
#endif
#ifndef AE_PHI_STAGES
#define AE_PHI_STAGES 8
#endif
#ifndef AE_PHI_RED_BUFS
#define AE_PHI_RED_BUFS 4        // depth of the reduction ring: absorbs the skew between the compute warps and the reducer
#endif
constexpr int kPhiR = AE_PHI_ROWS, kPhiAC = kPhiCW * kPhiR;
constexpr int kPhiStages = AE_PHI_STAGES, kPhiRB = AE_PHI_RED_BUFS;
constexpr int kPhiStageLen = 2 * kTile + 2;      // hs tile (general tiles only), q tile, record
constexpr int kPhiThreads = (kPhiCW + 2) * 32;
static_assert(kPhiAC <= kMaxBlk, "mailbox slots per CTA");

// General tile of the rows-per-warp kernel (region interface, padded last tile): rows one after the other with the
// per-cell arithmetic of sweep_kernel, weighted sum written straight to the reduction slot.  Out of line, so that
// its registers do not count against the lock-step path.  m / hw / carry: R entries (carry in and out).
__device__ __noinline__ void phi_general_tile(const double *stage, const double *m, const double *hw, double *carry, bool neg,
                                              int lane, int64_t off, int64_t Z, double *rw, uint64_t *empty_bar,
                                              uint64_t *red_empty_bar, uint32_t red_parity)
{
    double g[kL];
#pragma unroll 1
    for (int r = 0; r < kPhiR; ++r) {
        double ca[kL], cb[kL], f[kL];
        tile_maps<false>(stage, stage + kTile, nullptr, nullptr, m[r], 2.0 * m[r], ca, cb);
        if (r == kPhiR - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar);          // last read of the stage
        }
        double c = carry[r];
        if (neg) tile_solve<true>(ca, cb, c, f, lane);
        else tile_solve<false>(ca, cb, c, f, lane);
        carry[r] = c;
#pragma unroll
        for (int j = 0; j < kL; ++j) g[j] = r == 0 ? hw[r] * f[j] : fma(hw[r], f[j], g[j]);
    }
    if (off + kTile > Z) {                                  // only the last tile holds padding cells: exact zeros there
#pragma unroll
        for (int jj = 0; jj < kL; ++jj)
            if (off + lane * kL + jj >= Z) g[jj] = 0.0;
    }
    mbar_wait(red_empty_bar, red_parity);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) *reinterpret_cast<double2 *>(rw + kk * 64) = make_double2(g[2 * kk], g[2 * kk + 1]);
}

__global__ void __launch_bounds__(kPhiThreads, 1) phi_sweep_kernel(const SweepArgs p, const ae_ctrl *ctrl, int gate)
{
    constexpr int AC = kPhiAC, R = kPhiR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);                    // [kPhiStages][kPhiStageLen]
    double *red = ring + kPhiStages * kPhiStageLen;                         // [kPhiRB][kPhiCW][kTile]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + kPhiRB * kPhiCW * kTile);
    uint64_t *empty = full + kPhiStages;
    uint64_t *red_full = empty + kPhiStages;
    uint64_t *red_empty = red_full + kPhiRB;
    __shared__ int s_cta;
    __shared__ __align__(16) double s_ap[AC][kL];                           // per row: a^(jj+1)
    __shared__ double s_hw[AC];                                             // per row: half quadrature weight
    if (gate && ctrl->inner_done) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    SweepSync *sy = p.sync;
    if (tid == 0) {
        s_cta = (int)atomicInc(&sy->ticket, gridDim.x - 1);
        for (int s = 0; s < kPhiStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kPhiCW); }
        for (int b = 0; b < kPhiRB; ++b) { mbar_init(red_full + b, kPhiCW); mbar_init(red_empty + b, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int cta = s_cta;
    const int gang = cta / p.gs, mb = cta - gang * p.gs;
    const unsigned stamp = sy->epoch + 1;
    const int nt = (int)(p.Zp / kTile);
    const int64_t T = p.total_tiles, NC = gridDim.x / p.gs;
    const int64_t t0 = gang * T / NC, t1 = (gang + 1) * T / NC;
    const bool publishes = (t1 % nt) != 0;
    Segments segs;
    segs.init(t0, t1, nt);
    const bool info = p.tinfo != nullptr;
    RowCtx c;
    int st = 0, b = 0;
    uint32_t ph = 0, bph = 0;

    if (warp == kPhiCW + 1) {
        // ================= producer warp ================================================================
        const uint64_t pol = l2_policy(p.l2_hints ? 2 : 0);
        for (int k = 0; k < segs.count; ++k) {
            int rb, ti0, n;
            segs.get(k, rb, ti0, n);
            row_setup<AC>(p, unit_row(p, rb, mb), 0, c);
            const int dir = c.neg ? -1 : 1;
            int tile = c.neg ? nt - 1 - ti0 : ti0;
            const double *recs = info ? p.tinfo + 2 * ((int64_t)c.g * nt + tile) : nullptr;
            double rec_cur = 0.0, rec_nxt = 0.0;
            if (info && lane < n) rec_nxt = __ldg(recs + 2 * dir * lane);
            for (int i = 0; i < n; ++i, tile += dir) {
                if ((i & 31) == 0) {
                    rec_cur = rec_nxt;
                    if (info && i + 32 + lane < n) rec_nxt = __ldg(recs + 2 * dir * (i + 32 + lane));
                }
                const double rec_h = __shfl_sync(0xffffffffu, rec_cur, i & 31);
                const bool uniform = info && rec_h == rec_h;
                double *dst = ring + (size_t)st * kPhiStageLen;
                const int64_t goff = (int64_t)c.g * p.Zp + (int64_t)tile * kTile;
                mbar_wait(empty + st, ph ^ 1);
                if (lane == 0) mbar_expect_tx(full + st, (uint32_t)((uniform ? 1 : 2) * kTileBytes + (info ? 16 : 0)));
                __syncwarp();
                if (lane == 0 && !uniform) bulk_g2s(dst, p.hs + goff, kTileBytes, full + st, pol);
                if (lane == 1) bulk_g2s(dst + kTile, p.q + goff, kTileBytes, full + st, pol);
                if (info && lane == 31) bulk_g2s(dst + 2 * kTile, recs + 2 * dir * i, 16, full + st, pol);
                if (++st == kPhiStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == kPhiCW) {
        // ================= reducer warp: the slots already carry the quadrature weights ==============
        for (int k = 0; k < segs.count; ++k) {
            int rb, ti0, n;
            segs.get(k, rb, ti0, n);
            row_setup<AC>(p, unit_row(p, rb, mb), 0, c);
            const int step = c.neg ? -kTile : kTile;
            double *dst = p.partial + c.poff + (int64_t)(c.neg ? nt - 1 - ti0 : ti0) * kTile + lane * 2;
            for (int i = 0; i < n; ++i, dst += step) {
                const double *rb_ = red + b * (kPhiCW * kTile) + lane * 2;
                mbar_wait(red_full + b, bph);
                double2 acc[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) acc[kk] = *reinterpret_cast<const double2 *>(rb_ + kk * 64);
#pragma unroll
                for (int w = 1; w < kPhiCW; ++w)                            // warps in order: deterministic
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const double2 v = *reinterpret_cast<const double2 *>(rb_ + w * kTile + kk * 64);
                        acc[kk].x += v.x;
                        acc[kk].y += v.y;
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(red_empty + b);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) *reinterpret_cast<double2 *>(dst + kk * 64) = acc[kk];
                if (++b == kPhiRB) { b = 0; bph ^= 1; }
            }
        }
    } else {
        // ================= compute warps: rows warp, warp + 8, ... of the block ========================
        const double nan_v = __longlong_as_double(0x7ff8000000000000LL);
        UniRow u[R];
        const double *ap = &s_ap[warp][0];                                  // row r: + r * kPhiCW * kL
        const double *hws = &s_hw[warp];                                    // row r: + r * kPhiCW
        for (int k = 0; k < segs.count; ++k) {
            int rb, ti0, n;
            segs.get(k, rb, ti0, n);
            const int row = unit_row(p, rb, mb);
            row_setup<AC>(p, row, 0, c);
            const int a0 = c.neg ? (row % p.nbg) * AC : p.n_neg + (row % p.nbg - p.nb_neg) * AC;
            const int nblk = c.nblk;
            const bool neg = c.neg;
            double cin[R];                                                  // fluxes entering the tile
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int slot = warp + kPhiCW * r;
                if (lane == 0) s_hw[slot] = slot < nblk ? 0.5 * p.wgt[a0 + slot] : 0.0;     // half weights (0: absent row)
                // vacuum boundary (group.py:48,54) or the angle's boundary_condition (one_direction.py:12)
                cin[r] = (p.inflow && ti0 == 0 && slot < nblk) ? p.inflow[(int64_t)c.g * p.A + a0 + slot] : 0.0;
            }
            __syncwarp();
            if (ti0 != 0) {                                                 // continue rows another CTA started
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int slot = warp + kPhiCW * r;
                    if (lane == 0 && slot < nblk) {
                        while (ld_acquire(&sy->flag[cta - p.gs][slot]) != stamp) {}
                        cin[r] = sy->carry[cta - p.gs][slot];
                    }
                    cin[r] = __shfl_sync(0xffffffffu, cin[r], 0);
                }
            }
            double u_h = nan_v;
            const int step = neg ? -kTile : kTile;
            int64_t off = (int64_t)(neg ? nt - 1 - ti0 : ti0) * kTile;
            for (int i = 0; i < n; ++i, off += step) {
                const double *stage = ring + (size_t)st * kPhiStageLen + lane * 2;
                double *rw = red + b * (kPhiCW * kTile) + warp * kTile + lane * 2;
                mbar_wait(full + st, ph);
                double2 rec = make_double2(nan_v, 0.0);
                if (info) rec = *reinterpret_cast<const double2 *>(ring + (size_t)st * kPhiStageLen + 2 * kTile);
                if (rec.x == rec.x) {
                    if (!(rec.x == u_h)) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int slot = warp + kPhiCW * r;
                            const double m = fabs(p.mu_ihx[a0 + (slot < nblk ? slot : 0)]);
                            if (neg) row_setup_uniform<true>(m, rec.x, lane, &s_ap[slot][0], u[r]);
                            else row_setup_uniform<false>(m, rec.x, lane, &s_ap[slot][0], u[r]);
                        }
                        u_h = rec.x;
                    }
                    double x[R][kL], g[kL];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const double2 s = *reinterpret_cast<const double2 *>(stage + kTile + kk * 64);
#pragma unroll
                        for (int r = 0; r < R; ++r) { x[r][2 * kk] = s.x * u[r].r; x[r][2 * kk + 1] = s.y * u[r].r; }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + st);
#if defined(AE_PHI_SKELETON) && AE_PHI_SKELETON == 1   // diagnostic build: handshakes and memory traffic only, no scan
#pragma unroll
                    for (int j = 0; j < kL; ++j) { g[j] = 0.0; for (int r = 0; r < R; ++r) g[j] += x[r][j]; }
#else
                    if (neg) phi_rows<true, R>(x, u, ap, kPhiCW * kL, hws, kPhiCW, cin, g, lane);
                    else phi_rows<false, R>(x, u, ap, kPhiCW * kL, hws, kPhiCW, cin, g, lane);
#endif
                    mbar_wait(red_empty + b, bph ^ 1);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) *reinterpret_cast<double2 *>(rw + kk * 64) = make_double2(g[2 * kk], g[2 * kk + 1]);
                } else {
                    double tm[R], thw[R], tc[R];                            // copies: their address goes to the out-of-line path
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int slot = warp + kPhiCW * r;
                        tm[r] = fabs(p.mu_ihx[a0 + (slot < nblk ? slot : 0)]); thw[r] = hws[r * kPhiCW]; tc[r] = cin[r];
                    }
                    phi_general_tile(stage, tm, thw, tc, neg, lane, off, p.Z, rw, empty + st, red_empty + b, bph ^ 1);
#pragma unroll
                    for (int r = 0; r < R; ++r) cin[r] = tc[r];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(red_full + b);
                if (++st == kPhiStages) { st = 0; ph ^= 1; }
                if (++b == kPhiRB) { b = 0; bph ^= 1; }
            }
            if (k == 0 && publishes && lane == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int slot = warp + kPhiCW * r;
                    if (slot < nblk) { sy->carry[cta][slot] = cin[r]; st_release(&sy->flag[cta][slot], stamp); }
                }
            }
        }
    }
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


static int launch_phi(const SweepArgs &p_in, const ae_ctrl *ctrl, int gate, int *P_out, cudaStream_t st)
{
    SweepArgs p = p_in;
    // this kernel's own block shape: 8R angles per block
    p.nb_neg = (p.n_neg + kPhiAC - 1) / kPhiAC;
    p.nbg = p.nb_neg + (p.A - p.n_neg + kPhiAC - 1) / kPhiAC;
    p.total_tiles = (int64_t)p.G * p.nbg * (p.Zp / kTile);
    p.gs = 1; p.subs = p.nb_neg;
    const size_t smem = (size_t)(kPhiStages * kPhiStageLen + kPhiRB * kPhiCW * kTile) * sizeof(double) + (2 * kPhiStages + 2 * kPhiRB) * sizeof(uint64_t);
    static int ctas_per_sm = 0, sms = 0;
    if (!ctas_per_sm) {
        AE_CUDA(cudaFuncSetAttribute((const void *)phi_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 0, dev = 0;
        AE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, phi_sweep_kernel, kPhiThreads, smem));
        if (n < 1) { set_error("phi sweep kernel does not fit on an SM"); return AE_CUDA_ERROR; }
        AE_CUDA(cudaGetDevice(&dev));
        AE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        ctas_per_sm = n;
    }
    const int64_t grid = plan_grid(p, (int64_t)sms * ctas_per_sm);
    phi_sweep_kernel<<<(unsigned)grid, kPhiThreads, smem, st>>>(p, ctrl, gate);
    AE_CUDA(cudaGetLastError());
    *P_out = p.nbg;
    return AE_OK;
}

template <int AC>
static int launch_shape(const SweepArgs &p, const ae_ctrl *ctrl, int gate, int *P_out, cudaStream_t st)
{
    const bool rd = p.psi_in != nullptr, wr = p.psi_out != nullptr;
    static const char *rows1 = getenv("AE_SWEEP_PHI_ROWS");                // "1": phi-only sweeps through sweep_kernel (A/B)
    // phi-only sweeps of many-angle problems: the rows-per-warp kernel, when the caller can take its partial count and
    // its wider blocks (8R angles) are at least 7/8 full (angle-sharded ranks may hold too few angles per direction)
    const int nneg = p.n_neg, npos = p.A - p.n_neg;
    const int padded = ((nneg + kPhiAC - 1) / kPhiAC + (npos + kPhiAC - 1) / kPhiAC) * kPhiAC;
    if (AC == kACbig && !rd && !wr && P_out && 8 * p.A >= 7 * padded && !(rows1 && rows1[0] == '1'))
        return launch_phi(p, ctrl, gate, P_out, st);
    if (P_out) *P_out = p.nbg;
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
// *P_out (may be NULL): number of partial slabs written.  NULL pins it to ae_partial_count(A, n_neg).
int launch_sweep(const ae_slab *s, const double *q, const double *psi_in, double *psi_out, double *partial,
                 void *sync, const ae_ctrl *ctrl, int gate, int *P_out, cudaStream_t st)
{
    if (!s || !q || !partial || !sync || s->Zp % kTile || s->A <= 0 || s->G <= 0) {
        set_error("ae_sweep_set: bad argument");
        return AE_BAD_ARG;
    }
    if (psi_in && !s->cdt) { set_error("ae_sweep_set: psi_in given but slab has no cdt"); return AE_BAD_ARG; }
    SweepArgs p;
    // group-sharded rank: the sweep sees a slab of its gl groups -- rows [g0, g0+gl) of the per-cell arrays (contiguous),
    // psi and partial already hold only those groups
    const int G = s->gl > 0 ? s->gl : s->G;
    const int64_t goff = s->gl > 0 ? (int64_t)s->g0 * s->Zp : 0;
    const int ac = shape_args(s->A, s->n_neg, G, s->Zp, p);
    p.hs = s->hs + goff; p.cdt = s->cdt ? s->cdt + goff : nullptr; p.q = q + goff; p.mu_ihx = s->mu_ihx; p.wgt = s->wgt;
    p.psi_in = psi_in; p.psi_out = psi_out; p.partial = partial; p.sync = (SweepSync *)sync;
    p.inflow = s->inflow ? s->inflow + (s->gl > 0 ? (int64_t)s->g0 * s->A : 0) : nullptr;
    p.Z = s->Z;
    static const char *hints = getenv("AE_SWEEP_L2_HINTS");
    p.l2_hints = !(hints && hints[0] == '0');
    static const char *nofast = getenv("AE_SWEEP_NO_FAST");                // A/B and parity tests: general path only
    p.tinfo = ((nofast && nofast[0] == '1') || !s->tile_info) ? nullptr : s->tile_info + 2 * (goff / kTile);
    return ac == kACbig ? launch_shape<kACbig>(p, ctrl, gate, P_out, st) : launch_shape<kACsmall>(p, ctrl, gate, P_out, st);
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
    return ae::launch_sweep(s, q, psi_in, psi_out, partial, sync, nullptr, 0, nullptr, (cudaStream_t)stream);
}

extern "C" int ae_sweep_set_phi(const ae_slab *s, const double *q, double *partial, void *sync, int32_t *P_out, void *stream)
{
    if (!P_out) { ae::set_error("ae_sweep_set_phi: P_out missing"); return AE_BAD_ARG; }
    int P = 0;
    const int rc = ae::launch_sweep(s, q, nullptr, nullptr, partial, sync, nullptr, 0, &P, (cudaStream_t)stream);
    *P_out = P;
    return rc;
}

extern "C" int ae_reduce_partials(const ae_slab *s, const double *partial, int32_t P, double *phi, void *stream)
{
    if (!s || !partial || !phi || P < 1) { ae::set_error("ae_reduce_partials: bad argument"); return AE_BAD_ARG; }
    // group-sharded rank: the partials hold its gl groups; their sum goes to rows [g0, g0+gl) of phi
    const int64_t n = (int64_t)(s->gl > 0 ? s->gl : s->G) * s->Zp;
    if (s->gl > 0) phi += (int64_t)s->g0 * s->Zp;
    int64_t bx = (n / 2 + 255) / 256;
    if (bx > 148 * 32) bx = 148 * 32;
    ae::reduce_partials_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(partial, P, n, phi);
    AE_CUDA(cudaGetLastError());
    return AE_OK;
}
