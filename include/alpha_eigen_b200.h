/*
 * alpha_eigen_b200.h -- C ABI of the B200 (sm_100a) slab S_N hot path.
 *
 * The reference (SoftwareDevEngResearch/alpha-eigen) is pure Python and has no FFI of its
 * own; its boundary is the method surface of OneGroup / TimeUnit.  Each entry point below
 * replaces the body of one reference method and is what a ctypes binding inside that
 * method would call (INTEGRATION.md shows the stubs).  Plain pointers and sizes only: no
 * torch, no C++ types.  Every `double*` / `int32_t*` inside ae_slab / ae_work is a DEVICE
 * pointer unless its comment says host.  All functions return an ae_status.
 *
 * Cell storage order ("lane-blocked tiles")
 *   Per-cell arrays are stored in tiles of AE_TILE = 256 cells; inside a tile logical cell
 *   r = lane*8 + j (lane 0..31, j 0..7) lives at offset (j/2)*64 + lane*2 + (j%2), so that a
 *   warp whose lane owns 8 consecutive cells reads them with four perfectly coalesced 16-byte
 *   loads.  Z is padded to Zp = ae_padded_cells(Z); padding cells carry zeros in every input.
 *   ae_cell_offset() gives the map; the Python host applies it once at upload / download.
 *   Array shapes below are written [slow]...[fast] with the cell index always fastest.
 */
#ifndef ALPHA_EIGEN_B200_H
#define ALPHA_EIGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AE_ABI_VERSION 2
#define AE_TILE 256          /* cells per tile */
#define AE_LANE_CELLS 8      /* consecutive cells owned by one lane */

typedef enum ae_status {
    AE_OK = 0,
    AE_NOT_CONVERGED = 1,    /* the reference raises AssertionError (group.py:72,102; time_unit.py:35,51) */
    AE_BAD_ARG = 2,          /* includes |mu| <= 1e-10 (group.py:44) */
    AE_CUDA_ERROR = 3,       /* CUDA / NCCL / cuSOLVER failure; see ae_last_error() */
} ae_status;

/* Device-side control block (one per solver instance; also mirrored in pinned host memory). */
typedef struct ae_ctrl {
    double change;           /* last ||phi-phi_old||/||phi|| of the inner (scatter) iteration */
    double outer_change;     /* same for the outer (fission) loop of a time step */
    double k_new, k_old;     /* power iteration (group.py:108,114) */
    double k_change;         /* |k_new - k_old| before k_old is overwritten */
    int32_t inner_done, inner_iters;
    int32_t power_done, power_iters;
    uint32_t ticket;         /* last-block election counter */
    int32_t reserved;
} ae_ctrl;

/* Read-only problem description.  Built once per (slab, dt). */
typedef struct ae_slab {
    int64_t Z;               /* logical cells  (slab.py:8 num_zones) */
    int64_t Zp;              /* padded cells, multiple of AE_TILE */
    int32_t A;               /* angles held by THIS rank (slab.py:10), negatives first */
    int32_t n_neg;           /* how many of them have mu < 0 */
    int32_t G;               /* energy groups (slab.py:12) */
    int32_t M;               /* materials */
    const double *mu_ihx;    /* [A] mu/hx, signed (group.py:46,50) */
    const double *wgt;       /* [A] quadrature weights (slab.py:11) */
    const int32_t *mat;      /* [Zp] material index per cell (slab.py:41-56) */
    const double *scat;      /* [M][G][G] sigma_s(g'->g) at [m][g'][g] (material.py:38, zone.py:50) */
    const double *chi;       /* [M][G] (material.py:10) */
    const double *nusf;      /* [M][G] (material.py:13-19) */
    const double *hs;        /* [G][Zp] 0.5*sigT seen by the sweep (group.py:50; time_unit.py:24 adds 1/(v dt)) */
    const double *cdt;       /* [G][Zp] inv_speed/dt (time_unit.py:23) or NULL in steady state */
    const double *s_ext;     /* [G][Zp] fixed external source (group.source at time_unit.py:19) or NULL */
    const double *tile_info; /* [G][Zp/AE_TILE][2] records written by ae_tile_info for THESE hs / cdt arrays, or NULL.
                                Lets the sweep skip the per-cell reciprocals in tiles whose cells share one sigT
                                (material regions, slab.py:33-56); results are bit-identical with and without it */
    const double *inflow;    /* [G][A] flux entering the slab at each row's upstream boundary (x = 0 for mu > 0, x = L for mu < 0):
                                OneDirection.boundary_condition (one_direction.py:12,24-31).  NULL = vacuum, which is what the
                                reference always runs with (group.py:48,54) */
    int32_t g0, gl;          /* group-sharded ranks (several GPUs, G >= 2 ranks): this rank sweeps groups [g0, g0+gl) with ALL
                                angles and psi / partial hold only those groups ([gl][A][Zp], [P][gl][Zp]); every other array
                                keeps all G rows.  A group's flux is then complete on its owner: no reduction, one all-gather
                                of the new flux per source iteration.  gl = 0: not group-sharded (one GPU, or angle-sharded) */
} ae_slab;

/* Mutable buffers, all caller-allocated (torch owns the memory). */
typedef struct ae_work {
    double *phi;             /* [G][Zp] current scalar-flux iterate (phi_old of group.py:80) */
    double *phi_alt;         /* [G][Zp] the other half of the ping-pong: the flux-update kernel reads one and writes
                                the other, and the solver entry points SWAP these two pointers so that `phi`
                                is the newest iterate when they return */
    double *phi_outer;       /* [G][Zp] power-iteration phi.old / time-step outer phi */
    double *base;            /* [G][Zp] source that is fixed during the inner iteration */
    double *q0, *q1;         /* [G][Zp] ping-pong isotropic sweep source */
    double *partial;         /* [P][G][Zp] per-angle-block partial scalar flux */
    double *reduced;         /* [G][Zp] sum of the partials, reduce-scattered over ranks (multi-GPU only, else NULL) */
    double *phi_fixed;       /* [G][Zp] time steps only: scalar flux driven by the previous step's angular flux alone,
                                sum_a w_a L_a(cdt*psi^n).  The sweep is linear in its source and this part of the source
                                does not change inside a step (time_unit.py:23), so it is swept ONCE per step and the
                                inner iterations sweep the isotropic source only (no psi traffic).  May be NULL when
                                no time-dependent entry point is used */
    double *norm_scratch;    /* [4 * ae_norm_blocks(Zp) + 2 * G] */
    ae_ctrl *ctrl;           /* device */
    ae_ctrl *ctrl_host;      /* pinned host */
    void *sweep_sync;        /* device, ae_sweep_sync_bytes() bytes, zero-filled once */
    void *comm;              /* ae_comm handle for the flux exchange, or NULL (single GPU) */
    int32_t P;               /* ae_partial_count(A, n_neg) */
    int32_t batch;           /* source iterations launched between host polls (>=1) */
    int32_t flags;           /* AE_WORK_* bits, 0 for the default behaviour */
    int32_t reserved;
} ae_work;

#define AE_WORK_NO_FUSED 1   /* never take the fused small-slab kernels (streaming kernels + host-driven loops only) */
#define AE_WORK_REREAD_PSI 2 /* time steps: every inner sweep re-reads psi^n (the literal scheme of time_unit.py:53)
                                instead of sweeping its contribution once per step into phi_fixed */

/* ---- introspection ---------------------------------------------------------------- */
int ae_abi_version(void);
const char *ae_last_error(void);
int64_t ae_padded_cells(int64_t Z);
int64_t ae_padded_cells_sharded(int64_t Z, int32_t nranks);   /* multiple of 256*nranks: equal cell ranges per rank */
int64_t ae_cell_offset(int64_t z);              /* storage offset of logical cell z */
int32_t ae_partial_count(int32_t A, int32_t n_neg);
int64_t ae_norm_blocks(int64_t Zp);

/* ---- sweep: replaces the loop of OneGroup.sweep1D calls (group.py:41-60, 74-77) ------ */
/* One sweep set over all local angles x groups with isotropic source q [G][Zp].
 *   psi_in  != NULL : time-dependent source q + cdt*psi_in (time_unit.py:23,53); [G][A][Zp]
 *   psi_out != NULL : angle.psi_midpoint (group.py:59) is written; may alias psi_in
 *   partial         : [P][G][Zp] receives sum_a w_a psi_a per angle block (group.py:77)
 */
int ae_sweep_set(const ae_slab *s, const double *q, const double *psi_in, double *psi_out,
                 double *partial, void *sync, void *stream);
/* The same for a sweep without angular flux in or out (every steady-state iteration, the inner iterations of a
 * time step): may run a kernel with wider angle blocks that writes FEWER partial slabs; *P_out receives how many
 * (<= ae_partial_count), the value to hand to ae_reduce_partials / ae_flux_update. */
int ae_sweep_set_phi(const ae_slab *s, const double *q, double *partial, void *sync, int32_t *P_out, void *stream);
/* Uniform-tile records for ae_slab.tile_info: info[g][t] = {hs, cdt} of tile t of group g when its 256 cells are all
 * real and share bitwise-identical hs (and cdt, when cdt != NULL) values, {NaN, NaN} otherwise.  hs / cdt: [G][Zp]. */
int ae_tile_info(const double *hs, const double *cdt, int64_t Z, int64_t Zp, int32_t G, double *info, void *stream);
/* `sync` is a device buffer of ae_sweep_sync_bytes() bytes, zero-filled ONCE by the caller: the
 * mailbox through which persistent CTAs hand running boundary fluxes to each other.  One buffer
 * per stream of sweeps (ae_work carries its own). */
int64_t ae_sweep_sync_bytes(void);
/* Host-only replay of the sweep kernel's work assignment for `resident_ctas` co-resident CTAs: writes the (row block,
 * tile index in sweep order) pairs logical CTA `cta` visits, in order, into row_blocks / tiles (capacity cap) and
 * returns how many there are (-1: bad argument); *grid_out = CTAs launched, *gang_out = gang size (CTAs that walk
 * the same cells together so that hs/q/cdt are fetched from HBM once per gang).  No GPU needed (test hook). */
int64_t ae_sweep_plan(int32_t A, int32_t n_neg, int32_t G, int64_t Zp, int32_t resident_ctas, int32_t cta,
                      int32_t *grid_out, int32_t *gang_out, int32_t *row_blocks, int32_t *tiles, int64_t cap);
/* phi[G][Zp] = sum_p partial[p] in block order; padding cells are zeroed. */
int ae_reduce_partials(const ae_slab *s, const double *partial, int32_t P, double *phi, void *stream);

/* Scalar-flux assembly after a sweep set (group.py:77-80 + group.py:73 for the next iterate):
 *   phi_new = sum_p partial[p]; change = ||phi_new - w->phi|| / ||phi_new||; phi_new is written to
 *   w->phi_alt and the two pointers are swapped; q_next = w->base + 0.5 * sum_g' sigma_s(g'->g) phi_new(g').
 * ctrl.change / inner_iters / inner_done are updated on the device (done when change < tol).
 * The solver loops below are this call alternated with ae_sweep_set. */
int ae_flux_update(const ae_slab *s, ae_work *w, const double *partial, int32_t P, double *q_next,
                   double tol, int32_t iteration, void *stream);

/* ae_sweep_set followed by ae_flux_update as one call: the unit a time-dependent source iteration repeats (the sweep's
 * partials go to w->partial).  Group-sharded ranks (ae_slab.gl > 0) overlap the flux exchange with the sweep: their
 * groups are swept in two chunks and the first chunk's flux is assembled and all-gathered on a second stream while the
 * second chunk is swept.  Same results as the two calls (the norm's sum is taken chunk by chunk). */
int ae_sweep_and_update(const ae_slab *s, ae_work *w, const double *q, const double *psi_in, double *psi_out,
                        double *q_next, double tol, int32_t iteration, void *stream);

/* Multi-GPU only: w->reduced := sum of the local angle-block partials, then reduce-scattered by rows over
 * the ranks so that this rank holds the rank-summed flux of its own cell range [r*Zp/N, (r+1)*Zp/N).
 * Follow with ae_flux_update(s, w, w->reduced, 1, ...), which updates that range and all-gathers the
 * new source.  Zp must be a multiple of 256*N (ae_padded_cells_sharded). */
int ae_exchange_partials(const ae_slab *s, ae_work *w, void *stream);

/* ---- layout conversion at the boundary (device to device) ------------------------------- */
/* cells: logical [Z][G] (phi.new of group.py:18) <-> storage [G][Zp] */
int ae_pack_cells(const double *logical, int64_t Z, int64_t Zp, int32_t G, double *storage, void *stream);
int ae_unpack_cells(const double *storage, int64_t Z, int64_t Zp, int32_t G, double *logical, void *stream);
/* Cell-range variants (cell-sharded host I/O on several GPUs): `logical` holds `count` cells [count][G],
 * `storage` points at the first cell of a range of `cells` storage cells (whole tiles) inside rows of stride ld. */
int ae_pack_cells_range(const double *logical, int64_t count, int64_t cells, int64_t ld, int32_t G,
                        double *storage, void *stream);
int ae_unpack_cells_range(const double *storage, int64_t count, int64_t cells, int64_t ld, int32_t G,
                          double *logical, void *stream);
/* psi: logical [Z][A_total][G] (TimeUnit.psi[..., step], time_unit.py:14) <-> storage [G][A][Zp];
 * angle_index [A] (device) maps local angle a to the reference angle number. */
int ae_pack_psi(const double *logical, int64_t Z, int64_t Zp, int32_t A_total, int32_t G, const int32_t *angle_index,
                int32_t A, double *storage, void *stream);
int ae_unpack_psi(const double *storage, int64_t Z, int64_t Zp, int32_t A_total, int32_t G, const int32_t *angle_index,
                  int32_t A, double *logical, void *stream);

/* ---- source iteration: OneGroup.source_iteration_for_scalar_flux (group.py:62-81) and
 *      TimeUnit.time_dependent_source_iteration (time_unit.py:44-59) ------------------------ */
/* w->base holds the fixed source; cold_start is 1e-12 (group.py:66) or 0 (time_unit.py:45).
 * On return w->phi is the converged flux and *last_q_index says which of q0/q1 fed the last sweep. */
int ae_source_iteration(const ae_slab *s, ae_work *w, const double *psi_in, double cold_start,
                        double tol, int32_t max_iterations, int32_t *iters, int32_t *last_q_index,
                        void *stream);

/* ---- power iteration: OneGroup.perform_power_iteration (group.py:97-116) --------------- */
/* w->phi_outer is phi.old (in/out); w->phi returns phi.new.  k_hist (host, may be NULL) receives
 * k.new per iteration so the caller can print the reference's banner (group.py:110-113).
 * psi_out (device, may be NULL) receives angle.psi_midpoint of the last sweep; *last_q_index
 * says which of q0/q1 fed that sweep so it can also be replayed later with ae_sweep_set. */
int ae_power_iteration(const ae_slab *s, ae_work *w, double tol, int32_t max_iterations,
                       double si_tol, int32_t si_max_iterations, double *psi_out,
                       double *k_new, double *k_old, int32_t *iters, double *k_hist, int32_t *si_hist,
                       int32_t *last_q_index, void *stream);

/* ---- time stepping: TimeUnit.calculate_one_step_flux (time_unit.py:29-42) and
 *      calculate_time_dependent_flux (time_unit.py:17-27), with repairs R2-R5 ---- */
/* One backward-Euler step: reads psi_prev [G][A][Zp], writes psi_next (may alias psi_prev). */
int ae_time_step(const ae_slab *s, ae_work *w, const double *psi_prev, double *psi_next, double tol,
                 int32_t *outer_iters, int32_t *inner_iters, void *stream);
/* num_steps steps.  psi [G][A][Zp] is psi0 on entry and the last step's psi on exit (in place)
 * unless psi_snap != NULL, in which case step n reads slot n-1 and writes slot n of
 * psi_snap [num_steps][G][A][Zp] (slot 0 reads `psi`).  phi_snap [num_steps][G][Zp] may be NULL. */
int ae_time_march(const ae_slab *s, ae_work *w, double *psi, double *psi_snap, double *phi_snap,
                  int32_t num_steps, double tol, int32_t *outer_hist, int32_t *inner_hist, void *stream);

/* ---- ensembles of small slabs (BASELINE configs[4]) ----------------------------------------------------
 * One member = one TimeUnit.calculate_time_dependent_flux (time_unit.py:17-27) on its own slab and buffers. */
typedef struct ae_member {
    const ae_slab *slab;     /* host pointers to the member's problem and work buffers (all members: same Z, A, */
    ae_work *work;           /* n_neg, G; single GPU; small enough for the fused kernels)                       */
    double *psi;             /* [G][A][Zp] psi0 on entry, the last step's psi on exit */
    double *psi_snap;        /* [num_steps][G][A][Zp] or NULL, as in ae_time_march */
    double *phi_snap;        /* [num_steps][G][Zp] or NULL */
    int32_t *hist;           /* DEVICE [2][num_steps]: outer, then inner iterations per step */
} ae_member;
/* Marches all members in ONE cooperative launch: the grid is cut into teams of CTAs (one team = the row blocks
 * of one member) that pull members from a device-side queue and synchronise only among themselves; every member
 * keeps its own convergence history, so results equal `count` separate ae_time_march calls.  status [count]
 * (host, may be NULL) receives AE_OK / AE_NOT_CONVERGED per member; *teams_out (may be NULL) the number of
 * members in flight.  On return each member's work->phi is the flux of its last step. */
int ae_ensemble_time_march(const ae_member *members, int32_t count, int32_t num_steps, double tol,
                           int32_t *status, int32_t *teams_out, void *stream);

/* ---- DMD: TimeUnit.compute_alpha_eigenvalues (time_unit.py:61-92) ---------------------- */
/* gram[n][n] (device, row-major) = S S^T for snapshot rows S [n][ld] restricted to len columns. */
int ae_gram(const double *snap, int64_t ld, int64_t len, int32_t n, double *gram, void *stream);
/* Small dense part from the (allreduced) Gram matrix of snapshots first..first+K (K+1 rows):
 * rank rule of time_unit.py:75-76 (keep sigma_j while 1 - cumsum/sum > rank_tol), Atilde,
 * eigenvalues, alpha = (1-1/lambda)/dt (time_unit.py:91).  gram is a DEVICE pointer
 * [K+1][K+1]; alpha_re / alpha_im / sigma are HOST arrays of length K (sigma may be NULL);
 * *rank is the number of alpha values written. */
int ae_dmd_from_gram(const double *gram, int32_t K, double dt, double rank_tol,
                     double *alpha_re, double *alpha_im, double *sigma, int32_t *rank, void *stream);

/* R-only TSQR of the snapshot window: r[n][n] (device, row-major, upper triangular) with
 * S^T = Q r for the tall matrix S^T (len x n); Q is never formed.  Householder, backward stable:
 * resolves the singular values the reference's rank rule looks at (down to 1e-13 of the sum),
 * which the Gram route cannot (it squares the condition number). */
int ae_tsqr(const double *snap, int64_t ld, int64_t len, int32_t n, double *r, void *stream);
/* QR of `count` stacked n x n triangles (device, row-major): the cross-rank step after an
 * allgather of the per-rank factors of a row-sharded snapshot matrix. */
int ae_tsqr_merge(const double *r_stack, int32_t count, int32_t n, double *r, void *stream);
/* As ae_dmd_from_gram, from the triangular factor r [K+1][K+1] (device): thin SVD of r[:, 0:K]
 * gives sigma and V of X; Atilde = U_r^T r[:, 1:K+1] V S^-1; eigenvalues of Atilde. */
int ae_dmd_from_r(const double *r, int32_t K, double dt, double rank_tol,
                  double *alpha_re, double *alpha_im, double *sigma, int32_t *rank, void *stream);
/* The same for `count` windows at once (an ensemble's members): r_stack [count][K+1][K+1] (device); host outputs
 * alpha_re / alpha_im / sigma [count][K], rank / status [count] (status[m] is the code ae_dmd_from_r would have
 * returned for window m).  Windows of K <= 118 columns run as one thread block each of ONE launch (one-sided Jacobi
 * SVD, rank rule, Atilde, balanced Hessenberg-QR eigenvalues in shared memory; ae_dmd_from_r takes the same kernel);
 * wider ones go through cuSOLVER one by one. */
int ae_dmd_from_r_batched(const double *r_stack, int32_t count, int32_t K, double dt, double rank_tol,
                          double *alpha_re, double *alpha_im, double *sigma, int32_t *rank, int32_t *status, void *stream);
/* Diagnostic: Jacobi sweeps the first window of this host thread's last block-kernel DMD needed. */
int ae_dmd_last_jacobi_sweeps(void);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ---------------------------------- */
int ae_comm_unique_id(void *id128);                                     /* host, 128 bytes */
int ae_comm_init(void **comm, int32_t rank, int32_t nranks, const void *id128);
int ae_comm_allreduce_sum(void *comm, double *buf, int64_t count, void *stream);
/* Group-sharded ranks: rows [G][Zp], rank r contributes rows [g0_r, g0_r + gl_r) of the balanced contiguous
 * partition of the G groups (ae_group_range), in place. */
int ae_comm_gather_groups(void *comm, double *rows, int32_t G, int64_t Zp, void *stream);
/* How the group-sharded flux exchange of this communicator travels: 1 = copy engines over NVLink into IPC-shared peer
 * buffers, signalled by stream memory operations (no SM; opt-in with AE_COMM_PEER_COPY=1); 0 = NCCL all-gather (the
 * default, and the fallback when the peers cannot be mapped); -1 = not decided yet (no exchange has run). */
int32_t ae_comm_exchange_mode(void *comm);
void ae_group_range(int32_t G, int32_t nranks, int32_t rank, int32_t *g0, int32_t *gl);
/* rows [G][Zp]: every rank contributes its cell range of every row, in place (all-gather). */
int ae_comm_gather_rows(void *comm, double *rows, int32_t G, int64_t Zp, void *stream);
int ae_comm_destroy(void *comm);

#ifdef __cplusplus
}
#endif
#endif
