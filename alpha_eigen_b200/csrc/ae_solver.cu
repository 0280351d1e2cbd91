// Host-side iteration drivers behind the C ABI: source iteration, power iteration, time stepping.
//
// The loops of the reference (group.py:62-81, 97-116; time_unit.py:17-59) run here as streams of
// kernel launches.  Convergence is decided ON THE DEVICE by the finalize / k-update / outer-check
// kernels with the reference's formulas and tolerances; the host only polls a small control block
// (one pinned-memory read per `batch` source iterations) so that iteration counts -- and therefore
// results -- are identical to the reference's, which stops on the first iterate that passes.
#include <stdlib.h>
#include "ae_common.cuh"

namespace ae {
int launch_sweep(const ae_slab *s, const double *q, const double *psi_in, double *psi_out, double *partial,
                 void *sync, const ae_ctrl *ctrl, int gate, int *P_out, cudaStream_t st);
int launch_finalize(const ae_slab *s, ae_work *w, const double *partial, int P, const double *fixed, const double *phi_in,
                    double *phi_out, double *q_next, int init, double cold, double tol, int iteration, int gate, cudaStream_t st);
int launch_fission_base(const ae_slab *s, ae_work *w, const double *phi, double *base, int with_ext, cudaStream_t st);
int launch_small_source_iteration(const ae_slab *s, ae_work *w, const double *psi_in, double cold, double tol,
                                  int max_iterations, int *launched, cudaStream_t st);
int launch_small_time_march(const ae_slab *s, ae_work *w, double *psi, double *psi_snap, double *phi_snap, int num_steps,
                            double tol, int32_t *outer_hist, int32_t *inner_hist, int *launched, cudaStream_t st);
int comm_reduce_scatter_rows(void *comm, double *rows, int G, int64_t Zp, cudaStream_t st);
int comm_all_gather_rows(void *comm, double *rows, int G, int64_t Zp, int with_pairs, cudaStream_t st);

// multi-GPU: local sum of the angle-block partials, then reduce-scatter by rows so that this rank holds the
// rank-summed flux of ITS cell range in w->reduced (the one exchange of the solve)
static int exchange_partials(const ae_slab *s, ae_work *w, int P, cudaStream_t st)
{
    AE_TRY(ae_reduce_partials(s, w->partial, P, w->reduced, st));
    return comm_reduce_scatter_rows(w->comm, w->reduced, s->G, s->Zp, st);
}
int launch_k_update(const ae_slab *s, ae_work *w, double tol, int iteration, cudaStream_t st);
int launch_assemble_rows(const ae_slab *s, ae_work *w, const double *partial, int P, const double *fixed, const double *phi_in,
                         double *phi_out, int g0, int gl, int slot, cudaStream_t st);
int launch_norm_and_scatter(const ae_slab *s, ae_work *w, const double *phi, double *q_next, int nslots, double tol, int iteration,
                            cudaStream_t st);
int comm_all_gather_group_chunk(void *comm, double *rows, int G, int64_t Zp, int count, int chunk, int with_pairs, cudaStream_t st);
int group_chunks(int G, int nranks);
void group_chunk_range(int gl, int count, int chunk, int *off, int *len);
int comm_size(const void *comm);
cudaStream_t comm_aux_stream(void *comm);
cudaEvent_t comm_event(void *comm, int i);
bool comm_peer_exchange_ready(void *comm, int G, int64_t Zp, cudaStream_t st);
int launch_outer_check(const ae_slab *s, ae_work *w, cudaStream_t st);
int launch_ctrl_init(ae_work *w, cudaStream_t st);

static int poll_ctrl(ae_work *w, cudaStream_t st)
{
    AE_CUDA(cudaMemcpyAsync(w->ctrl_host, w->ctrl, sizeof(ae_ctrl), cudaMemcpyDeviceToHost, st));
    AE_CUDA(cudaStreamSynchronize(st));
    return AE_OK;
}

int check_work(const ae_slab *s, const ae_work *w, const char *who)
{
    if (!s || !w || !w->phi || !w->phi_alt || !w->phi_outer || !w->base || !w->q0 || !w->q1 || !w->partial || !w->norm_scratch ||
        !w->ctrl || !w->ctrl_host || !w->sweep_sync || w->P != ae_partial_count(s->A, s->n_neg)) {
        set_error("%s: incomplete ae_work / ae_slab", who);
        return AE_BAD_ARG;
    }
    if (w->comm && s->gl <= 0 && !w->reduced) { set_error("%s: angle-sharded multi-GPU work needs the `reduced` buffer", who); return AE_BAD_ARG; }
    return AE_OK;
}

// Time-dependent solves.  The angular source of a step is q + cdt*psi^n (time_unit.py:23,53) and only q changes
// between the sweeps of the step, while the sweep is LINEAR in its source: phi = sum_a w_a L_a(q) + sum_a w_a L_a(cdt psi^n).
// The second term is swept once (this function, the only pass that reads psi^n) into w->phi_fixed; the ~40 inner
// iterations of the step then sweep q alone -- 24/A bytes per update instead of 8 + 24/A -- and the flux update adds
// phi_fixed.  AE_WORK_REREAD_PSI (or AE_TD_REREAD_PSI=1) restores the literal scheme (psi^n re-read by every inner sweep) for A/B runs.
static bool reread_psi(const ae_work *w)
{
    static const char *e = getenv("AE_TD_REREAD_PSI");
    return (e && e[0] == '1') || (w->flags & AE_WORK_REREAD_PSI);
}
static int compute_fixed(const ae_slab *s, ae_work *w, const double *psi_in, cudaStream_t st)
{
    if (!w->phi_fixed) { set_error("time-dependent solve: ae_work.phi_fixed is missing"); return AE_BAD_ARG; }
    AE_CUDA(cudaMemsetAsync(w->q1, 0, (size_t)s->G * s->Zp * sizeof(double), st));        // q1 is free until the first update
    // with incident boundary fluxes the sweep is affine, not linear: they belong to the sweeps of q (every inner
    // iteration carries them), so this one runs with vacuum boundaries
    ae_slab vac = *s;
    vac.inflow = nullptr;
    AE_TRY(launch_sweep(&vac, w->q1, psi_in, nullptr, w->partial, w->sweep_sync, w->ctrl, 0, nullptr, st));
    AE_TRY(ae_reduce_partials(s, w->partial, w->P, w->phi_fixed, st));
    if (w->comm && s->gl <= 0) AE_TRY(comm_reduce_scatter_rows(w->comm, w->phi_fixed, s->G, s->Zp, st)); // valid on this rank's cell range
    return AE_OK;
}

// group.py:62-81 / time_unit.py:44-59.  The flux iterate ping-pongs between w->phi and w->phi_alt (the
// update kernel must not read and write the same buffer); on return w->phi is the converged flux.
// *fixed_ready (time-dependent callers): w->phi_fixed already holds the contribution of psi_in.
static int source_iteration(const ae_slab *s, ae_work *w, const double *psi_in, int *fixed_ready, double cold, double tol,
                            int max_iterations, int *iters, int *last_q, cudaStream_t st)
{
    double *q[2] = {w->q0, w->q1};
    double *phi[2] = {w->phi, w->phi_alt};
    const int batch = w->batch > 0 ? w->batch : 1;
    // small slabs: the whole loop runs in one cooperative kernel (ae_small.cu), one host read at the end
    int fused = 0;
    AE_TRY(launch_small_source_iteration(s, w, psi_in, cold, tol, max_iterations, &fused, st));
    if (fused) {
        AE_TRY(poll_ctrl(w, st));
        const int n = w->ctrl_host->inner_iters;
        if (iters) *iters = n;
        if (last_q) *last_q = (n - 1) & 1;
        if (n & 1) { w->phi = phi[1]; w->phi_alt = phi[0]; }
        return w->ctrl_host->inner_done ? AE_OK : AE_NOT_CONVERGED;
    }
    const double *fixed = nullptr, *sweep_psi = psi_in;
    if (psi_in && !reread_psi(w)) {
        if (!fixed_ready || !*fixed_ready) AE_TRY(compute_fixed(s, w, psi_in, st));
        if (fixed_ready) *fixed_ready = 1;
        fixed = w->phi_fixed;
        sweep_psi = nullptr;
    }
    // phi_old = cold (group.py:66 / time_unit.py:45); q0 = base + 0.5*sigma_s*phi_old (group.py:73)
    AE_TRY(launch_finalize(s, w, nullptr, 0, nullptr, nullptr, phi[0], q[0], 1, cold, tol, 0, 0, st));
    int it = 0;         // iterations launched so far
    int status = AE_OK, n_done = 0;
    for (;;) {
        // `assert iteration < max_iterations` (group.py:72) fires when iteration reaches the cap
        const int room = (max_iterations - 1) - it;
        if (room <= 0) { status = AE_NOT_CONVERGED; n_done = it; break; }
        const int nb = batch < room ? batch : room;
        for (int b = 0; b < nb; ++b) {
            const int i = it + b;                       // 0-based; reference iteration number is i+1
            const int gate = b > 0;                     // first of a batch always runs
            int Pw = w->P;                              // partial slabs this sweep writes (fewer from the phi-only kernel)
            AE_TRY(launch_sweep(s, q[i & 1], sweep_psi, nullptr, w->partial, w->sweep_sync, w->ctrl, gate, &Pw, st));
            if (w->comm && s->gl <= 0) {
                // angle-sharded sweeps, cell-sharded update: reduce-scatter the flux, update 1/N of the cells,
                // all-gather the new source (inside launch_finalize)
                AE_TRY(exchange_partials(s, w, Pw, st));
                AE_TRY(launch_finalize(s, w, w->reduced, 1, fixed, phi[i & 1], phi[(i + 1) & 1], q[(i + 1) & 1], 0, cold, tol,
                                       i + 1, gate, st));
            } else {
                AE_TRY(launch_finalize(s, w, w->partial, Pw, fixed, phi[i & 1], phi[(i + 1) & 1], q[(i + 1) & 1], 0, cold, tol,
                                       i + 1, gate, st));
            }
        }
        AE_TRY(poll_ctrl(w, st));
        if (w->ctrl_host->inner_done) { n_done = w->ctrl_host->inner_iters; break; }
        it += nb;
    }
    if (iters) *iters = n_done;
    if (last_q) *last_q = (n_done - 1) & 1;             // the sweep of iteration n read q[(n-1)&1]
    if (n_done & 1) { w->phi = phi[1]; w->phi_alt = phi[0]; }   // iteration n wrote phi[n&1]
    if (w->comm && s->gl <= 0) AE_TRY(comm_all_gather_rows(w->comm, w->phi, s->G, s->Zp, 0, st));     // callers see the whole flux
    return status;
}

}  // namespace ae

using namespace ae;

extern "C" int ae_exchange_partials(const ae_slab *s, ae_work *w, void *stream)
{
    AE_TRY(check_work(s, w, "ae_exchange_partials"));
    if (!w->comm || s->gl > 0) { set_error("ae_exchange_partials: needs an angle-sharded communicator (group-sharded ranks have nothing to reduce)"); return AE_BAD_ARG; }
    return exchange_partials(s, w, w->P, (cudaStream_t)stream);
}

extern "C" int ae_flux_update(const ae_slab *s, ae_work *w, const double *partial, int32_t P, double *q_next, double tol,
                              int32_t iteration, void *stream)
{
    AE_TRY(check_work(s, w, "ae_flux_update"));
    if (!partial || !q_next || P < 1) { set_error("ae_flux_update: bad argument"); return AE_BAD_ARG; }
    AE_TRY(launch_finalize(s, w, partial, P, nullptr, w->phi, w->phi_alt, q_next, 0, 0.0, tol, iteration, 0, (cudaStream_t)stream));
    double *t = w->phi; w->phi = w->phi_alt; w->phi_alt = t;      // w->phi is always the newest iterate
    return AE_OK;
}

// One sweep set + flux update as ONE call (what a time-dependent source iteration consists of).
// AE_EXCHANGE_OVERLAP=1 (group-sharded ranks with the peer exchange of ae_comm.cu, which needs no SM) runs it in two
// chunks of the rank's groups: the first chunk's flux is assembled right after its sweep and travels on a second stream
// while the second, two-group chunk is swept.  It is NOT the default: the exchange does hide, but a sweep of two groups
// has too few rows for the wavefront over cell segments that keeps the SMs busy, and loses more than the exchange costs
// (C4 on 2 GPUs: 30.6 ms per step split, 28.0 ms unsplit; with NCCL on the second stream and 16 SMs held back for it,
// 8 GPUs: 10.1 against 8.35 ms).
extern "C" int ae_sweep_and_update(const ae_slab *s, ae_work *w, const double *q, const double *psi_in, double *psi_out,
                                   double *q_next, double tol, int32_t iteration, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    AE_TRY(check_work(s, w, "ae_sweep_and_update"));
    if (!q || !q_next) { set_error("ae_sweep_and_update: bad argument"); return AE_BAD_ARG; }
    static const char *force = getenv("AE_EXCHANGE_OVERLAP");
    int chunks = 1;
    if (w->comm && s->gl > 0 && force && force[0] == '1' && comm_peer_exchange_ready(w->comm, s->G, s->Zp, st))
        chunks = group_chunks(s->G, comm_size(w->comm));
    if (chunks < 2) {
        AE_TRY(launch_sweep(s, q, psi_in, psi_out, w->partial, w->sweep_sync, w->ctrl, 0, nullptr, st));
        if (w->comm && s->gl <= 0) {
            AE_TRY(exchange_partials(s, w, w->P, st));
            AE_TRY(launch_finalize(s, w, w->reduced, 1, nullptr, w->phi, w->phi_alt, q_next, 0, 0.0, tol, iteration, 0, st));
        } else {
            AE_TRY(launch_finalize(s, w, w->partial, w->P, nullptr, w->phi, w->phi_alt, q_next, 0, 0.0, tol, iteration, 0, st));
        }
    } else {
        cudaStream_t aux = comm_aux_stream(w->comm);
        const size_t row = (size_t)s->A * s->Zp;                                // psi elements per group
        size_t part_off = 0;
        AE_CUDA(cudaEventRecord(comm_event(w->comm, 3), st));                   // aux starts after everything queued so far
        AE_CUDA(cudaStreamWaitEvent(aux, comm_event(w->comm, 3), 0));
        for (int c = 0; c < chunks; ++c) {
            int off, len;
            group_chunk_range(s->gl, chunks, c, &off, &len);
            ae_slab view = *s;
            view.g0 = s->g0 + off; view.gl = len;
            double *part = w->partial + part_off;                               // this chunk's own [P][len][Zp] region
            AE_TRY(launch_sweep(&view, q, psi_in ? psi_in + off * row : nullptr, psi_out ? psi_out + off * row : nullptr,
                                part, w->sweep_sync, w->ctrl, 0, nullptr, st));
            AE_TRY(launch_assemble_rows(s, w, part, w->P, nullptr, w->phi, w->phi_alt, view.g0, len, c, st));
            AE_CUDA(cudaEventRecord(comm_event(w->comm, c), st));
            AE_CUDA(cudaStreamWaitEvent(aux, comm_event(w->comm, c), 0));
            AE_TRY(comm_all_gather_group_chunk(w->comm, w->phi_alt, s->G, s->Zp, chunks, c, c == chunks - 1, aux));
            part_off += (size_t)w->P * len * s->Zp;
        }
        AE_CUDA(cudaEventRecord(comm_event(w->comm, 2), aux));
        AE_CUDA(cudaStreamWaitEvent(st, comm_event(w->comm, 2), 0));
        AE_TRY(launch_norm_and_scatter(s, w, w->phi_alt, q_next, chunks, tol, iteration, st));
    }
    double *t = w->phi; w->phi = w->phi_alt; w->phi_alt = t;                   // w->phi is always the newest iterate
    return AE_OK;
}

extern "C" int ae_source_iteration(const ae_slab *s, ae_work *w, const double *psi_in, double cold_start, double tol,
                                   int32_t max_iterations, int32_t *iters, int32_t *last_q_index, void *stream)
{
    AE_TRY(check_work(s, w, "ae_source_iteration"));
    return source_iteration(s, w, psi_in, nullptr, cold_start, tol, max_iterations, iters, last_q_index, (cudaStream_t)stream);
}

// group.py:97-116
extern "C" int ae_power_iteration(const ae_slab *s, ae_work *w, double tol, int32_t max_iterations, double si_tol,
                                  int32_t si_max_iterations, double *psi_out, double *k_new, double *k_old,
                                  int32_t *iters, double *k_hist, int32_t *si_hist, int32_t *last_q_index,
                                  void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    AE_TRY(check_work(s, w, "ae_power_iteration"));
    AE_TRY(launch_ctrl_init(w, st));                    // k.new = 1, k.old = 1.1 (eigenvalue.py:3-4)
    int iteration = 0, status = AE_OK, last_q = 0;
    double kn = 1.0, ko = 1.1;
    for (;;) {
        iteration += 1;
        if (!(iteration < max_iterations)) { status = AE_NOT_CONVERGED; iteration -= 1; break; }   // group.py:102
        AE_TRY(launch_fission_base(s, w, w->phi_outer, w->base, 0, st));    // group.py:104: source = zeros + fission
        int si = 0;
        status = source_iteration(s, w, nullptr, nullptr, 1e-12, si_tol, si_max_iterations, &si, &last_q, st);
        if (si_hist) si_hist[iteration - 1] = si;
        if (status != AE_OK) { iteration -= 1; break; }     // *iters = completed power iterations
        AE_TRY(launch_k_update(s, w, tol, iteration, st));
        AE_TRY(poll_ctrl(w, st));
        kn = w->ctrl_host->k_new;
        ko = w->ctrl_host->k_old;
        if (k_hist) k_hist[iteration - 1] = kn;
        if (w->ctrl_host->power_done) break;
    }
    if (status == AE_OK && psi_out) {
        // angle.psi_midpoint of the last sweep executed (group.py:59): replay it with the same source
        const double *q = last_q ? w->q1 : w->q0;
        AE_TRY(launch_sweep(s, q, nullptr, psi_out, w->partial, w->sweep_sync, w->ctrl, 0, nullptr, st));
        AE_CUDA(cudaStreamSynchronize(st));
    }
    if (k_new) *k_new = kn;
    if (k_old) *k_old = ko;
    if (iters) *iters = iteration;
    if (last_q_index) *last_q_index = last_q;
    return status;
}

// time_unit.py:29-42 (+R3, R4): one backward-Euler step
extern "C" int ae_time_step(const ae_slab *s, ae_work *w, const double *psi_prev, double *psi_next, double tol,
                            int32_t *outer_iters, int32_t *inner_iters, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    AE_TRY(check_work(s, w, "ae_time_step"));
    if (!psi_prev || !psi_next || !s->cdt) { set_error("ae_time_step: psi / cdt missing"); return AE_BAD_ARG; }
    const size_t nbytes = (size_t)s->G * s->Zp * sizeof(double);
    AE_CUDA(cudaMemsetAsync(w->phi_outer, 0, nbytes, st));      // time_unit.py:30 phi = zeros
    int outer = 0, inner_total = 0, last_q = 0, status = AE_OK;
    int fixed_ready = 0;                                        // phi_fixed is swept once per step, on first use
    for (;;) {
        outer += 1;
        if (!(outer < 100)) { status = AE_NOT_CONVERGED; break; }               // time_unit.py:35
        AE_TRY(launch_fission_base(s, w, w->phi_outer, w->base, 1, st));        // time_unit.py:37-39
        int si = 0;
        status = source_iteration(s, w, psi_prev, &fixed_ready, 0.0, tol, 100, &si, &last_q, st);    // time_unit.py:40
        inner_total += si;
        if (status != AE_OK) break;
        AE_TRY(launch_outer_check(s, w, st));                                   // time_unit.py:41 (+R4)
        AE_TRY(poll_ctrl(w, st));
        if (w->ctrl_host->outer_change < tol) break;                            // time_unit.py:42
    }
    if (outer_iters) *outer_iters = outer;
    if (inner_iters) *inner_iters = inner_total;
    if (status != AE_OK) return status;
    // commit: replay the last sweep with its own source, now writing psi (time_unit.py:26-27)
    const double *q = last_q ? w->q1 : w->q0;
    AE_TRY(launch_sweep(s, q, psi_prev, psi_next, w->partial, w->sweep_sync, w->ctrl, 0, nullptr, st));
    return AE_OK;
}

// time_unit.py:17-27 (+R2)
extern "C" int ae_time_march(const ae_slab *s, ae_work *w, double *psi, double *psi_snap, double *phi_snap,
                             int32_t num_steps, double tol, int32_t *outer_hist, int32_t *inner_hist, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    AE_TRY(check_work(s, w, "ae_time_march"));
    if (!psi) { set_error("ae_time_march: psi missing"); return AE_BAD_ARG; }
    // psi holds the groups this rank sweeps (all of them unless group-sharded); phi always has all G rows
    const size_t n_phi = (size_t)s->G * s->Zp, n_psi = (size_t)(s->gl > 0 ? s->gl : s->G) * s->Zp * s->A;
    // small slabs: every step, loop and snapshot in one cooperative kernel (ae_small.cu)
    if (num_steps > 0) {
        int32_t *d_hist = nullptr;
        AE_CUDA(cudaMallocAsync((void **)&d_hist, sizeof(int32_t) * 2 * num_steps, st));
        AE_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * 2 * num_steps, st));
        int fused = 0;
        const int rc = launch_small_time_march(s, w, psi, psi_snap, phi_snap, num_steps, tol, d_hist, d_hist + num_steps, &fused, st);
        if (rc != AE_OK || !fused) {
            cudaFreeAsync(d_hist, st);
            if (rc != AE_OK) return rc;
        } else {
            if (outer_hist) AE_CUDA(cudaMemcpyAsync(outer_hist, d_hist, sizeof(int32_t) * num_steps, cudaMemcpyDeviceToHost, st));
            if (inner_hist) AE_CUDA(cudaMemcpyAsync(inner_hist, d_hist + num_steps, sizeof(int32_t) * num_steps, cudaMemcpyDeviceToHost, st));
            AE_TRY(poll_ctrl(w, st));
            cudaFreeAsync(d_hist, st);
            if (w->ctrl_host->inner_iters & 1) { double *t = w->phi; w->phi = w->phi_alt; w->phi_alt = t; }
            if (w->ctrl_host->reserved) return AE_NOT_CONVERGED;
            if (psi_snap)
                AE_CUDA(cudaMemcpyAsync(psi, psi_snap + (size_t)(num_steps - 1) * n_psi, n_psi * sizeof(double),
                                        cudaMemcpyDeviceToDevice, st));
            AE_CUDA(cudaStreamSynchronize(st));
            return AE_OK;
        }
    }
    for (int step = 0; step < num_steps; ++step) {
        const double *prev = psi;
        double *next = psi;
        if (psi_snap) {
            prev = step == 0 ? psi : psi_snap + (size_t)(step - 1) * n_psi;
            next = psi_snap + (size_t)step * n_psi;
        }
        int outer = 0, inner = 0;
        const int status = ae_time_step(s, w, prev, next, tol, &outer, &inner, stream);
        if (outer_hist) outer_hist[step] = outer;
        if (inner_hist) inner_hist[step] = inner;
        if (status != AE_OK) return status;
        if (phi_snap)                                                            // time_unit.py:25
            AE_CUDA(cudaMemcpyAsync(phi_snap + (size_t)step * n_phi, w->phi, n_phi * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
    }
    if (psi_snap && num_steps > 0)
        AE_CUDA(cudaMemcpyAsync(psi, psi_snap + (size_t)(num_steps - 1) * n_psi, n_psi * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
    AE_CUDA(cudaStreamSynchronize(st));
    return AE_OK;
}
