// NCCL plumbing for the angle-sharded multi-GPU path (one process per GPU).
//
// The reference is single-process; sharding the (angle, group) rows over ranks adds exactly one
// exchange per source iteration: the sum over ranks of the partial scalar flux (group.py:77 summed
// over ALL angles).  NCCL is resolved at run time from the copy torch already loaded (dlopen by
// soname), so this library has no link-time NCCL dependency and loads on CPU-only hosts.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include "ae_common.cuh"

namespace ae {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static Nccl g_nccl;

// What an ae_comm handle points to.
struct Comm {
    ncclComm_t nccl;
    int rank, nranks;
    double *pair_local;     // device [2]: this rank's (num, den) of a sharded norm
    double *pairs;          // device [nranks][2]: everybody's, after the all-gather
    double *stage;          // device [nranks][block]: equal-sized blocks for the all-gather of unequal group ranges
    size_t stage_elems;
    cudaStream_t aux;       // second stream: flux assembly + all-gather of a group chunk while the next chunk is swept
    cudaEvent_t ev[4];      // main -> aux (chunk swept), aux -> main (exchange done)
};
constexpr int kPairSlots = 4;                  // (num, den) pairs per rank: one per group chunk

static int load_nccl()
{
    if (g_nccl.h) return AE_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("NCCL not found: %s", dlerror()); return AE_CUDA_ERROR; }
#define AE_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name)
    AE_SYM(GetUniqueId, "ncclGetUniqueId");
    AE_SYM(CommInitRank, "ncclCommInitRank");
    AE_SYM(AllReduce, "ncclAllReduce");
    AE_SYM(ReduceScatter, "ncclReduceScatter");
    AE_SYM(AllGather, "ncclAllGather");
    AE_SYM(Broadcast, "ncclBroadcast");
    AE_SYM(GroupStart, "ncclGroupStart");
    AE_SYM(GroupEnd, "ncclGroupEnd");
    AE_SYM(CommDestroy, "ncclCommDestroy");
    AE_SYM(GetErrorString, "ncclGetErrorString");
#undef AE_SYM
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.ReduceScatter || !g_nccl.AllGather || !g_nccl.Broadcast ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.CommDestroy) {
        set_error("NCCL symbols missing");
        return AE_CUDA_ERROR;
    }
    g_nccl.h = h;
    return AE_OK;
}

static int nccl_fail(int rc, const char *what)
{
    set_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
    return AE_CUDA_ERROR;
}
#define AE_NCCL(call)                                    \
    do {                                                 \
        const int rc_ = (call);                          \
        if (rc_) return nccl_fail(rc_, #call);           \
    } while (0)

int comm_rank(const void *comm) { return comm ? ((const Comm *)comm)->rank : 0; }
int comm_size(const void *comm) { return comm ? ((const Comm *)comm)->nranks : 1; }
double *comm_pair_local(void *comm) { return ((Comm *)comm)->pair_local; }
double *comm_pairs(void *comm) { return ((Comm *)comm)->pairs; }
cudaStream_t comm_aux_stream(void *comm) { return ((Comm *)comm)->aux; }
cudaEvent_t comm_event(void *comm, int i) { return ((Comm *)comm)->ev[i]; }

// rows[g][Zp]: sum over ranks of every row, rank r keeping cells [r*Zs, (r+1)*Zs) of each row (in place).
int comm_reduce_scatter_rows(void *comm, double *rows, int G, int64_t Zp, cudaStream_t st)
{
    Comm *c = (Comm *)comm;
    const int64_t Zs = Zp / c->nranks;
    AE_NCCL(g_nccl.GroupStart());
    for (int g = 0; g < G; ++g) {
        double *row = rows + (int64_t)g * Zp;
        AE_NCCL(g_nccl.ReduceScatter(row, row + c->rank * Zs, (size_t)Zs, ncclFloat64, ncclSum, c->nccl, st));
    }
    AE_NCCL(g_nccl.GroupEnd());
    return AE_OK;
}

// rows[g][Zp]: every rank contributes its cell range of every row (in place); optionally the (num, den)
// pairs of a sharded norm travel in the same group.
int comm_all_gather_rows(void *comm, double *rows, int G, int64_t Zp, int with_pairs, cudaStream_t st)
{
    Comm *c = (Comm *)comm;
    const int64_t Zs = Zp / c->nranks;
    AE_NCCL(g_nccl.GroupStart());
    for (int g = 0; g < G && rows; ++g) {
        double *row = rows + (int64_t)g * Zp;
        AE_NCCL(g_nccl.AllGather(row + c->rank * Zs, row, (size_t)Zs, ncclFloat64, c->nccl, st));
    }
    if (with_pairs) AE_NCCL(g_nccl.AllGather(c->pair_local, c->pairs, 2 * kPairSlots, ncclFloat64, c->nccl, st));
    AE_NCCL(g_nccl.GroupEnd());
    return AE_OK;
}

// Balanced contiguous partition of G groups over the ranks: the first G % n ranks own one group more.
void group_range(int G, int nranks, int rank, int *g0, int *gl)
{
    const int base = G / nranks, rem = G % nranks;
    *gl = base + (rank < rem ? 1 : 0);
    *g0 = rank * base + (rank < rem ? rank : rem);
}

// Group-sharded ranks: rows[g][Zp], rank r owns the rows of its groups; every rank ends up with all rows.  The blocks
// have unequal sizes (70 groups over 8 ranks: 9,9,9,9,9,9,8,8) and are not equally spaced, which ncclAllGather cannot
// express.  One broadcast per owner inside an NCCL group does it in place but reached only ~210 GB/s per rank on
// 8 B200 (2.3 ms for the 490 MB of C4); instead the own block is copied into an equal-sized slot of a staging buffer,
// ONE ncclAllGather fills the other slots, and the foreign blocks are copied to their rows (device-to-device, a few
// per cent of the collective).  Optionally the (num, den) pairs of a sharded norm travel in the same NCCL group.
// A rank's groups are exchanged in up to two chunks so that the first can travel while the second is swept: the last
// `tail` groups form chunk 1, the rest chunk 0 (ranks with fewer than 3 groups have one chunk: count = 1 for everybody).
int group_chunks(int G, int nranks) { return G / nranks >= 3 ? 2 : 1; }
void group_chunk_range(int gl, int count, int chunk, int *off, int *len)
{
    const int tail = count == 2 ? 2 : 0;
    if (chunk == 0) { *off = 0; *len = gl - tail; }
    else { *off = gl - tail; *len = tail; }
}

// All-gather of chunk `chunk` (of `count`) of every rank's groups; chunk < 0: whole blocks.
int comm_all_gather_group_chunk(void *comm, double *rows, int G, int64_t Zp, int count, int chunk, int with_pairs, cudaStream_t st)
{
    Comm *c = (Comm *)comm;
    const int glmax = (G + c->nranks - 1) / c->nranks;
    int lmax = glmax;
    if (chunk >= 0) { int o; group_chunk_range(glmax, count, chunk, &o, &lmax); }
    const size_t block = (size_t)lmax * Zp;
    if (c->stage_elems < (size_t)glmax * Zp * c->nranks) {
        if (c->stage) AE_CUDA(cudaFree(c->stage));
        c->stage = nullptr; c->stage_elems = 0;
        AE_CUDA(cudaMalloc((void **)&c->stage, (size_t)glmax * Zp * c->nranks * sizeof(double)));
        c->stage_elems = (size_t)glmax * Zp * c->nranks;
    }
    auto range_of = [&](int r, int *a, int *n) {
        int g0, gl;
        group_range(G, c->nranks, r, &g0, &gl);
        int off = 0, len = gl;
        if (chunk >= 0) group_chunk_range(gl, count, chunk, &off, &len);
        *a = g0 + off; *n = len;
    };
    int a, n;
    range_of(c->rank, &a, &n);
    if (n > 0)
        AE_CUDA(cudaMemcpyAsync(c->stage + block * c->rank, rows + (int64_t)a * Zp, (size_t)n * Zp * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
    AE_NCCL(g_nccl.GroupStart());
    AE_NCCL(g_nccl.AllGather(c->stage + block * c->rank, c->stage, block, ncclFloat64, c->nccl, st));
    if (with_pairs) AE_NCCL(g_nccl.AllGather(c->pair_local, c->pairs, 2 * kPairSlots, ncclFloat64, c->nccl, st));
    AE_NCCL(g_nccl.GroupEnd());
    for (int r = 0; r < c->nranks; ++r) {
        range_of(r, &a, &n);
        if (r != c->rank && n > 0)
            AE_CUDA(cudaMemcpyAsync(rows + (int64_t)a * Zp, c->stage + block * r, (size_t)n * Zp * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
    }
    return AE_OK;
}

int comm_all_gather_groups(void *comm, double *rows, int G, int64_t Zp, int with_pairs, cudaStream_t st)
{
    Comm *c = (Comm *)comm;
    static const char *bc = getenv("AE_COMM_GROUPS_BROADCAST");            // "1": the grouped-broadcast variant (A/B)
    const bool broadcast = bc && bc[0] == '1';
    const int glmax = (G + c->nranks - 1) / c->nranks;
    const size_t block = (size_t)glmax * Zp;
    int g0 = 0, gl = 0;
    group_range(G, c->nranks, c->rank, &g0, &gl);
    if (rows && !broadcast) {
        if (c->stage_elems < block * c->nranks) {
            if (c->stage) AE_CUDA(cudaFree(c->stage));
            c->stage = nullptr; c->stage_elems = 0;
            AE_CUDA(cudaMalloc((void **)&c->stage, block * c->nranks * sizeof(double)));
            c->stage_elems = block * c->nranks;
        }
        if (gl > 0)
            AE_CUDA(cudaMemcpyAsync(c->stage + block * c->rank, rows + (int64_t)g0 * Zp, (size_t)gl * Zp * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
    }
    AE_NCCL(g_nccl.GroupStart());
    if (rows && !broadcast) AE_NCCL(g_nccl.AllGather(c->stage + block * c->rank, c->stage, block, ncclFloat64, c->nccl, st));
    for (int r = 0; r < c->nranks && rows && broadcast; ++r) {
        int a, n;
        group_range(G, c->nranks, r, &a, &n);
        double *blk = rows + (int64_t)a * Zp;
        if (n > 0) AE_NCCL(g_nccl.Broadcast(blk, blk, (size_t)n * Zp, ncclFloat64, r, c->nccl, st));
    }
    if (with_pairs) AE_NCCL(g_nccl.AllGather(c->pair_local, c->pairs, 2 * kPairSlots, ncclFloat64, c->nccl, st));
    AE_NCCL(g_nccl.GroupEnd());
    for (int r = 0; r < c->nranks && rows && !broadcast; ++r) {
        int a, n;
        group_range(G, c->nranks, r, &a, &n);
        if (r != c->rank && n > 0)
            AE_CUDA(cudaMemcpyAsync(rows + (int64_t)a * Zp, c->stage + block * r, (size_t)n * Zp * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
    }
    return AE_OK;
}

}  // namespace ae

using namespace ae;

extern "C" void ae_group_range(int32_t G, int32_t nranks, int32_t rank, int32_t *g0, int32_t *gl)
{
    int a = 0, b = 0;
    group_range(G, nranks > 0 ? nranks : 1, rank, &a, &b);
    if (g0) *g0 = a;
    if (gl) *gl = b;
}

extern "C" int ae_comm_gather_groups(void *comm, double *rows, int32_t G, int64_t Zp, void *stream)
{
    if (!comm || !g_nccl.h) { set_error("ae_comm_gather_groups: no communicator"); return AE_BAD_ARG; }
    return comm_all_gather_groups(comm, rows, G, Zp, 0, (cudaStream_t)stream);
}

extern "C" int ae_comm_unique_id(void *id128)
{
    AE_TRY(load_nccl());
    ncclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return AE_OK;
}

extern "C" int ae_comm_init(void **comm, int32_t rank, int32_t nranks, const void *id128)
{
    AE_TRY(load_nccl());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t nc = nullptr;
    const int rc = g_nccl.CommInitRank(&nc, nranks, id, rank);
    if (rc) return nccl_fail(rc, "ncclCommInitRank");
    Comm *c = new Comm{nc, rank, nranks, nullptr, nullptr, nullptr, 0, nullptr, {nullptr, nullptr, nullptr, nullptr}};
    AE_CUDA(cudaMalloc((void **)&c->pair_local, 2 * kPairSlots * sizeof(double)));
    AE_CUDA(cudaMalloc((void **)&c->pairs, 2 * kPairSlots * sizeof(double) * nranks));
    AE_CUDA(cudaMemset(c->pair_local, 0, 2 * kPairSlots * sizeof(double)));
    AE_CUDA(cudaMemset(c->pairs, 0, 2 * kPairSlots * sizeof(double) * nranks));
    AE_CUDA(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) AE_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
    *comm = c;
    return AE_OK;
}

extern "C" int ae_comm_allreduce_sum(void *comm, double *buf, int64_t count, void *stream)
{
    if (!comm || !g_nccl.h) { set_error("ae_comm_allreduce_sum: no communicator"); return AE_BAD_ARG; }
    AE_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, ((Comm *)comm)->nccl, (cudaStream_t)stream));
    return AE_OK;
}

extern "C" int ae_comm_gather_rows(void *comm, double *rows, int32_t G, int64_t Zp, void *stream)
{
    if (!comm || !g_nccl.h) { set_error("ae_comm_gather_rows: no communicator"); return AE_BAD_ARG; }
    return comm_all_gather_rows(comm, rows, G, Zp, 0, (cudaStream_t)stream);
}

extern "C" int ae_comm_destroy(void *comm)
{
    if (!comm || !g_nccl.h) return AE_OK;
    Comm *c = (Comm *)comm;
    const int rc = g_nccl.CommDestroy(c->nccl);
    cudaFree(c->pair_local);
    cudaFree(c->pairs);
    cudaFree(c->stage);
    if (c->aux) cudaStreamDestroy(c->aux);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
    if (rc) return nccl_fail(rc, "ncclCommDestroy");
    return AE_OK;
}
