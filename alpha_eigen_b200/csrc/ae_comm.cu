// NCCL plumbing for the angle-sharded multi-GPU path (one process per GPU).
//
// The reference is single-process; sharding the (angle, group) rows over ranks adds exactly one
// exchange per source iteration: the sum over ranks of the partial scalar flux (group.py:77 summed
// over ALL angles).  NCCL is resolved at run time from the copy torch already loaded (dlopen by
// soname), so this library has no link-time NCCL dependency and loads on CPU-only hosts.
//
// Group-sharded ranks exchange whole flux rows: by default one ncclAllGather through a staging buffer.  The opt-in
// alternative (AE_COMM_PEER_COPY=1, peer_exchange below) does not go through NCCL: every rank owns one IPC-shared
// buffer, peers push their rows into it with copy engines over NVLink and signal with stream memory operations
// (cuStreamWaitValue64 on a flag the pusher's copy engine writes), so the exchange uses NO SM and can run under a
// persistent sweep kernel that occupies all of them.
#include <cuda.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include "ae_common.cuh"

namespace ae {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt8 = 0, ncclFloat64 = 8, ncclSum = 0 };

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

constexpr int kMaxPeers = 16;                  // ranks the peer exchange supports (one NVSwitch domain)
constexpr size_t kXData = 0, kXAck = 128, kXPairs = 1024, kXHeader = 8192;      // byte offsets in an exchange buffer

// IPC-shared exchange buffer of one rank: [data flags | ack flags | (num, den) pairs x 2 parities | rows [G][Zp]].
struct Exchange {
    int state = 0;                              // 0 not set up, 1 ready, -1 unavailable (NCCL does the exchange)
    unsigned char *base = nullptr;              // this rank's buffer
    unsigned char *peer[kMaxPeers] = {};        // every rank's buffer mapped into this process (peer[rank] == base)
    size_t elems = 0;                           // doubles of row space
    unsigned long long epoch = 0;               // chunk exchanges issued so far (value of the data flags)
    unsigned long long round = 0;               // whole exchanges issued so far (value of the ack flags)
    unsigned long long *src = nullptr;          // device [16]: the words the flag copies read (written by stream memops)
};

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
    Exchange x;
    double *pairs_own;      // the cudaMalloc'ed pairs array (`pairs` points into the exchange buffer once that exists)
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

// ---- peer exchange: copy engines + stream memory operations, no SM ------------------------------------------------
struct Driver {
    int state = 0;          // 0 untried, 1 ok, -1 missing
    CUresult (*WaitValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
    CUresult (*WriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
};
static Driver g_drv;

static bool load_driver()
{
    if (g_drv.state) return g_drv.state > 0;
    cudaDriverEntryPointQueryResult q1, q2;
    void *a = nullptr, *b = nullptr;
    const bool ok = cudaGetDriverEntryPoint("cuStreamWaitValue64", &a, cudaEnableDefault, &q1) == cudaSuccess &&
                    cudaGetDriverEntryPoint("cuStreamWriteValue64", &b, cudaEnableDefault, &q2) == cudaSuccess &&
                    q1 == cudaDriverEntryPointSuccess && q2 == cudaDriverEntryPointSuccess && a && b;
    cudaGetLastError();
    *(void **)(&g_drv.WaitValue64) = a;
    *(void **)(&g_drv.WriteValue64) = b;
    g_drv.state = ok ? 1 : -1;
    return ok;
}

#define AE_DRV(call)                                                             \
    do {                                                                         \
        const CUresult rc_ = (call);                                             \
        if (rc_ != CUDA_SUCCESS) { set_error("%s failed: CUresult %d", #call, (int)rc_); return AE_CUDA_ERROR; } \
    } while (0)

// Collective when the exchange is up: nobody frees a buffer a peer still has mapped.
static void exchange_release(Comm *c, cudaStream_t st)
{
    Exchange &x = c->x;
    for (int r = 0; r < c->nranks && r < kMaxPeers; ++r)
        if (x.peer[r] && r != c->rank) cudaIpcCloseMemHandle(x.peer[r]);
    if (x.state > 0) {
        g_nccl.AllReduce(c->pair_local, c->pair_local, 1, ncclFloat64, ncclSum, c->nccl, st);
        cudaStreamSynchronize(st);
    }
    if (x.base) cudaFree(x.base);
    if (x.src) cudaFree(x.src);
    c->pairs = c->pairs_own;
    x = Exchange();
}

// Collective over the communicator (every rank calls it with the same `elems`): allocate this rank's buffer, trade IPC
// handles, map the peers.  Any failure on any rank leaves the exchange unavailable on all of them (NCCL carries on).
static int exchange_setup(Comm *c, size_t elems, cudaStream_t st)
{
    Exchange &x = c->x;
    if (x.state < 0 || (x.state > 0 && x.elems >= elems)) return AE_OK;
    // Opt-in (AE_COMM_PEER_COPY=1).  Measured on C4 (70 groups, 1M cells: 560 MB of flux): the exchange alone takes 1.40 ms
    // on 8 B200 against 1.13 ms for the staged ncclAllGather (7 serial 72 MB pushes per rank reach ~460 GB/s), and inside
    // the step 3.1 ms (10.5 against 8.5 ms per step); on 2 GPUs it is the faster one (28.0 against 28.2 ms per step).
    static const char *env = getenv("AE_COMM_PEER_COPY");
    if (!(env && env[0] == '1') || c->nranks > kMaxPeers || c->nranks < 2) { x.state = -1; return AE_OK; }
    AE_CUDA(cudaDeviceSynchronize());
    if (x.state > 0) {
        // a larger problem: nobody may still be pushing into the old buffers when they go away
        double *z = c->pair_local;
        AE_NCCL(g_nccl.AllReduce(z, z, 1, ncclFloat64, ncclSum, c->nccl, st));
        AE_CUDA(cudaStreamSynchronize(st));
        exchange_release(c, st);
    }
    int fail = load_driver() ? 0 : 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (!fail && cudaMalloc((void **)&x.base, kXHeader + elems * sizeof(double)) != cudaSuccess) fail = 1;
    if (!fail && cudaMalloc((void **)&x.src, 16 * sizeof(unsigned long long)) != cudaSuccess) fail = 1;
    if (!fail && cudaMemset(x.base, 0, kXHeader) != cudaSuccess) fail = 1;
    if (!fail && cudaMemset(x.src, 0, 16 * sizeof(unsigned long long)) != cudaSuccess) fail = 1;
    if (!fail && cudaIpcGetMemHandle(&mine, x.base) != cudaSuccess) fail = 1;
    cudaGetLastError();
    // handles (and a failure count) travel through NCCL: [nranks][handle | fail]
    const size_t slot = sizeof(cudaIpcMemHandle_t) + 8;
    unsigned char *dev = nullptr, host[kMaxPeers * (sizeof(cudaIpcMemHandle_t) + 8)];
    AE_CUDA(cudaMalloc((void **)&dev, slot * c->nranks));
    memset(host, 0, sizeof(host));
    memcpy(host + slot * c->rank, &mine, sizeof(mine));
    host[slot * c->rank + sizeof(mine)] = (unsigned char)fail;
    AE_CUDA(cudaMemcpyAsync(dev + slot * c->rank, host + slot * c->rank, slot, cudaMemcpyHostToDevice, st));
    AE_NCCL(g_nccl.AllGather(dev + slot * c->rank, dev, slot, ncclInt8, c->nccl, st));
    AE_CUDA(cudaMemcpyAsync(host, dev, slot * c->nranks, cudaMemcpyDeviceToHost, st));
    AE_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < c->nranks; ++r) fail += host[slot * r + sizeof(mine)];
    for (int r = 0; r < c->nranks && !fail; ++r) {
        if (r == c->rank) { x.peer[r] = x.base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, host + slot * r, sizeof(h));
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { fail = 1; cudaGetLastError(); break; }
        x.peer[r] = (unsigned char *)ptr;
    }
    // the mapping can fail on one rank only: agree once more
    double *z = (double *)dev;
    const double mine_failed = fail ? 1.0 : 0.0;
    double all_failed = 0.0;
    AE_CUDA(cudaMemcpyAsync(z, &mine_failed, sizeof(double), cudaMemcpyHostToDevice, st));
    AE_NCCL(g_nccl.AllReduce(z, z, 1, ncclFloat64, ncclSum, c->nccl, st));
    AE_CUDA(cudaMemcpyAsync(&all_failed, z, sizeof(double), cudaMemcpyDeviceToHost, st));
    AE_CUDA(cudaStreamSynchronize(st));
    AE_CUDA(cudaFree(dev));
    if (all_failed > 0.0) { exchange_release(c, st); x.state = -1; return AE_OK; }
    x.elems = elems; x.epoch = 0; x.round = 0; x.state = 1;
    return AE_OK;
}

// One chunk of the all-gather of group rows through the peer buffers, everything on stream `st`:
//   first chunk of an exchange: wait until every peer has drained the previous exchange out of ITS view of my rows
//   push my rows of the chunk (and, with the last chunk, my pairs) into every peer's buffer, then my data flag
//   wait for every peer's data flag, copy their rows from my buffer to `rows`
//   last chunk: tell every peer that my buffer is free again (ack flag)
// Copies are cudaMemcpyAsync between device pointers (copy engines), flags are 8-byte copies of a word a stream memory
// operation wrote, waits are cuStreamWaitValue64: no kernel anywhere.
static int peer_exchange(Comm *c, double *rows, int G, int64_t Zp, int count, int chunk, int with_pairs, cudaStream_t st)
{
    Exchange &x = c->x;
    const bool first = chunk <= 0, last = chunk < 0 || chunk == count - 1;
    const int me = c->rank, n = c->nranks;
    auto range_of = [&](int r, int64_t *a, int64_t *len) {
        int g0, gl;
        group_range(G, n, r, &g0, &gl);
        int off = 0, l = gl;
        if (chunk >= 0) group_chunk_range(gl, count, chunk, &off, &l);
        *a = (int64_t)(g0 + off) * Zp; *len = rows ? (int64_t)l * Zp : 0;
    };
    auto flag_at = [&](int r, size_t section, int slot) { return (CUdeviceptr)(x.peer[r] + section + 8 * (size_t)slot); };
    double *xrows = (double *)(x.base + kXHeader);
    if (first)
        for (int k = 1; k < n; ++k)
            AE_DRV(g_drv.WaitValue64((CUstream)st, flag_at(me, kXAck, (me + k) % n), x.round, CU_STREAM_WAIT_VALUE_GEQ));
    const unsigned long long e = ++x.epoch;
    const size_t pair_bytes = 2 * kPairSlots * sizeof(double);
    const size_t parity = (size_t)(x.round & 1) * kMaxPeers * pair_bytes;       // pairs of consecutive exchanges alternate
    int64_t a, len;
    range_of(me, &a, &len);
    for (int k = 1; k < n; ++k) {
        const int p = (me + k) % n;
        if (len > 0)
            AE_CUDA(cudaMemcpyAsync(x.peer[p] + kXHeader + a * sizeof(double), rows + a, len * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (with_pairs && last)
            AE_CUDA(cudaMemcpyAsync(x.peer[p] + kXPairs + parity + me * pair_bytes, c->pair_local, pair_bytes, cudaMemcpyDeviceToDevice, st));
    }
    if (with_pairs && last)
        AE_CUDA(cudaMemcpyAsync(x.base + kXPairs + parity + me * pair_bytes, c->pair_local, pair_bytes, cudaMemcpyDeviceToDevice, st));
    unsigned long long *word = x.src + (e & 7);
    AE_DRV(g_drv.WriteValue64((CUstream)st, (CUdeviceptr)word, e, CU_STREAM_WRITE_VALUE_DEFAULT));
    for (int k = 1; k < n; ++k)
        AE_CUDA(cudaMemcpyAsync((void *)flag_at((me + k) % n, kXData, me), word, 8, cudaMemcpyDeviceToDevice, st));
    for (int k = 1; k < n; ++k) {
        const int p = (me + n - k) % n;
        AE_DRV(g_drv.WaitValue64((CUstream)st, flag_at(me, kXData, p), e, CU_STREAM_WAIT_VALUE_GEQ));
        range_of(p, &a, &len);
        if (len > 0) AE_CUDA(cudaMemcpyAsync(rows + a, xrows + a, len * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    if (last) {
        if (with_pairs) c->pairs = (double *)(x.base + kXPairs + parity);
        const unsigned long long r = ++x.round;
        unsigned long long *ack = x.src + 8 + (r & 7);
        AE_DRV(g_drv.WriteValue64((CUstream)st, (CUdeviceptr)ack, r, CU_STREAM_WRITE_VALUE_DEFAULT));
        for (int k = 1; k < n; ++k)
            AE_CUDA(cudaMemcpyAsync((void *)flag_at((me + k) % n, kXAck, me), ack, 8, cudaMemcpyDeviceToDevice, st));
    }
    return AE_OK;
}

bool comm_peer_exchange_ready(void *comm, int G, int64_t Zp, cudaStream_t st)
{
    Comm *c = (Comm *)comm;
    if (exchange_setup(c, (size_t)G * Zp, st) != AE_OK) return false;
    return c->x.state > 0;
}

// All-gather of chunk `chunk` (of `count`) of every rank's groups; chunk < 0: whole blocks.
int comm_all_gather_group_chunk(void *comm, double *rows, int G, int64_t Zp, int count, int chunk, int with_pairs, cudaStream_t st)
{
    Comm *c = (Comm *)comm;
    AE_TRY(exchange_setup(c, (size_t)G * Zp, st));
    if (c->x.state > 0) return peer_exchange(c, rows, G, Zp, count, chunk, with_pairs, st);
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
    if (!broadcast) {
        AE_TRY(exchange_setup(c, (size_t)G * Zp, st));
        if (c->x.state > 0) return peer_exchange(c, rows, G, Zp, 1, -1, with_pairs, st);
    }
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

extern "C" int32_t ae_comm_exchange_mode(void *comm)
{
    if (!comm) return -1;
    const int st = ((Comm *)comm)->x.state;
    return st > 0 ? 1 : (st < 0 ? 0 : -1);
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
    Comm *c = new Comm{nc, rank, nranks, nullptr, nullptr, nullptr, 0, nullptr, {nullptr, nullptr, nullptr, nullptr}, Exchange(), nullptr};
    AE_CUDA(cudaMalloc((void **)&c->pair_local, 2 * kPairSlots * sizeof(double)));
    AE_CUDA(cudaMalloc((void **)&c->pairs, 2 * kPairSlots * sizeof(double) * nranks));
    AE_CUDA(cudaMemset(c->pair_local, 0, 2 * kPairSlots * sizeof(double)));
    AE_CUDA(cudaMemset(c->pairs, 0, 2 * kPairSlots * sizeof(double) * nranks));
    c->pairs_own = c->pairs;
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
    cudaDeviceSynchronize();
    exchange_release(c, nullptr);
    const int rc = g_nccl.CommDestroy(c->nccl);
    cudaFree(c->pair_local);
    cudaFree(c->pairs_own);
    cudaFree(c->stage);
    if (c->aux) cudaStreamDestroy(c->aux);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
    if (rc) return nccl_fail(rc, "ncclCommDestroy");
    return AE_OK;
}
