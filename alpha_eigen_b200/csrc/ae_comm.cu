// NCCL plumbing for the angle-sharded multi-GPU path (one process per GPU).
//
// The reference is single-process; sharding the (angle, group) rows over ranks adds exactly one
// exchange per source iteration: the sum over ranks of the partial scalar flux (group.py:77 summed
// over ALL angles).  NCCL is resolved at run time from the copy torch already loaded (dlopen by
// soname), so this library has no link-time NCCL dependency and loads on CPU-only hosts.
#include <dlfcn.h>
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
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static Nccl g_nccl;

static int load_nccl()
{
    if (g_nccl.h) return AE_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("NCCL not found: %s", dlerror()); return AE_CUDA_ERROR; }
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
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

}  // namespace ae

using namespace ae;

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
    ncclComm_t c = nullptr;
    const int rc = g_nccl.CommInitRank(&c, nranks, id, rank);
    if (rc) return nccl_fail(rc, "ncclCommInitRank");
    *comm = c;
    return AE_OK;
}

extern "C" int ae_comm_allreduce_sum(void *comm, double *buf, int64_t count, void *stream)
{
    if (!comm || !g_nccl.h) { set_error("ae_comm_allreduce_sum: no communicator"); return AE_BAD_ARG; }
    const int rc = g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream);
    if (rc) return nccl_fail(rc, "ncclAllReduce");
    return AE_OK;
}

extern "C" int ae_comm_destroy(void *comm)
{
    if (!comm || !g_nccl.h) return AE_OK;
    const int rc = g_nccl.CommDestroy((ncclComm_t)comm);
    if (rc) return nccl_fail(rc, "ncclCommDestroy");
    return AE_OK;
}
