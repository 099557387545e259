// comm.cu -- see comm.h.  NCCL entry points are resolved with dlopen/dlsym ("libnccl.so.2": the copy already mapped into
// the process when the host is PyTorch, else the system one); the few types needed are declared here to the NCCL 2.x ABI.
#include "comm.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace sbmf {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;          // ncclSuccess = 0
enum { ncclFloat32 = 7, ncclFloat64 = 8 };   // nccl.h ncclDataType_t
enum { ncclSum = 0 };                        // nccl.h ncclRedOp_t

struct Api {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
Api g_api;

bool load(std::string& err)
{
    if (g_api.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) {
        err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
#define SYM(field, name)                                                        \
    *(void**)(&g_api.field) = dlsym(g_api.lib, name);                           \
    if (!g_api.field) {                                                         \
        err = std::string("libnccl: missing symbol ") + name;                   \
        return false;                                                           \
    }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(Broadcast, "ncclBroadcast") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_api.ok = true;
    return true;
}

#define NC(call)                                                                \
    do {                                                                        \
        ncclResult_t r_ = (call);                                               \
        if (r_ != 0) {                                                          \
            err = std::string(#call) + ": " + g_api.GetErrorString(r_);         \
            return -1;                                                          \
        }                                                                       \
    } while (0)
// inside an open ncclGroupStart: close the group before reporting, otherwise the next collective hangs instead of failing
#define NCG(call)                                                               \
    do {                                                                        \
        ncclResult_t r_ = (call);                                               \
        if (r_ != 0) {                                                          \
            err = std::string(#call) + ": " + g_api.GetErrorString(r_);         \
            g_api.GroupEnd();                                                   \
            return -1;                                                          \
        }                                                                       \
    } while (0)
}  // namespace

int comm_unique_id(uint8_t out[128], std::string& err)
{
    if (!load(err)) return -1;
    ncclUniqueId id;
    NC(g_api.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return 0;
}

int comm_init(Comm& c, const uint8_t idb[128], int rank, int world, std::string& err)
{
    if (!load(err)) return -1;
    ncclUniqueId id;
    memcpy(id.internal, idb, 128);
    ncclComm_t comm = nullptr;
    NC(g_api.CommInitRank(&comm, world, id, rank));
    c.nccl = comm;
    c.rank = rank;
    c.world = world;
    return 0;
}

void comm_destroy(Comm& c)
{
    if (c.nccl && g_api.ok) g_api.CommDestroy((ncclComm_t)c.nccl);
    c.nccl = nullptr;
}

int comm_allreduce_sum_f64(Comm& c, double* buf, size_t count, cudaStream_t st, std::string& err)
{
    NC(g_api.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)c.nccl, st));
    return 0;
}

// In-place all-gather-v of nseg strided segments, point-to-point: every rank sends its own part of every segment straight
// to every peer (one grouped NCCL launch, all NVLink links busy at once; measured faster than one broadcast per root).
template <typename T>
static int allgatherv(Comm& c, T* base, size_t stride, int nseg, const size_t* offsets, const size_t* counts, int dtype, cudaStream_t st,
                      std::string& err)
{
    NC(g_api.GroupStart());
    for (int q = 0; q < c.world; ++q) {
        if (q == c.rank) continue;
        for (int s = 0; s < nseg; ++s) {
            if (counts[c.rank]) NCG(g_api.Send(base + (size_t)s * stride + offsets[c.rank], counts[c.rank], dtype, q, (ncclComm_t)c.nccl, st));
            if (counts[q]) NCG(g_api.Recv(base + (size_t)s * stride + offsets[q], counts[q], dtype, q, (ncclComm_t)c.nccl, st));
        }
    }
    NC(g_api.GroupEnd());
    return 0;
}

int comm_allgatherv_f32(Comm& c, float* buf, const size_t* offsets, const size_t* counts, cudaStream_t st, std::string& err)
{
    return allgatherv<float>(c, buf, 0, 1, offsets, counts, ncclFloat32, st, err);
}
int comm_allgatherv_f64(Comm& c, double* buf, const size_t* offsets, const size_t* counts, cudaStream_t st, std::string& err)
{
    return allgatherv<double>(c, buf, 0, 1, offsets, counts, ncclFloat64, st, err);
}
int comm_allgatherv_strided_f32(Comm& c, float* base, size_t stride, int nseg, const size_t* offsets, const size_t* counts, cudaStream_t st,
                                std::string& err)
{
    return allgatherv<float>(c, base, stride, nseg, offsets, counts, ncclFloat32, st, err);
}

int comm_group_begin(std::string& err)
{
    if (!load(err)) return -1;
    NC(g_api.GroupStart());
    return 0;
}
int comm_group_end(std::string& err)
{
    NC(g_api.GroupEnd());
    return 0;
}

int comm_alltoallv_f32(Comm& c, const float* send, const size_t* send_off, const size_t* send_cnt, float* recv, const size_t* recv_off,
                       const size_t* recv_cnt, cudaStream_t st, std::string& err)
{
    NC(g_api.GroupStart());
    for (int q = 0; q < c.world; ++q) {
        if (send_cnt[q]) NCG(g_api.Send(send + send_off[q], send_cnt[q], ncclFloat32, q, (ncclComm_t)c.nccl, st));
        if (recv_cnt[q]) NCG(g_api.Recv(recv + recv_off[q], recv_cnt[q], ncclFloat32, q, (ncclComm_t)c.nccl, st));
    }
    NC(g_api.GroupEnd());
    return 0;
}

}  // namespace sbmf
