// vg_comm_*: the gradient exchange of the data-parallel train step behind the C ABI.
//
// Replaces what tf.distribute.MirroredStrategy does inside `optimizer.minimize` (vangan.py:426-438: SUM all-reduce of every
// network's gradients across replicas before clip + Adam) and `strategy.reduce(SUM)` of the result dict (vangan.py:459-473).
// One process per GPU; the communicator is NCCL over NVLink / NVSwitch.  The library never links NCCL: the symbols are resolved
// at run time from the libnccl.so.2 that is already in the process (torch's bundled copy) or on the loader path, so a single-GPU
// user needs no NCCL at all.  All collectives are enqueued on the caller's stream -- a dedicated communication stream ordered by
// events, which is also how they are captured into the CUDA graph of the step next to the backward sweeps they overlap.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* names[] = {getenv("VG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) return a;
#define VG_SYM(field, name)                                       \
    *(void**)(&a.field) = dlsym(a.handle, name);                  \
    if (!a.field) return a;
    VG_SYM(GetUniqueId, "ncclGetUniqueId")
    VG_SYM(CommInitRank, "ncclCommInitRank")
    VG_SYM(CommDestroy, "ncclCommDestroy")
    VG_SYM(AllReduce, "ncclAllReduce")
    VG_SYM(GroupStart, "ncclGroupStart")
    VG_SYM(GroupEnd, "ncclGroupEnd")
    VG_SYM(GetErrorString, "ncclGetErrorString")
    VG_SYM(GetVersion, "ncclGetVersion")
#undef VG_SYM
    a.ok = true;
    return a;
}

}  // namespace

struct vg_comm {
    ncclComm_t comm;
    int world, rank, device;
    unsigned long long collectives;   // all-reduce messages enqueued so far (buckets count one each)
};

#define VG_NCCL(call)                                                                            \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess && r_ != ncclInProgress) {                                         \
            if (getenv("VG_DEBUG")) fprintf(stderr, "[vg_comm] %s: %s\n", #call, api().GetErrorString(r_)); \
            return VG_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

extern "C" {

int vg_comm_available(void) { return api().ok ? 1 : 0; }

int vg_comm_nccl_version(void) {
    int v = 0;
    if (!api().ok || api().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int vg_comm_unique_id(void* id128) {
    VG_REQUIRE(id128);
    if (!api().ok) return VG_ERR_UNSUPPORTED;
    ncclUniqueId id;
    VG_NCCL(api().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return VG_OK;
}

int vg_comm_init(vg_comm** out, const void* id128, int world, int rank, int device) {
    VG_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world && device >= 0);
    if (!api().ok) return VG_ERR_UNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return VG_ERR_CUDA;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    vg_comm* c = (vg_comm*)calloc(1, sizeof(vg_comm));
    if (!c) return VG_ERR_CUDA;
    c->world = world; c->rank = rank; c->device = device;
    ncclResult_t r = api().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        if (getenv("VG_DEBUG")) fprintf(stderr, "[vg_comm] ncclCommInitRank: %s\n", api().GetErrorString(r));
        free(c);
        return VG_ERR_CUDA;
    }
    *out = c;
    return VG_OK;
}

int vg_comm_world(const vg_comm* c) { return c ? c->world : 0; }
int vg_comm_rank(const vg_comm* c) { return c ? c->rank : -1; }
unsigned long long vg_comm_collectives(const vg_comm* c) { return c ? c->collectives : 0ull; }

// In-place SUM all-reduce of a flat fp32 gradient buffer, cut into buckets of `bucket_elems` elements (one NCCL message each, grouped
// into one launch) so that a consumer can be ordered after individual buckets later on; bucket_elems <= 0 = one message.
int vg_comm_allreduce_bucket(vg_comm* c, float* buf, long long count, long long bucket_elems, void* stream) {
    VG_REQUIRE(c && buf && count > 0);
    if (c->world == 1) return VG_OK;
    if (bucket_elems <= 0 || bucket_elems > count) bucket_elems = count;
    VG_NCCL(api().GroupStart());
    for (long long off = 0; off < count; off += bucket_elems) {
        const long long n = count - off < bucket_elems ? count - off : bucket_elems;
        ncclResult_t r = api().AllReduce(buf + off, buf + off, (size_t)n, ncclFloat32, ncclSum, c->comm, (cudaStream_t)stream);
        if (r != ncclSuccess && r != ncclInProgress) {
            api().GroupEnd();
            return VG_ERR_CUDA;
        }
        c->collectives++;
    }
    VG_NCCL(api().GroupEnd());
    return VG_OK;
}

// strategy.reduce(SUM) of the result dict (vangan.py:459-473): n fp64 scalars in device memory, in place
int vg_comm_reduce_scalars(vg_comm* c, double* vals, int n, void* stream) {
    VG_REQUIRE(c && vals && n > 0);
    if (c->world == 1) return VG_OK;
    VG_NCCL(api().AllReduce(vals, vals, (size_t)n, ncclFloat64, ncclSum, c->comm, (cudaStream_t)stream));
    c->collectives++;
    return VG_OK;
}

int vg_comm_destroy(vg_comm* c) {
    if (!c) return VG_OK;
    if (api().ok && c->comm) api().CommDestroy(c->comm);
    free(c);
    return VG_OK;
}

}  // extern "C"
