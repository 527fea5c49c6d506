// comm.cu -- the collective of the hot path inside the C ABI: one NCCL communicator per context
// (one process per GPU), used only for the small shared blocks of the fit:
//   cc_reproj_jtj's 21-double block [J'J_ii | J'r_i | sum r^2]  (the reduction the reference runs
//   serially over views, src/buildcalibrations.jl:28-31,60-65), the LM Schur share and step norms.
// Frames and points never cross GPUs.
//
// libnccl is resolved at run time (dlopen "libnccl.so.2"): the library loads and every single-GPU
// entry point works on a box without NCCL; a process that already has NCCL mapped (torch) shares
// that copy.  Bootstrap (carrying the 128-byte unique id from rank 0 to the other ranks) belongs
// to the caller: MPI, a file, torch.distributed -- it is not on the data path.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace cc {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return CC_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(CC_ERR_UNSUPPORTED, "NCCL not available: %s", dlerror());
    NcclApi a;
    a.handle = h;
    *(void**)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&a.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
    *(void**)&a.GetVersion = dlsym(h, "ncclGetVersion");
    *(void**)&a.CommInitAll = dlsym(h, "ncclCommInitAll");
    *(void**)&a.GroupStart = dlsym(h, "ncclGroupStart");
    *(void**)&a.GroupEnd = dlsym(h, "ncclGroupEnd");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString)
        return set_error(CC_ERR_UNSUPPORTED, "libnccl.so.2 lacks an expected symbol");
    g_nccl = a;
    return CC_OK;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    return set_error(CC_ERR_CUDA, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
}

// sum `count` doubles over the ranks of the context's communicator, in place, on `st`;
// a context without a communicator is a world of one: nothing to do
int comm_allreduce_sum(cc_ctx* ctx, double* buf, size_t count, cudaStream_t st) {
    if (!ctx->nccl_comm || ctx->comm_nranks <= 1 || count == 0) return CC_OK;
    const ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum,
                                            static_cast<ncclComm_t>(ctx->nccl_comm), st);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    ctx->collectives++;
    return CC_OK;
}

void comm_free(cc_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
    ctx->comm_nranks = 1;
    ctx->comm_rank = 0;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_comm_unique_id(void* id128) {
    CC_REQUIRE(id128 != nullptr, "id buffer is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == CC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return CC_OK;
}

int cc_comm_init_rank(cc_ctx* ctx, int nranks, int rank, const void* id128) {
    CC_REQUIRE(ctx != nullptr, "ctx is NULL");
    CC_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    CC_REQUIRE(ctx->nccl_comm == nullptr, "this context already has a communicator");
    if (nranks == 1) { ctx->comm_nranks = 1; ctx->comm_rank = 0; return CC_OK; }
    CC_REQUIRE(id128 != nullptr, "unique id is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    int prev = -1;
    CC_CUDA(cudaGetDevice(&prev));
    CC_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->comm_nranks = nranks;
    ctx->comm_rank = rank;
    return CC_OK;
}

int cc_comm_destroy(cc_ctx* ctx) {
    CC_REQUIRE(ctx != nullptr, "ctx is NULL");
    comm_free(ctx);
    return CC_OK;
}

int cc_comm_size(const cc_ctx* ctx, int* nranks, int* rank) {
    CC_REQUIRE(ctx != nullptr, "ctx is NULL");
    if (nranks) *nranks = ctx->comm_nranks > 0 ? ctx->comm_nranks : 1;
    if (rank) *rank = ctx->comm_rank;
    return CC_OK;
}

int cc_allreduce_shared(cc_ctx* ctx, double* buf, size_t count, void* stream) {
    CC_REQUIRE(ctx != nullptr, "ctx is NULL");
    CC_REQUIRE(buf != nullptr || count == 0, "buffer is NULL");
    return comm_allreduce_sum(ctx, buf, count, static_cast<cudaStream_t>(stream));
}

// ---- one process, several GPUs ("cc_ctx_create(ndev, devs): owns streams, NCCL comms, scratch").
// A single Julia session that drives all GPUs of a box creates one context per device, joined in one
// communicator; the shared block is then summed with ONE grouped call.  Entry points that
// synchronise internally on a collective (cc_lm_fit_f64) must be called from one host thread per
// context; everything else is asynchronous and can be issued from one thread, device after device.
int cc_ctx_create_group(int ndev, const int* devs, cc_ctx** ctxs) {
    CC_REQUIRE(ndev >= 1 && devs && ctxs, "bad device list");
    for (int i = 0; i < ndev; ++i) {
        ctxs[i] = nullptr;
        for (int j = 0; j < i; ++j) CC_REQUIRE(devs[i] != devs[j], "a device appears twice in the list");
    }
    int rc = CC_OK;
    for (int i = 0; i < ndev && !rc; ++i) rc = cc_ctx_create(devs[i], &ctxs[i]);
    if (!rc && ndev > 1) {
        rc = nccl_load();
        if (!rc && (!g_nccl.CommInitAll || !g_nccl.GroupStart || !g_nccl.GroupEnd))
            rc = set_error(CC_ERR_UNSUPPORTED, "libnccl.so.2 lacks ncclCommInitAll / ncclGroupStart");
        if (!rc) {
            ncclComm_t comms[64];
            if (ndev > 64) rc = set_error(CC_ERR_INVALID_ARG, "at most 64 devices per group");
            else {
                int prev = -1;
                cudaGetDevice(&prev);
                const ncclResult_t r = g_nccl.CommInitAll(comms, ndev, devs);
                if (prev >= 0) cudaSetDevice(prev);
                if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitAll");
                else
                    for (int i = 0; i < ndev; ++i) { ctxs[i]->nccl_comm = comms[i]; ctxs[i]->comm_nranks = ndev; ctxs[i]->comm_rank = i; }
            }
        }
    }
    if (rc)
        for (int i = 0; i < ndev; ++i)
            if (ctxs[i]) { cc_ctx_destroy(ctxs[i]); ctxs[i] = nullptr; }
    return rc;
}

int cc_ctx_destroy_group(int ndev, cc_ctx** ctxs) {
    CC_REQUIRE(ndev >= 0 && (ndev == 0 || ctxs), "bad context list");
    int rc = CC_OK;
    for (int i = 0; i < ndev; ++i)
        if (ctxs[i]) { const int r = cc_ctx_destroy(ctxs[i]); ctxs[i] = nullptr; if (r && !rc) rc = r; }
    return rc;
}

// bufs[i] (device memory of ctxs[i]'s GPU) <- sum over i, in place, each on streams[i] (NULL: default streams)
int cc_allreduce_shared_group(cc_ctx* const* ctxs, int ndev, double* const* bufs, size_t count, void* const* streams) {
    CC_REQUIRE(ndev >= 1 && ctxs && bufs, "bad arguments");
    for (int i = 0; i < ndev; ++i) {
        CC_REQUIRE(ctxs[i] != nullptr && (bufs[i] != nullptr || count == 0), "NULL context / buffer");
        CC_REQUIRE(ctxs[i]->comm_nranks == ndev || ndev == 1, "the contexts are not one group (cc_ctx_create_group)");
    }
    if (ndev == 1 || count == 0) return CC_OK;
    ncclResult_t r = g_nccl.GroupStart();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
    for (int i = 0; i < ndev && r == ncclSuccess; ++i) {
        r = g_nccl.AllReduce(bufs[i], bufs[i], count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctxs[i]->nccl_comm),
                             static_cast<cudaStream_t>(streams ? streams[i] : nullptr));
        ctxs[i]->collectives++;
    }
    const ncclResult_t e = g_nccl.GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    if (e != ncclSuccess) return nccl_fail(e, "ncclGroupEnd");
    return CC_OK;
}

int cc_comm_nccl_version(int* version) {
    CC_REQUIRE(version != nullptr, "version is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    *version = 0;
    if (g_nccl.GetVersion) g_nccl.GetVersion(version);
    return CC_OK;
}

}  // extern "C"
