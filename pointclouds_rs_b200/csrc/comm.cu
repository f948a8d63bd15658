// comm.cu -- NCCL plumbing for the one exchange the path has: the per-iteration all-reduce of the
// ICP normal equations (30 doubles) when the source cloud is sharded over GPUs.
//
// NCCL is resolved at run time with dlopen so that (a) the library has no link-time dependency on
// it (single-GPU users never load it) and (b) inside a process that already carries a NCCL (e.g.
// one that imported torch) the SAME copy is reused instead of a second one.
#include "pcr_internal.cuh"

#include <dlfcn.h>

#include <vector>

namespace pcr {

namespace {

typedef struct {
    char internal[128];
} ncclUniqueId_t;
typedef void *ncclComm_t_;
typedef int ncclResult_t_;
enum { kNcclUint8 = 1, kNcclUint32 = 3, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t_ (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    ncclResult_t_ (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_t, int) = nullptr;
    ncclResult_t_ (*CommDestroy)(ncclComm_t_) = nullptr;
    ncclResult_t_ (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
    ncclResult_t_ (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t_) = nullptr;
    bool ok = false;
};

NcclApi load_nccl() {
    NcclApi a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {  // a copy already mapped into the process wins
        a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) {
        const char *env = getenv("PCR_NCCL_LIBRARY");
        if (env) a.handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    for (const char *n : names) {
        if (a.handle) break;
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!a.handle) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather;
    return a;
}

// resolved once, by whichever thread asks first (a C++11 magic static: concurrent callers wait for the loader)
NcclApi &api() {
    static NcclApi a = load_nccl();
    return a;
}

}  // namespace

int comm_unique_id(void *out) {
    static_assert(sizeof(ncclUniqueId_t) == PCR_UNIQUE_ID_BYTES, "id size");
    NcclApi &a = api();
    if (!a.ok) {
        set_thread_error("NCCL library not found (set PCR_NCCL_LIBRARY)");
        return PCR_ERR_NCCL;
    }
    ncclUniqueId_t id;
    int r = a.GetUniqueId(&id);
    if (r != 0) {
        set_thread_error(a.GetErrorString ? a.GetErrorString(r) : "ncclGetUniqueId failed");
        return PCR_ERR_NCCL;
    }
    memcpy(out, &id, sizeof(id));
    return PCR_OK;
}

int comm_allgather_bytes(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes);
int comm_allreduce_u32(Ctx *ctx, uint32_t *d_buf, size_t count);

// ---- exchange blocks for the one-shot all-reduce of the ICP sums (icp.cu: icp_reduce_kernel) ----------------------------
// Every rank allocates one block with cudaMalloc (pool memory cannot be exported), exports it with cudaIpcGetMemHandle, the
// handles travel through ncclAllGather, every rank opens its peers' blocks (one process per GPU: CUDA IPC; NVLink peer access is
// enabled lazily by the open).  All ranks must take the same path, so the outcome is agreed on with an integer all-reduce: if
// any rank failed anywhere, all fall back to ncclAllReduce.  PCR_ICP_NCCL=1 forces that (A/B hook).
static void peer_teardown(Ctx *ctx) {
    for (int p = 0; p < kPeerMaxWorld; p++) {
        if (ctx->peer_open[p]) cudaIpcCloseMemHandle(ctx->peer_open[p]);
        ctx->peer_open[p] = nullptr;
    }
    if (ctx->peer_block) cudaFree(ctx->peer_block);
    ctx->peer_block = nullptr;
    ctx->peer_ok = false;
    ctx->peer = {};
    cudaGetLastError();
}

static void peer_setup(Ctx *ctx) {
    const int W = ctx->world, R = ctx->rank;
    if (W <= 1 || W > kPeerMaxWorld) return;
    bool ok = getenv("PCR_ICP_NCCL") == nullptr;
    unsigned char *d_tmp = nullptr;  // [0, 64 W): the handles; then one u32 for the agreement
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    std::vector<cudaIpcMemHandle_t> handles(W);
    if (cudaMalloc((void **)&d_tmp, hb * W + 64) != cudaSuccess) {
        cudaGetLastError();
        return;  // (no collective has been issued yet: returning here cannot desynchronise the ranks only if ALL fail; see below)
    }
    if (ok && cudaMalloc(&ctx->peer_block, kPeerBlockBytes) != cudaSuccess) ok = false;
    if (ok && cudaMemset(ctx->peer_block, 0, kPeerBlockBytes) != cudaSuccess) ok = false;
    cudaIpcMemHandle_t mine = {};
    if (ok && cudaIpcGetMemHandle(&mine, ctx->peer_block) != cudaSuccess) ok = false;
    cudaGetLastError();
    // the collectives below are issued by every rank whatever happened above
    cudaMemcpy(d_tmp + hb * R, &mine, hb, cudaMemcpyHostToDevice);
    bool coll = comm_allgather_bytes(ctx, d_tmp + hb * R, d_tmp, hb) == PCR_OK;
    coll = coll && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (coll) cudaMemcpy(handles.data(), d_tmp, hb * W, cudaMemcpyDeviceToHost);
    ok = ok && coll;
    for (int p = 0; ok && p < W; p++) {
        if (p == R) continue;
        if (cudaIpcOpenMemHandle(&ctx->peer_open[p], handles[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
    }
    cudaGetLastError();
    uint32_t good = ok ? 1u : 0u, *d_good = (uint32_t *)(d_tmp + hb * W);
    cudaMemcpy(d_good, &good, sizeof(good), cudaMemcpyHostToDevice);
    if (comm_allreduce_u32(ctx, d_good, 1) == PCR_OK && cudaStreamSynchronize(ctx->stream) == cudaSuccess)
        cudaMemcpy(&good, d_good, sizeof(good), cudaMemcpyDeviceToHost);
    else
        good = 0;
    cudaFree(d_tmp);
    cudaGetLastError();
    if (good != (uint32_t)W) {
        peer_teardown(ctx);
        return;
    }
    ctx->peer.rank = R;
    ctx->peer.world = W;
    for (int p = 0; p < W; p++) {
        char *base = (char *)(p == R ? ctx->peer_block : ctx->peer_open[p]);
        ctx->peer.data[p] = (double *)base;
        ctx->peer.flag[p] = (unsigned long long *)(base + kPeerDataBytes);
    }
    ctx->peer_seq = 0;
    ctx->peer_ok = true;
}

int comm_init(Ctx *ctx, const void *id, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, PCR_ERR_INVALID_ARG, "bad rank/world");
    comm_destroy(ctx);
    ctx->rank = rank;
    ctx->world = world;
    if (world == 1) return PCR_OK;
    NcclApi &a = api();
    if (!a.ok) return fail(ctx, PCR_ERR_NCCL, "NCCL library not found (set PCR_NCCL_LIBRARY)");
    ncclUniqueId_t uid;
    memcpy(&uid, id, sizeof(uid));
    PCR_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclComm_t_ comm = nullptr;
    int r = a.CommInitRank(&comm, world, uid, rank);
    if (r != 0) {
        ctx->world = 1;
        ctx->rank = 0;
        return fail(ctx, PCR_ERR_NCCL, "ncclCommInitRank: %s", a.GetErrorString ? a.GetErrorString(r) : "error");
    }
    ctx->nccl_comm = comm;
    peer_setup(ctx);  // (best effort: without it the ICP sums go through ncclAllReduce)
    return PCR_OK;
}

void comm_destroy(Ctx *ctx) {
    ctx->fake_comm = false;
    peer_teardown(ctx);
    if (ctx->nccl_comm) {
        NcclApi &a = api();
        if (a.ok) a.CommDestroy((ncclComm_t_)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->world = 1;
    ctx->rank = 0;
}

int comm_allreduce_f64(Ctx *ctx, double *d_buf, size_t count) {
    if (ctx->world <= 1) return PCR_OK;
    NcclApi &a = api();
    if (!a.ok || !ctx->nccl_comm) return fail(ctx, PCR_ERR_NCCL, "communicator not initialised");
    int r = a.AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, PCR_ERR_NCCL, "ncclAllReduce: %s", a.GetErrorString ? a.GetErrorString(r) : "error");
    return PCR_OK;
}

int comm_allreduce_u32(Ctx *ctx, uint32_t *d_buf, size_t count) {
    if (ctx->world <= 1 || count == 0 || ctx->fake_comm) return PCR_OK;
    NcclApi &a = api();
    if (!a.ok || !ctx->nccl_comm) return fail(ctx, PCR_ERR_NCCL, "communicator not initialised");
    int r = a.AllReduce(d_buf, d_buf, count, kNcclUint32, kNcclSum, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, PCR_ERR_NCCL, "ncclAllReduce: %s", a.GetErrorString ? a.GetErrorString(r) : "error");
    return PCR_OK;
}

// every rank contributes `bytes` bytes at d_send; d_recv receives world * bytes, rank-major
int comm_allgather_bytes(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes) {
    if (ctx->world <= 1) {
        if (d_send != d_recv && bytes) PCR_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return PCR_OK;
    }
    NcclApi &a = api();
    if (!a.ok || !ctx->nccl_comm) return fail(ctx, PCR_ERR_NCCL, "communicator not initialised");
    int r = a.AllGather(d_send, d_recv, bytes, kNcclUint8, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, PCR_ERR_NCCL, "ncclAllGather: %s", a.GetErrorString ? a.GetErrorString(r) : "error");
    return PCR_OK;
}

}  // namespace pcr
