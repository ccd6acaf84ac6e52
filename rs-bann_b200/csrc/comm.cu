// Host side of the peer-memory exchange (comm.cuh): inbox allocation, handle export, peer mapping.
// One process per GPU exchanges the handles through the host layer (torch.distributed all-gather of
// bytes, rs-bann_b200/dist.py); contexts of the SAME process (several ranks emulated on one GPU in
// the tests) are connected through the raw pointer carried in the handle.
#include <unistd.h>

#include "store.cuh"

using namespace bann;

namespace {
struct HandleBlob {               // BANN_COMM_HANDLE_BYTES
    cudaIpcMemHandle_t ipc;       // 64 bytes
    uint64_t pid;
    uint64_t raw;                 // device pointer in the exporting process
    int32_t device;
    int32_t has_ipc;
    uint8_t pad[BANN_COMM_HANDLE_BYTES - 64 - 24];
};
static_assert(sizeof(HandleBlob) == BANN_COMM_HANDLE_BYTES, "handle blob size");
size_t inbox_bytes(int world) { return (size_t)2 * world * kXrCap * sizeof(uint2); }
}  // namespace

namespace bann {
// one device allocation -> BANN_COMM_HANDLE_BYTES blob / blob of a peer -> a pointer valid in this process
int comm_export(void* dptr, int device, uint8_t* out) {
    HandleBlob h;
    memset(&h, 0, sizeof(h));
    h.pid = (uint64_t)getpid();
    h.raw = (uint64_t)(uintptr_t)dptr;
    h.device = device;
    h.has_ipc = cudaIpcGetMemHandle(&h.ipc, dptr) == cudaSuccess ? 1 : 0;
    if (!h.has_ipc) cudaGetLastError();
    memcpy(out, &h, sizeof(h));
    return 0;
}
int comm_import(bann_ctx* ctx, const uint8_t* blob, void** out, bool* opened_ipc) {
    HandleBlob h;
    memcpy(&h, blob, sizeof(h));
    if (h.pid == (uint64_t)getpid()) {             // same process: the pointer is valid as it is
        if (h.device != ctx->device) {
            int can = 0;
            BANN_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
            if (!can) BANN_FAIL("no peer access between the devices of two ranks");
            cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) BANN_CUDA(e);
            cudaGetLastError();
        }
        *out = reinterpret_cast<void*>((uintptr_t)h.raw);
        *opened_ipc = false;
    } else {
        if (!h.has_ipc) BANN_FAIL("a peer rank could not export a CUDA IPC handle");
        void* p = nullptr;
        BANN_CUDA(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
        *out = p;
        *opened_ipc = true;
    }
    return 0;
}
XrComm xr_next(bann_ctx* ctx, int* error_flag) {
    XrComm c;
    memset(&c, 0, sizeof(c));
    c.rank = (uint32_t)ctx->rank;
    c.world = ctx->xr_connected ? (uint32_t)ctx->world : 1u;
    c.error_flag = error_flag;
    if (ctx->xr_connected) {
        for (int r = 0; r < ctx->world; ++r) c.inbox[r] = ctx->xr_inbox[r];
        c.epoch = ++ctx->xr_epoch;
    }
    return c;
}
void xr_release(bann_ctx* ctx) {
    for (int r = 0; r < kXrMaxWorld; ++r) {
        if (!ctx->xr_inbox[r]) continue;
        if (r == ctx->rank) cudaFree(ctx->xr_inbox[r]);
        else if (ctx->xr_ipc[r]) cudaIpcCloseMemHandle(ctx->xr_inbox[r]);
        ctx->xr_inbox[r] = nullptr;
    }
    ctx->xr_connected = false;
}
}  // namespace bann

extern "C" {

int bann_ctx_comm_handle(bann_ctx* ctx, uint8_t* out) {
    if (!ctx || !out) BANN_FAIL("NULL argument");
    if (ctx->world > kXrMaxWorld) BANN_FAIL("peer-memory exchange supports at most 8 ranks");
    BANN_CUDA(cudaSetDevice(ctx->device));
    uint2*& mine = ctx->xr_inbox[ctx->rank];
    if (!mine) {
        BANN_CUDA(cudaMalloc(&mine, inbox_bytes(ctx->world)));
        BANN_CUDA(cudaMemset(mine, 0, inbox_bytes(ctx->world)));   // epoch 0 = "nothing yet"
        BANN_CUDA(cudaDeviceSynchronize());
    }
    return comm_export(mine, ctx->device, out);
}

int bann_ctx_comm_connect(bann_ctx* ctx, const uint8_t* handles) {
    if (!ctx || !handles) BANN_FAIL("NULL argument");
    if (!ctx->xr_inbox[ctx->rank]) BANN_FAIL("bann_ctx_comm_connect before bann_ctx_comm_handle");
    BANN_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->world; ++r) {
        if (r == ctx->rank) continue;
        void* p = nullptr;
        bool ipc = false;
        BANN_CHECK(comm_import(ctx, handles + (size_t)r * BANN_COMM_HANDLE_BYTES, &p, &ipc));
        ctx->xr_inbox[r] = reinterpret_cast<uint2*>(p);
        ctx->xr_ipc[r] = ipc;
    }
    ctx->xr_connected = true;
    ctx->xr_epoch = 0;
    return 0;
}

int bann_ctx_comm_connected(bann_ctx* ctx) { return ctx && ctx->xr_connected ? 1 : 0; }

}  // extern "C"
