// api.cu -- error plumbing shared by the C-ABI entry points (include/rgbdgan_b200.h).
#include <stdarg.h>

#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace rgbd {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
thread_local cudaEvent_t g_hook_start = nullptr, g_hook_stop = nullptr;

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch-time errors only (bad configuration, missing kernel image for the device, ...).
// Never synchronises: asynchronous faults surface at the caller's next sync.
int check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

}  // namespace rgbd

extern "C" {

RGBD_API int rgbd_version(void) { return RGBD_B200_VERSION; }

RGBD_API const char *rgbd_last_error(void) { return rgbd::g_err; }

RGBD_API unsigned long long rgbd_launch_count(void) { return rgbd::g_launches.load(); }

RGBD_API int rgbd_peer_comm_create(int rank, int world, void **comm_out, unsigned char *ipc_handle_out)
{
    using namespace rgbd;
    if (!comm_out || !ipc_handle_out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) {
        set_error("rgbd_peer_comm_create: bad arguments (rank %d world %d, max %d)", rank, world, kMaxPeers);
        return RGBD_E_ARG;
    }
    rgbd_peer_comm *pc = new rgbd_peer_comm();
    memset(pc, 0, sizeof(*pc));
    cudaError_t e = cudaMalloc((void **)&pc->mine, sizeof(rgbd_mailbox));
    if (e == cudaSuccess) e = cudaMemset(pc->mine, 0, sizeof(rgbd_mailbox));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, pc->mine);
    if (e != cudaSuccess) {
        set_error("rgbd_peer_comm_create: %s", cudaGetErrorString(e));
        if (pc->mine) cudaFree(pc->mine);
        delete pc;
        return (int)e;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &h, 64);
    pc->args.rank = rank; pc->args.world = world; pc->args.box[rank] = pc->mine;
    {   // RGBD_B200_PEER_TIMEOUT_MS (default 2000): bound of every wait for a peer's flag inside the finalize kernel
        const char *e = getenv("RGBD_B200_PEER_TIMEOUT_MS");
        const long ms = e && *e ? atol(e) : 2000;
        pc->args.timeout_ns = (unsigned long long)(ms > 0 ? ms : 2000) * 1000000ull;
    }
    e = cudaStreamCreateWithFlags(&pc->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pc->ev_main_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pc->ev_fin_done, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("rgbd_peer_comm_create: %s", cudaGetErrorString(e)); return (int)e; }
    *comm_out = pc;
    return 0;
}

RGBD_API int rgbd_peer_comm_connect(void *comm, const unsigned char *all_handles)
{
    using namespace rgbd;
    rgbd_peer_comm *pc = (rgbd_peer_comm *)comm;
    if (!pc || !all_handles) { set_error("rgbd_peer_comm_connect: null argument"); return RGBD_E_ARG; }
    for (int r = 0; r < pc->args.world; ++r) {
        if (r == pc->args.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * r, 64);
        void *ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("rgbd_peer_comm_connect: rank %d: %s", r, cudaGetErrorString(e));
            return (int)e;
        }
        pc->args.box[r] = (rgbd_mailbox *)ptr;
    }
    pc->connected = true;
    return 0;
}

RGBD_API int rgbd_peer_comm_destroy(void *comm)
{
    using namespace rgbd;
    rgbd_peer_comm *pc = (rgbd_peer_comm *)comm;
    if (!pc) return 0;
    for (int r = 0; r < pc->args.world; ++r) {
        if (pc->loopback[r]) cudaFree(pc->loopback[r]);
        else if (r != pc->args.rank && pc->args.box[r]) cudaIpcCloseMemHandle(pc->args.box[r]);
    }
    if (pc->side) { cudaStreamSynchronize(pc->side); cudaStreamDestroy(pc->side); }
    if (pc->ev_main_done) cudaEventDestroy(pc->ev_main_done);
    if (pc->ev_fin_done) cudaEventDestroy(pc->ev_fin_done);
    if (pc->mine) cudaFree(pc->mine);
    delete pc;
    return 0;
}

RGBD_API int rgbd_peer_comm_wait(void *comm, void *stream)
{
    using namespace rgbd;
    rgbd_peer_comm *pc = (rgbd_peer_comm *)comm;
    if (!pc) { set_error("rgbd_peer_comm_wait: null comm"); return RGBD_E_ARG; }
    if (pc->lazy_pending) return launch_peer_collect(pc, (cudaStream_t)stream);
    if (!pc->fin_pending) return 0;
    const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, pc->ev_fin_done, 0);
    if (e != cudaSuccess) { set_error("rgbd_peer_comm_wait: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

RGBD_API int rgbd_peer_comm_status(void *comm, void *stream, int *status_host)
{
    using namespace rgbd;
    rgbd_peer_comm *pc = (rgbd_peer_comm *)comm;
    if (!pc || !status_host) { set_error("rgbd_peer_comm_status: null argument"); return RGBD_E_ARG; }
    if (pc->side) cudaStreamSynchronize(pc->side);
    unsigned err = 0;
    cudaError_t e = cudaMemcpyAsync(&err, &pc->mine->error, sizeof(err), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("rgbd_peer_comm_status: %s", cudaGetErrorString(e)); return (int)e; }
    *status_host = err ? 1 : 0;
    return 0;
}

RGBD_API int rgbd_debug_peer_comm_loopback(void *comm)
{
    using namespace rgbd;
    rgbd_peer_comm *pc = (rgbd_peer_comm *)comm;
    if (!pc || pc->connected) { set_error("rgbd_debug_peer_comm_loopback: null or already connected comm"); return RGBD_E_ARG; }
    for (int r = 0; r < pc->args.world; ++r) {
        if (r == pc->args.rank) continue;
        cudaError_t e = cudaMalloc((void **)&pc->loopback[r], sizeof(rgbd_mailbox));
        if (e == cudaSuccess) e = cudaMemset(pc->loopback[r], 0, sizeof(rgbd_mailbox));
        if (e != cudaSuccess) { set_error("rgbd_debug_peer_comm_loopback: %s", cudaGetErrorString(e)); return (int)e; }
        pc->args.box[r] = pc->loopback[r];
    }
    pc->connected = true;
    return 0;
}

RGBD_API void rgbd_profile_hook(void *ev_start, void *ev_stop)
{
    rgbd::g_hook_start = (cudaEvent_t)ev_start;
    rgbd::g_hook_stop = (cudaEvent_t)ev_stop;
}

}
