// api.cu -- error plumbing shared by the C-ABI entry points (include/rgbdgan_b200.h).
#include <stdarg.h>

#include <atomic>
#include <stdio.h>

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

RGBD_API void rgbd_profile_hook(void *ev_start, void *ev_stop)
{
    rgbd::g_hook_start = (cudaEvent_t)ev_start;
    rgbd::g_hook_stop = (cudaEvent_t)ev_stop;
}

}
