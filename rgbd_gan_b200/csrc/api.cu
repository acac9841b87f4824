// api.cu -- error plumbing shared by the C-ABI entry points (include/rgbdgan_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace rgbd {

static thread_local char g_err[512] = "";

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

}
