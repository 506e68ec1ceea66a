// Error reporting + device check for libvitae_b200.so.
#include "common.h"
#include <stdlib.h>

namespace vitae {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("VITAE_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
}  // namespace vitae

extern "C" int vitae_abi_version(void) { return VITAE_ABI_VERSION; }
extern "C" const char* vitae_last_error(void) { return vitae::g_err; }
extern "C" long long vitae_launch_count(void) { return vitae::g_launch_count.load(std::memory_order_relaxed); }
extern "C" int vitae_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return vitae::set_error(-1, "cudaGetDevice: %s", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) return vitae::set_error(-1, "libvitae_b200 needs an sm_100a device, found sm_%d%d", major, minor);
    return 0;
}
