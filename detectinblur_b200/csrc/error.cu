// Thread-local error message + small device-info entry points of the C ABI (include/dib.h).
#include <stdarg.h>

#include "dib_common.cuh"

namespace dib {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dib

extern "C" int dib_abi_version(void) { return DIB_ABI_VERSION; }

extern "C" const char* dib_last_error(void) { return dib::g_err; }

extern "C" int dib_device_info(int* sm_count, int* cc) {
    int dev = 0;
    DIB_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    DIB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DIB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DIB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc) *cc = major * 10 + minor;
    return DIB_OK;
}

extern "C" int dib_tapset_layout_for(int n_psfs, int max_taps, dib_tapset_layout* out) {
    DIB_CHECK_ARG(n_psfs > 0 && max_taps > 0 && out != nullptr, "dib_tapset_layout_for: n_psfs, max_taps must be > 0");
    *out = dib::tapset_layout(n_psfs, max_taps);
    return DIB_OK;
}
