// Library plumbing for the C-ABI: version, thread-local error text, per-device init.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace fcwdm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int conv3d_init_device();       // conv3d.cu
int conv3d_pair_init_device();  // conv3d_pair.cu
int conv3d_wgrad_init_device(); // conv3d_wgrad.cu
int conv3d_chain_init_device(); // conv3d_chain.cu

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FCWDM_NO_PDL");
        v = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace fcwdm

extern "C" int fcwdm_version(void) { return FCWDM_VERSION; }

extern "C" const char* fcwdm_last_error(void) { return fcwdm::g_err; }

extern "C" int fcwdm_init(int device) {
    cudaError_t e = cudaSetDevice(device);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_init: cudaSetDevice(%d) failed: %s", device,
                  cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    FCWDM_REQUIRE(major == 10, FCWDM_ERR_ARCH,
                  "fcwdm_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, major,
                  minor);
    int rc = fcwdm::conv3d_init_device();
    if (rc) return rc;
    rc = fcwdm::conv3d_pair_init_device();
    if (rc) return rc;
    rc = fcwdm::conv3d_wgrad_init_device();
    if (rc) return rc;
    return fcwdm::conv3d_chain_init_device();
}
