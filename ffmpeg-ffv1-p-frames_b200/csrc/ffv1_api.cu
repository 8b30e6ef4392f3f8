// ffv1_api.cu -- library-level entry points: version, error text, device probing.
#include "../../include/ffv1_b200.h"
#include "ffv1_internal.h"
#include <cuda_runtime.h>
#include <string>

namespace ffv1 {
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
__global__ void k_probe(int *out) { if (threadIdx.x == 0) *out = 100; }
}

extern "C" {

const char *ffv1b200_version(void) { return "ffv1_b200 0.1 (sm_100a)"; }

const char *ffv1b200_last_error(void) { return ffv1::g_last_error.c_str(); }

const char *ffv1b200_strerror(int err)
{
    switch (err) {
    case 0: return "success";
    case FFV1B200_ERR_EINVAL: return "Invalid argument";
    case FFV1B200_ERR_ENOMEM: return "Cannot allocate memory";
    case FFV1B200_ERR_ENOSYS: return "Function not implemented";
    case FFV1B200_ERR_INVALIDDATA: return "Invalid data found when processing input";
    case FFV1B200_ERR_EXTERNAL: return "Generic error in an external library (CUDA)";
    case FFV1B200_ERR_BUFFER_TOO_SMALL: return "Buffer too small";
    default: return "unknown error";
    }
}

int ffv1b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        ffv1::set_last_error(std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                             " (this library has no CPU fallback)");
        cudaGetLastError();
        return FFV1B200_ERR_EXTERNAL;
    }
    // the kernel image is sm_100a only: make sure it can actually be loaded on device 0
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, ffv1::k_probe);
    if (e != cudaSuccess) {
        ffv1::set_last_error(std::string("sm_100a kernel image not loadable on this device: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return FFV1B200_ERR_EXTERNAL;
    }
    return n;
}

} // extern "C"
