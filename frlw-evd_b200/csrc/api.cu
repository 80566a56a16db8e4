// Library-level entry points: version, error strings, device query.
#include <string.h>

#include "common.cuh"

namespace evrep {

static thread_local char g_cuda_error[256] = "";

void set_cuda_error(cudaError_t e) {
    strncpy(g_cuda_error, cudaGetErrorString(e), sizeof(g_cuda_error) - 1);
    g_cuda_error[sizeof(g_cuda_error) - 1] = 0;
    (void)cudaGetLastError();   // clear the sticky-less error state for the next call
}

int sm_count() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

}  // namespace evrep

extern "C" {

int evrep_version(void) { return EVREP_VERSION; }

const char* evrep_strerror(int code) {
    switch (code) {
        case EVREP_OK: return "ok";
        case EVREP_ERR_ARG: return "invalid argument";
        case EVREP_ERR_CUDA: return "CUDA error (see evrep_last_cuda_error)";
        case EVREP_ERR_SCRATCH: return "scratch buffer too small";
        case EVREP_ERR_RANGE: return "size out of range for the packed formats";
        default: return "unknown error";
    }
}

const char* evrep_last_cuda_error(void) { return evrep::g_cuda_error; }

int evrep_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    EVREP_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    EVREP_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return EVREP_OK;
}

}  // extern "C"
