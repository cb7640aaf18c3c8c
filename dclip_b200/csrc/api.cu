// Library-wide C-ABI entry points: version, error string, device capability.
#include "common.cuh"

namespace dclip {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

bool device_is_sm100() {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    return major == 10;
}

}  // namespace dclip

extern "C" {

int dclip_version(void) { return DCLIP_B200_VERSION; }

int dclip_last_error(char* buf, size_t buf_len) {
    if (!buf || buf_len == 0) return DCLIP_ERR_INVALID;
    strncpy(buf, dclip::last_error_buffer(), buf_len - 1);
    buf[buf_len - 1] = 0;
    return DCLIP_OK;
}

int dclip_device_supports_tensor_path(int device) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return major == 10 ? 1 : 0;
}

}  // extern "C"
