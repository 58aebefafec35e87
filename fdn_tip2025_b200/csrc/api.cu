// libfdn_b200: error plumbing and library-level entry points of the C ABI (see include/fdn_b200.h).
#include "fdn_common.cuh"

static thread_local std::string g_fdn_error;

void fdn_set_error(const std::string& msg) { g_fdn_error = msg; }

int fdn_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_fdn_error = std::string(what) + ": " + cudaGetErrorString(e);
        return (int)e;
    }
    return 0;
}

FDN_API const char* fdn_last_error_string() { return g_fdn_error.c_str(); }

FDN_API int fdn_abi_version() { return 1; }

// 1 when the library was built for the GPU (sm_100a), 0 for the host emulation build used by tests/emu only.
FDN_API int fdn_is_device_build() {
#ifdef FDN_EMU
    return 0;
#else
    return 1;
#endif
}
