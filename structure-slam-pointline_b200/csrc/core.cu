// core.cu — error reporting, version, pinned host memory of libsslpl_b200.
#include "common.cuh"
#include <cstdlib>
#include <cstdarg>

namespace sslpl {
static thread_local char t_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return t_err; }
}  // namespace sslpl

extern "C" {
const char* sslpl_last_error(void) { return sslpl::get_error(); }
int sslpl_version(void) { return SSLPL_VERSION; }
int sslpl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int sslpl_default_device(void) {
    const char* e = getenv("SSLPL_DEVICE");
    return e && *e ? atoi(e) : 0;
}
int sslpl_host_alloc(void** p, size_t bytes) {
    SSLPL_REQUIRE(p, SSLPL_ERR_ARG, "null argument");
    SSLPL_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return SSLPL_OK;
}
int sslpl_host_free(void* p) {
    if (p) SSLPL_CUDA(cudaFreeHost(p));
    return SSLPL_OK;
}
}
