// Library-level entry points: version, error text, device check, memset.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace mvae {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MVAE_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
long long g_kernel_launches = 0;
bool pdl_chain_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MVAE_PDL_CHAIN"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
}  // namespace mvae

using namespace mvae;

namespace mvae {
__global__ void accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, int n, float alpha) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = fmaf(alpha, src[i], dst[i]);
}
}  // namespace mvae

extern "C" int mvae_version(void) { return 100; }   // 0.1.0

extern "C" int mvae_last_error(char* buf, size_t n) {
    if (!buf || n == 0) return MVAE_ERR_ARG;
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
    return MVAE_OK;
}

extern "C" int mvae_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    MVAE_CUDA(cudaGetDevice(&dev));
    MVAE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MVAE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}

extern "C" int mvae_memset_zero(void* ptr, size_t bytes, mvae_stream_t stream) {
    MVAE_REQUIRE(ptr || bytes == 0, "memset_zero: null pointer");
    if (bytes) MVAE_CUDA(cudaMemsetAsync(ptr, 0, bytes, as_stream(stream)));
    return MVAE_OK;
}

extern "C" int mvae_accumulate(float* dst, const float* src, int n, float alpha, mvae_stream_t stream) {
    MVAE_REQUIRE(dst && src && n >= 0, "accumulate: bad arguments");
    if (n == 0) return MVAE_OK;
    MVAE_CUDA(launch_pdl_ex(false, accumulate_kernel, dim3((n + 255) / 256), dim3(256), 0, as_stream(stream), dst, src, n, alpha));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_stream_create(int priority, mvae_stream_t* stream) {
    MVAE_REQUIRE(stream != nullptr, "stream_create: null output");
    int lo = 0, hi = 0;
    MVAE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // numerically lower = higher priority: lo = least, hi = greatest
    const int p = priority <= 0 ? lo : priority >= 2 ? hi : (lo + hi) / 2;
    cudaStream_t s = nullptr;
    MVAE_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, p));
    *stream = reinterpret_cast<mvae_stream_t>(s);
    return MVAE_OK;
}

extern "C" int mvae_stream_destroy(mvae_stream_t stream) {
    if (stream) MVAE_CUDA(cudaStreamDestroy(as_stream(stream)));
    return MVAE_OK;
}

extern "C" long long mvae_kernel_launch_count(void) { return g_kernel_launches; }
