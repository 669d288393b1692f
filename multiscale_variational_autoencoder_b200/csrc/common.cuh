// Shared helpers for the mvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/mvae_b200.h"

namespace mvae {

// tuning knob read from the environment (diagnostics / experiments; every default is the measured best)
static inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

void set_error(const char* fmt, ...);

#define MVAE_REQUIRE(cond, ...)                    \
    do {                                           \
        if (!(cond)) {                             \
            mvae::set_error(__VA_ARGS__);          \
            return MVAE_ERR_ARG;                   \
        }                                          \
    } while (0)

#define MVAE_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            mvae::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return MVAE_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define MVAE_LAUNCH_CHECK() MVAE_CUDA(cudaGetLastError())

// Function attributes (dynamic shared-memory limit, carve-out) are per DEVICE: a call site keeps one flag per device, so a
// second model on another GPU of the same process configures its kernels too.
struct DeviceOnce {
    bool seen[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (seen[d]) return false;
        seen[d] = true;
        return true;
    }
};

static inline cudaStream_t as_stream(mvae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;   // B200

// Programmatic dependent launch: the next kernel of the stream is scheduled (and runs its prologue: barrier init, TMEM
// allocation, descriptor prefetch) while this one drains; it blocks in pdl_wait() until every prerequisite grid has
// completed and flushed.  Kernels call pdl_wait() before their first global access and pdl_launch() right after, so at most
// one successor overlaps.  Captured into CUDA graphs as programmatic dependency edges.  Opt-in (MVAE_PDL=1): measured on
// cfg2 it does not pay (4.27 vs 4.15 ms/step) because the graph is bound by kernel durations, not by launch gaps.
extern long long g_kernel_launches;      // every kernel launch of the library goes through launch_pdl_ex and counts here
bool pdl_enabled();
// MVAE_PDL_CHAIN=1: programmatic dependent launch for the kernels of a mobilenetV3 chain only (the fused tile kernels and the
// squeeze-excite gate kernels between them): each stages its weights before griddepcontrol.wait, i.e. under its predecessor
bool pdl_chain_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_ex(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        Args&&... args);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
    return launch_pdl_ex(pdl_enabled(), kernel, grid, block, smem, s, static_cast<Args&&>(args)...);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_ex(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        Args&&... args) {
    // Every kernel of the library asks for the maximum shared-memory carve-out: the tcgen05 / TMA kernels need it, and a
    // kernel that lets the driver pick a small carve-out forces an SM reconfiguration (a drain of several microseconds)
    // between itself and its neighbours in the stream.
    {
        static const void* seen[512];
        static int seen_dev[512];
        static int nseen = 0;
        const void* key = reinterpret_cast<const void*>(kernel);
        int dev = 0;
        cudaGetDevice(&dev);
        bool found = false;
        for (int i = 0; i < nseen; ++i) found |= (seen[i] == key && seen_dev[i] == dev);
        if (!found) {
            cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            if (nseen < 512) { seen[nseen] = key; seen_dev[nseen] = dev; ++nseen; }
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    ++g_kernel_launches;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_launch(); }
#endif

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// TensorFlow 'SAME' padding: out = ceil(in/s), pad_before = max((out-1)*s + k - in, 0) / 2
static inline void same_pad(int in, int k, int s, int* out, int* pad_before) {
    int o = (in + s - 1) / s;
    int total = (o - 1) * s + k - in;
    if (total < 0) total = 0;
    *out = o;
    *pad_before = total / 2;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == MVAE_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == MVAE_ACT_ELU) return v > 0.f ? v : expm1f(v);
    return v;
}

// derivative of the activation expressed through its OUTPUT o: relu' = (o>0); elu' = o>0 ? 1 : o+1
__device__ __forceinline__ float act_grad_from_out(float o, int act) {
    if (act == MVAE_ACT_RELU) return o > 0.f ? 1.f : 0.f;
    if (act == MVAE_ACT_ELU) return o > 0.f ? 1.f : o + 1.f;
    return 1.f;
}

}  // namespace mvae
