// Data-parallel gradient exchange over NVLink 5 / NVSwitch peer memory (mvae_comm_*).
//
// One process per GPU.  Every rank maps every peer's gradient buffer and signal block into its address space (CUDA IPC
// handles travel through torch.distributed once, at set-up) and ONE kernel per step does the whole sum all-reduce:
//
//   barrier   every rank's gradients are final                       (peer stores of a counter, system scope)
//   phase 1   reduce-scatter: rank r sums slice r of all `world` buffers, reading the peers' slices over NVLink, always in
//             rank order 0..world-1 (so the result does not depend on which rank owns the slice), into its own buffer
//   barrier
//   phase 2   all-gather: rank r copies slice p of peer p (p != r) into its own buffer
//   barrier   nobody still reads this rank's buffer (the next step clears it)
//
// The barriers are per CTA: CTA b of rank r only ever touches the elements CTA b of the other ranks touches (same grid, same
// index map relative to the slice), so CTA b waits for the CTAs b of its peers and for nobody else.  No host involvement, no
// parameter that changes from step to step (the barrier counters live in device memory), hence capturable into the step's
// CUDA graph.  A rank that never arrives would make the others spin for ever; the wait gives up after MVAE_COMM_TIMEOUT_MS
// (default 4000) and raises the block's error word, which the host reads with mvae_comm_status.
#include "common.cuh"
#include <string.h>

namespace mvae {
namespace comm {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 128;
constexpr int kThreads = 512;

// one per rank, in its own cudaMalloc allocation (mvae_comm_alloc_signals), zero-initialised
struct Signals {
    unsigned int flag[2][kMaxCtas][kMaxWorld];   // flag[slot][cta][writer rank]
    unsigned int count[kMaxCtas];                // barriers this CTA has passed (only its owner touches it)
    unsigned int error;                          // != 0: a wait timed out
};

struct Peers {
    float* buf[kMaxWorld];
    Signals* sig[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float4 ld16(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// all CTAs `blockIdx.x` of the `world` ranks meet here; `val` is the number of this barrier (1, 2, 3 ... since allocation)
__device__ __forceinline__ void cta_barrier(const Peers& P, int rank, int world, unsigned int val, unsigned long long timeout_ns) {
    __syncthreads();                                    // this CTA's stores precede the release below (cumulativity)
    if (threadIdx.x < world) {
        const int slot = val & 1;                       // a peer can be at most one barrier ahead: two slots never collide
        st_release_sys(&P.sig[threadIdx.x]->flag[slot][blockIdx.x][rank], val);
        const unsigned int* mine = &P.sig[rank]->flag[slot][blockIdx.x][threadIdx.x];
        const unsigned long long t0 = globaltimer();
        while (ld_acquire_sys(mine) != val) {
            if (globaltimer() - t0 > timeout_ns) { P.sig[rank]->error = 1u; break; }
        }
    }
    __syncthreads();
}

template <int WORLD>
__global__ void __launch_bounds__(kThreads) allreduce_kernel(const Peers P, const int rank, const long long n4,
                                                             const unsigned long long timeout_ns) {
    Signals* self = P.sig[rank];
    const unsigned int base = self->count[blockIdx.x];              // written by this CTA only, in the previous launch
    const long long chunk = (n4 + WORLD - 1) / WORLD;               // 16-byte units per slice
    const long long stride = (long long)gridDim.x * kThreads;
    const long long j0 = (long long)blockIdx.x * kThreads + threadIdx.x;

    cta_barrier(P, rank, WORLD, base + 1, timeout_ns);

    {   // reduce-scatter of slice `rank`
        const long long lo = rank * chunk, hi = min(lo + chunk, n4);
        float* mine = P.buf[rank];
        for (long long i = lo + j0; i < hi; i += stride) {
            float4 v[WORLD];
#pragma unroll
            for (int r = 0; r < WORLD; ++r) v[r] = ld16(P.buf[r] + 4 * i);      // all peers' loads in flight together
            float4 a = v[0];
#pragma unroll
            for (int r = 1; r < WORLD; ++r) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
            *reinterpret_cast<float4*>(mine + 4 * i) = a;
        }
    }

    cta_barrier(P, rank, WORLD, base + 2, timeout_ns);

    {   // all-gather: slice p from its owner, the peers visited in a rank-dependent order so that the links share the load
        float* mine = P.buf[rank];
        const long long span = min(chunk, n4);
        for (long long j = j0; j < span; j += stride) {
            float4 v[WORLD - 1];
            bool ok[WORLD - 1];
#pragma unroll
            for (int k = 1; k < WORLD; ++k) {
                const int p = (rank + k) % WORLD;
                const long long i = p * chunk + j;
                ok[k - 1] = i < n4;
                if (ok[k - 1]) v[k - 1] = ld16(P.buf[p] + 4 * i);
            }
#pragma unroll
            for (int k = 1; k < WORLD; ++k) {
                const int p = (rank + k) % WORLD;
                if (ok[k - 1]) *reinterpret_cast<float4*>(mine + 4 * (p * chunk + j)) = v[k - 1];
            }
        }
    }

    cta_barrier(P, rank, WORLD, base + 3, timeout_ns);
    if (threadIdx.x == 0) self->count[blockIdx.x] = base + 3;
}

typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);

static cuMemGetAddressRange_t address_range_fn() {
    static cuMemGetAddressRange_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<cuMemGetAddressRange_t>(p);
    }
    return fn;
}

}  // namespace comm
}  // namespace mvae

using namespace mvae;

extern "C" size_t mvae_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int mvae_comm_alloc_signals(void** signals) {
    MVAE_REQUIRE(signals != nullptr, "mvae_comm_alloc_signals: null output");
    void* p = nullptr;
    MVAE_CUDA(cudaMalloc(&p, sizeof(comm::Signals)));
    MVAE_CUDA(cudaMemset(p, 0, sizeof(comm::Signals)));
    MVAE_CUDA(cudaDeviceSynchronize());
    *signals = p;
    return MVAE_OK;
}

extern "C" int mvae_comm_free_signals(void* signals) {
    if (signals) MVAE_CUDA(cudaFree(signals));
    return MVAE_OK;
}

extern "C" int mvae_comm_export(const void* ptr, void* handle, unsigned long long* offset) {
    MVAE_REQUIRE(ptr && handle && offset, "mvae_comm_export: null argument");
    unsigned long long base = reinterpret_cast<unsigned long long>(ptr);
    size_t size = 0;
    comm::cuMemGetAddressRange_t fn = comm::address_range_fn();
    MVAE_REQUIRE(fn != nullptr, "mvae_comm_export: cuMemGetAddressRange is unavailable");
    if (fn(&base, &size, reinterpret_cast<unsigned long long>(ptr)) != 0) {
        set_error("mvae_comm_export: cuMemGetAddressRange failed for %p", ptr);
        return MVAE_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    MVAE_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    memcpy(handle, &h, sizeof(h));
    *offset = reinterpret_cast<unsigned long long>(ptr) - base;
    return MVAE_OK;
}

extern "C" int mvae_comm_open(const void* handle, unsigned long long offset, void** mapped_base, void** ptr) {
    MVAE_REQUIRE(handle && mapped_base && ptr, "mvae_comm_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    MVAE_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *mapped_base = base;
    *ptr = static_cast<char*>(base) + offset;
    return MVAE_OK;
}

extern "C" int mvae_comm_close(void* mapped_base) {
    if (mapped_base) MVAE_CUDA(cudaIpcCloseMemHandle(mapped_base));
    return MVAE_OK;
}

extern "C" int mvae_comm_allreduce(float* const* bufs, void* const* signals, int rank, int world, long long n, int ctas,
                                   mvae_stream_t stream) {
    MVAE_REQUIRE(bufs && signals, "mvae_comm_allreduce: null argument");
    MVAE_REQUIRE(world >= 2 && world <= comm::kMaxWorld && rank >= 0 && rank < world, "mvae_comm_allreduce: world %d rank %d", world, rank);
    MVAE_REQUIRE(n > 0 && (n % 4) == 0, "mvae_comm_allreduce: n = %lld must be a positive multiple of 4 floats", n);
    comm::Peers P = {};
    for (int r = 0; r < world; ++r) {
        MVAE_REQUIRE(bufs[r] && signals[r] && (reinterpret_cast<uintptr_t>(bufs[r]) & 15) == 0,
                     "mvae_comm_allreduce: buffer of rank %d is null or not 16-byte aligned", r);
        P.buf[r] = bufs[r];
        P.sig[r] = static_cast<comm::Signals*>(signals[r]);
    }
    if (ctas <= 0) ctas = env_int("MVAE_COMM_CTAS", 64);
    if (ctas > comm::kMaxCtas) ctas = comm::kMaxCtas;
    const unsigned long long timeout_ns = 1000000ull * (unsigned long long)env_int("MVAE_COMM_TIMEOUT_MS", 4000);
    const long long n4 = n / 4;
    cudaStream_t s = as_stream(stream);
    switch (world) {
#define MVAE_COMM_CASE(W) \
        case W: MVAE_CUDA(launch_pdl_ex(false, comm::allreduce_kernel<W>, dim3(ctas), dim3(comm::kThreads), 0, s, P, rank, n4, timeout_ns)); break;
        MVAE_COMM_CASE(2) MVAE_COMM_CASE(3) MVAE_COMM_CASE(4) MVAE_COMM_CASE(5) MVAE_COMM_CASE(6) MVAE_COMM_CASE(7) MVAE_COMM_CASE(8)
#undef MVAE_COMM_CASE
    }
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_comm_status(const void* signals, int* timed_out) {
    MVAE_REQUIRE(signals && timed_out, "mvae_comm_status: null argument");
    unsigned int e = 0;
    MVAE_CUDA(cudaMemcpy(&e, &static_cast<const comm::Signals*>(signals)->error, sizeof(e), cudaMemcpyDeviceToHost));
    *timed_out = (int)e;
    return MVAE_OK;
}
