// Data-parallel gradient exchange over NVLink 5 / NVSwitch peer memory (mvae_comm_*).
//
// One process per GPU.  Every rank maps every peer's gradient buffer and signal block into its address space (CUDA IPC
// handles travel through torch.distributed once, at set-up) and ONE kernel per step does the whole sum all-reduce:
//
//   barrier   every rank's gradients are final                       (peer stores of a counter, system scope)
//   exchange  rank r sums slice r of all `world` buffers, reading the peers' slices over NVLink, always in rank order
//             0..world-1 (so the result does not depend on which rank owns the slice), and stores the sum into slice r of
//             EVERY rank's buffer (posted stores over NVLink: no second round trip, no third barrier)
//   barrier   every slice has landed everywhere, and nobody still reads this rank's buffer (the next step clears it)
//
// The barriers are per CTA: CTA b of rank r only ever touches the elements CTA b of the other ranks touches (same grid, same
// index map relative to the slice), so CTA b waits for the CTAs b of its peers and for nobody else.  No host involvement, no
// parameter that changes from step to step (the barrier counters live in device memory), hence capturable into the step's
// CUDA graph.  A rank that never arrives would make the others spin for ever; the wait gives up after MVAE_COMM_TIMEOUT_MS
// (default 20000) and raises the block's error word, which the host reads with mvae_comm_status.
#include "common.cuh"
#include <string.h>

namespace mvae {
namespace comm {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 256;
constexpr int kThreads = 256;
constexpr int kChannels = 16;      // exchanges on different channels may run concurrently (each has its own counters)
constexpr int kMaxRanges = 8;

// one per rank, in its own cudaMalloc allocation (mvae_comm_alloc_signals), zero-initialised
struct Signals {
    unsigned int flag[kChannels][2][kMaxCtas][kMaxWorld];   // flag[channel][slot][cta][writer rank]
    unsigned int count[kChannels][kMaxCtas];                // barriers this CTA has passed (only its owner touches it)
    unsigned int error;                                     // != 0: a wait timed out
};

struct Peers {
    float* buf[kMaxWorld];
    Signals* sig[kMaxWorld];
};

// element ranges of one exchange, in 16-byte units relative to the buffers' base
struct Ranges {
    int n;
    long long lo4[kMaxRanges], n4[kMaxRanges];
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
__device__ __forceinline__ void cta_barrier(const Peers& P, int ch, int rank, int world, unsigned int val,
                                            unsigned long long timeout_ns) {
    __syncthreads();                                    // this CTA's stores precede the release below (cumulativity)
    if (threadIdx.x < world) {
        const int slot = val & 1;                       // a peer can be at most one barrier ahead: two slots never collide
        st_release_sys(&P.sig[threadIdx.x]->flag[ch][slot][blockIdx.x][rank], val);
        const unsigned int* mine = &P.sig[rank]->flag[ch][slot][blockIdx.x][threadIdx.x];
        const unsigned long long t0 = globaltimer();
        while (ld_acquire_sys(mine) != val) {
            if (globaltimer() - t0 > timeout_ns) { P.sig[rank]->error = 1u; break; }
        }
    }
    __syncthreads();
}

// U iterations of the grid-stride loops are issued together: every load of a batch (U x WORLD 16-byte loads per thread) is in
// flight before the first add -- one NVLink round trip per batch instead of one per element (the loads are volatile asm, the
// compiler never overlaps two iterations by itself)
template <int WORLD>
struct Unroll { static constexpr int U = WORLD <= 2 ? 8 : WORLD <= 4 ? 4 : 2; };

// rank r sums slice r of the range over all ranks' buffers, in rank order, and writes the sum into EVERY rank's buffer (its
// own with a plain store, the peers' with posted stores over NVLink): slice r of any buffer is read and written by rank r
// only, each thread reads its elements everywhere before it writes them anywhere, so the exchange is in place
template <int WORLD>
__device__ __forceinline__ void reduce_and_broadcast_slice(const Peers& P, int rank, long long lo4, long long n4) {
    constexpr int U = Unroll<WORLD>::U;
    const long long chunk = (n4 + WORLD - 1) / WORLD;
    const long long stride = (long long)gridDim.x * kThreads;
    const long long lo = lo4 + rank * chunk, hi = lo4 + min((rank + 1) * chunk, n4);
    for (long long i0 = lo + (long long)blockIdx.x * kThreads + threadIdx.x; i0 < hi; i0 += U * stride) {
        float4 v[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
#pragma unroll
                for (int r = 0; r < WORLD; ++r) v[u][r] = ld16(P.buf[r] + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                float4 a = v[u][0];                                      // rank order 0..WORLD-1 whoever owns the slice
#pragma unroll
                for (int r = 1; r < WORLD; ++r) { a.x += v[u][r].x; a.y += v[u][r].y; a.z += v[u][r].z; a.w += v[u][r].w; }
#pragma unroll
                for (int k = 0; k < WORLD; ++k)                          // own buffer first, then the peers, staggered by rank
                    *reinterpret_cast<float4*>(P.buf[(rank + k) % WORLD] + 4 * i) = a;
            }
        }
    }
}

// barrier | pull + sum + push of this rank's slice | barrier.  The second barrier orders every pushed store before the kernel
// ends on the receiving rank (release / acquire at system scope through the signal words) and tells this rank that no peer
// still reads its buffer.
template <int WORLD>
__global__ void __launch_bounds__(kThreads) allreduce_kernel(const Peers P, const Ranges R, const int channel, const int rank,
                                                             const unsigned long long timeout_ns) {
    Signals* self = P.sig[rank];
    const unsigned int base = self->count[channel][blockIdx.x];     // written by this CTA only, in the previous launch
    cta_barrier(P, channel, rank, WORLD, base + 1, timeout_ns);
    for (int k = 0; k < R.n; ++k) reduce_and_broadcast_slice<WORLD>(P, rank, R.lo4[k], R.n4[k]);
    cta_barrier(P, channel, rank, WORLD, base + 2, timeout_ns);
    if (threadIdx.x == 0) self->count[channel][blockIdx.x] = base + 2;
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

extern "C" int mvae_comm_allreduce(float* const* bufs, void* const* signals, int rank, int world, int nranges,
                                   const long long* lo, const long long* n, int channel, int ctas, mvae_stream_t stream) {
    MVAE_REQUIRE(bufs && signals && lo && n, "mvae_comm_allreduce: null argument");
    MVAE_REQUIRE(world >= 2 && world <= comm::kMaxWorld && rank >= 0 && rank < world, "mvae_comm_allreduce: world %d rank %d", world, rank);
    MVAE_REQUIRE(nranges >= 1 && nranges <= comm::kMaxRanges, "mvae_comm_allreduce: %d ranges (1..%d)", nranges, comm::kMaxRanges);
    MVAE_REQUIRE(channel >= 0 && channel < comm::kChannels, "mvae_comm_allreduce: channel %d (0..%d)", channel, comm::kChannels - 1);
    comm::Ranges R = {};
    R.n = nranges;
    for (int k = 0; k < nranges; ++k) {
        MVAE_REQUIRE(lo[k] >= 0 && n[k] > 0 && (lo[k] % 4) == 0 && (n[k] % 4) == 0,
                     "mvae_comm_allreduce: range %d = [%lld, +%lld) must be multiples of 4 floats", k, lo[k], n[k]);
        R.lo4[k] = lo[k] / 4;
        R.n4[k] = n[k] / 4;
    }
    comm::Peers P = {};
    for (int r = 0; r < world; ++r) {
        MVAE_REQUIRE(bufs[r] && signals[r] && (reinterpret_cast<uintptr_t>(bufs[r]) & 15) == 0,
                     "mvae_comm_allreduce: buffer of rank %d is null or not 16-byte aligned", r);
        P.buf[r] = bufs[r];
        P.sig[r] = static_cast<comm::Signals*>(signals[r]);
    }
    if (ctas <= 0) ctas = env_int("MVAE_COMM_CTAS", 128);
    if (ctas > comm::kMaxCtas) ctas = comm::kMaxCtas;
    const unsigned long long timeout_ns = 1000000ull * (unsigned long long)env_int("MVAE_COMM_TIMEOUT_MS", 20000);
    cudaStream_t s = as_stream(stream);
    switch (world) {
#define MVAE_COMM_CASE(W) \
        case W: MVAE_CUDA(launch_pdl_ex(false, comm::allreduce_kernel<W>, dim3(ctas), dim3(comm::kThreads), 0, s, P, R, channel, rank, timeout_ns)); break;
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
