// Shared device/host helpers of the krypy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/krypy_b200.h"

namespace cg = cooperative_groups;

#define KRY_THREADS 256
#define KRY_MAX_PARTIAL_BLOCKS 2048   // upper bound on gridDim.x of any reducing kernel
#define KRY_MAX_SLOTS 64              // reduction slots (basis vectors) per pass

struct kry_ctx {
    int device;
    cudaStream_t stream;
    int sm_count;
    int cc;
    long long l2_bytes;
    long long smem_optin;
    int coop;
    double* d_partials;      // [2][KRY_MAX_SLOTS][KRY_MAX_PARTIAL_BLOCKS] ping-pong scratch
    unsigned int* d_ticket;  // last-block tickets (zeroed; kernels reset them)
    double* h_mailbox;       // pinned + mapped
    double* d_mailbox;       // device alias
    long long launches;
    int orth_blocks_f64, orth_blocks_f32;     // co-resident grid sizes (cached)
    int proj_blocks_f64, proj_blocks_f32;
};

void kry_set_error(const char* fmt, ...);

#define KRY_CHECK_CUDA(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            kry_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                          __FILE__, __LINE__);                                       \
            return KRY_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

#define KRY_REQUIRE(cond, msg)                                                       \
    do {                                                                             \
        if (!(cond)) {                                                               \
            kry_set_error("%s: requirement failed: %s", __func__, msg);              \
            return KRY_ERR_ARG;                                                      \
        }                                                                            \
    } while (0)

#define KRY_LAUNCHED(ctx)                                                            \
    do {                                                                             \
        (ctx)->launches++;                                                           \
        KRY_CHECK_CUDA(cudaGetLastError());                                          \
    } while (0)

static inline bool kry_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// ---------------------------------------------------------------------------
// vector access: VEC elements of T per 16-byte (or scalar) access
// ---------------------------------------------------------------------------
template <typename T, int VEC> struct Pack;
template <> struct Pack<double, 2> { typedef double2 type; };
template <> struct Pack<double, 1> { typedef double type; };
template <> struct Pack<float, 4> { typedef float4 type; };
template <> struct Pack<float, 1> { typedef float type; };

template <typename T> struct VecWidth;
template <> struct VecWidth<double> { static const int value = 2; };
template <> struct VecWidth<float> { static const int value = 4; };

template <typename T, int VEC>
struct VecIO {
    typedef typename Pack<T, VEC>::type P;
    // streaming read-only load (read once: do not pollute L1)
    static __device__ __forceinline__ void load(const T* __restrict__ p, long long i, double (&v)[VEC]) {
        P t = __ldg(reinterpret_cast<const P*>(p) + i);
        const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
        for (int u = 0; u < VEC; ++u) v[u] = (double)e[u];
    }
    // plain load (data may have been written earlier in this kernel)
    static __device__ __forceinline__ void loadrw(const T* p, long long i, double (&v)[VEC]) {
        P t = *(reinterpret_cast<const P*>(p) + i);
        const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
        for (int u = 0; u < VEC; ++u) v[u] = (double)e[u];
    }
    static __device__ __forceinline__ void store(T* p, long long i, const double (&v)[VEC]) {
        P t;
        T* e = reinterpret_cast<T*>(&t);
#pragma unroll
        for (int u = 0; u < VEC; ++u) e[u] = (T)v[u];
        *(reinterpret_cast<P*>(p) + i) = t;
    }
};

// ---------------------------------------------------------------------------
// deterministic block reductions (fixed shuffle tree + fixed warp order)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double kry_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in every thread.  sm: >= 32 doubles of scratch.
__device__ __forceinline__ double kry_block_sum(double v, double* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = kry_warp_sum(v);
    __syncthreads();  // protect sm reuse
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = (lane < nw) ? sm[lane] : 0.0;
    r = kry_warp_sum(r);
    return r;
}

// Sum partials[0..nblocks) in a fixed order with the whole CTA; valid in every thread.
__device__ __forceinline__ double kry_reduce_partials(const volatile double* partials, int nblocks,
                                                      double* sm) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v += partials[b];
    return kry_block_sum(v, sm);
}

// ---------------------------------------------------------------------------
// NVLink peer exchange helpers (shared by kry_dist.cu and the cooperative kernels)
// ---------------------------------------------------------------------------
#define PEER_MAX_RANKS 16
#define PEER_SLOT 64

struct PeerArgs {
    int world, rank;
    unsigned long long* epoch_dev;
    double* const* slots;                 // [rank] -> that rank's [2][world][PEER_SLOT] doubles
    unsigned long long* const* flags;     // [rank] -> that rank's [world] u64
};

#ifndef KRY_EMUL   // (the CPU tier's execution emulator supplies host versions, tests/csrc/cuda_emul)
__device__ __forceinline__ void dst_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double dld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dglobal_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

#endif

__device__ __forceinline__ double nan_f64() { return __longlong_as_double(0x7ff8000000000000ll); }

// the calling CTA stores vals[0..n) into every rank's slot row for `epoch` and releases the flag
__device__ __forceinline__ void peer_publish(const PeerArgs& pa, unsigned long long epoch, const double* vals_smem,
                                             int n) {
    const size_t par = (size_t)(epoch & 1ull) * (size_t)pa.world * PEER_SLOT;
    for (int idx = threadIdx.x; idx < pa.world * n; idx += blockDim.x) {
        const int r = idx / n, i = idx - r * n;
        pa.slots[r][par + (size_t)pa.rank * PEER_SLOT + i] = vals_smem[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pa.world) dst_release_sys(pa.flags[threadIdx.x] + pa.rank, epoch);
}

// every thread of the CTA returns once all ranks have published `epoch` (false on time-out)
__device__ __forceinline__ bool peer_wait(const PeerArgs& pa, unsigned long long epoch, int* flag_smem) {
    if (threadIdx.x == 0) *flag_smem = 1;
    __syncthreads();
    if (threadIdx.x < pa.world) {
        const unsigned long long* f = pa.flags[pa.rank] + threadIdx.x;
        const unsigned long long t0 = dglobal_timer_ns();
        while (dld_acquire_sys(f) < epoch) {
            if (dglobal_timer_ns() - t0 > 10000000000ull) {
                *flag_smem = 0;
                break;
            }
        }
    }
    // the polling threads' ld.acquire.sys + the CTA barrier order every later access of the CTA
    // after the peers' releases (causality is cumulative): no CTA-wide system fence needed
    __syncthreads();
    return *flag_smem != 0;
}

// fixed rank-order sum of slot i of `epoch` (my own slot array)
__device__ __forceinline__ double peer_sum(const PeerArgs& pa, unsigned long long epoch, int i) {
    const double* mine = pa.slots[pa.rank] + (size_t)(epoch & 1ull) * (size_t)pa.world * PEER_SLOT;
    double s = 0.0;
    for (int r = 0; r < pa.world; ++r) s += dld_volatile_f64(mine + (size_t)r * PEER_SLOT + i);
    return s;
}

// Row-partitioned run: turn the local sums c_s[0..cnt) (identical in every CTA) into global sums.
// CTA 0 stores them into every peer's slot array and releases its flag; every CTA acquires all
// flags and sums the per-rank partials in rank order (bitwise identical on all ranks).
__device__ __forceinline__ void peer_exchange(const PeerArgs& pa, unsigned long long epoch, double* c_s, int cnt,
                                              double* stage, int* okflag) {
    if (blockIdx.x == 0) peer_publish(pa, epoch, c_s, cnt);
    const bool ok = peer_wait(pa, epoch, okflag);
    const double* mine = pa.slots[pa.rank] + (size_t)(epoch & 1ull) * (size_t)pa.world * PEER_SLOT;
    for (int idx = threadIdx.x; idx < pa.world * cnt; idx += blockDim.x) {
        const int r = idx / cnt, j = idx - r * cnt;
        stage[r * PEER_SLOT + j] = dld_volatile_f64(mine + (size_t)r * PEER_SLOT + j);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        double sum = 0.0;
        for (int r = 0; r < pa.world; ++r) sum += stage[r * PEER_SLOT + j];
        c_s[j] = ok ? sum : nan_f64();
    }
    __syncthreads();
}
