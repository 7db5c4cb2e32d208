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
