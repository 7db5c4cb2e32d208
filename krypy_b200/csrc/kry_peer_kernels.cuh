// Device code of kry_peer_allreduce / kry_peer_barrier / kry_halo_gather (csrc/kry_peer.cu).  A header of its own
// so that the CPU test tier can run these kernels over emulated ranks (tests/csrc/cuda_emul,
// tests/csrc/cgdist_emul_host.cpp, tests/test_cgdist_emul_cpu.py).
#pragma once
#include "kry_common.cuh"

#define PEER_MAX_RANKS 16
#define PEER_SLOT 64     // doubles per rank per parity

#ifndef KRY_EMUL   // (the CPU tier's execution emulator supplies host versions)
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

#endif

// slots layout on every rank: [2 parities][world][PEER_SLOT] doubles; flags: [world] u64
__global__ void __launch_bounds__(128)
peer_allreduce_kernel(int world, int rank, unsigned long long* epoch_dev, int n, double* inout,
                      double* const* peer_slots, unsigned long long* const* peer_flags, int post, double* acc) {
    __shared__ int timed_out;
    const int tid = threadIdx.x;
    // the operation counter lives in device memory (one increment per peer operation, in stream
    // order, identical on every rank): the launch arguments never change, so the iteration can
    // be replayed from a CUDA graph
    const unsigned long long epoch = *epoch_dev + 1ull;
    if (tid == 0) timed_out = 0;
    __syncthreads();
    const size_t par = (size_t)(epoch & 1ull) * (size_t)world * PEER_SLOT;
    // 1. publish my partials into every rank's slot array (including my own)
    for (int idx = tid; idx < world * n; idx += blockDim.x) {
        const int r = idx / n, i = idx - r * n;
        peer_slots[r][par + (size_t)rank * PEER_SLOT + i] = inout[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(peer_flags[tid] + rank, epoch);
    // 2. wait until every rank has published this epoch into MY arrays
    if (tid < world) {
        const unsigned long long* f = peer_flags[rank] + tid;
        // bounded spin (10 s): a peer that died must not wedge this GPU; the result is poisoned instead
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < epoch) {
            if (global_timer_ns() - t0 > 10000000000ull) {
                timed_out = 1;
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    // 3. fixed rank-order sum: bitwise identical on every rank
    const double* mine = peer_slots[rank] + par;
    for (int i = tid; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += ld_volatile_f64(mine + (size_t)r * PEER_SLOT + i);
        if (post == 1) s = sqrt(fabs(s));
        if (timed_out) s = __longlong_as_double(0x7ff8000000000000ll);   // NaN: peer never arrived
        inout[i] = s;
        if (acc) acc[i] += s;
    }
    if (tid == 0) *epoch_dev = epoch;
}

template <typename T>
__global__ void __launch_bounds__(KRY_THREADS)
halo_gather_kernel(long long nhalo, const T* const* peer_bases, long long elem_offset,
                   const int* __restrict__ halo_peer, const int* __restrict__ halo_off, const double* div, T* dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double d = div ? div[0] : 1.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhalo; i += stride) {
        const T* src = peer_bases[__ldg(halo_peer + i)] + elem_offset;
        // remote HBM over NVLink: plain (non-.nc) load, the peer rewrites this buffer between uses
        const T v = *(const volatile T*)(src + __ldg(halo_off + i));
        dst[i] = div ? (T)((double)v / d) : v;
    }
}

