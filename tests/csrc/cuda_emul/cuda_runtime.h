// TEST INFRASTRUCTURE -- a minimal CUDA *execution* emulator for the CPU test tier (never part of the product).
//
// It lets g++ compile the device headers of krypy_b200/csrc unchanged and run a kernel with the CUDA execution
// model: every CUDA thread is an OS thread, every CTA is a process (so that `__shared__` variables -- compiled as
// function-local statics -- are per CTA), device memory is a MAP_SHARED mapping created before the CTAs are
// forked.  __syncthreads, warp shuffles, grid.sync, mbarriers and bulk copies are pthread barriers / mutex
// protected state.  What a green run proves: the indexing, the reduction plumbing and the synchronisation
// protocol of the kernel are right for the emulated grid.  It says nothing about performance or about
// hardware-specific behaviour (memory model, occupancy).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct float4 {
    float x, y, z, w;
};
struct float2 {
    float x, y;
};
struct uint3 {
    unsigned int x, y, z;
};
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int a = 1, unsigned int b = 1, unsigned int c = 1) : x(a), y(b), z(c) {}
};

extern thread_local uint3 threadIdx;   // per OS thread
extern uint3 blockIdx;                 // per process (one CTA per process)
extern dim3 blockDim, gridDim;

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
static inline double __dadd_rn(double a, double b) { return a + b; }      // (compile with -ffp-contract=off)
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, 8);
    return d;
}

// provided by the emulator runtime (tests/csrc/cplx_emul_host.cpp)
double __shfl_xor_sync(unsigned int mask, double v, int lane_mask);
void __syncthreads();
void __syncwarp();
void __threadfence();
void __threadfence_system();
unsigned int atomicAdd(unsigned int* p, unsigned int v);
unsigned char* kry_emul_dynamic_smem();
