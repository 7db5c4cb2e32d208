// Device code of kry_gram / kry_block_trsm (csrc/kry_block.cu): the two kernels of the CholQR2 set-up of the
// deflation projector.  A header of its own so that the CPU test tier can run a CholQR2 round with exactly these
// kernels over the CUDA execution emulator (tests/csrc/cuda_emul, tests/test_block_emul_cpu.py).
#pragma once
#include "kry_common.cuh"

#define GRAM_E 128        // rows per staged chunk
#define GRAM_LD 130       // shared-memory row stride (doubles): 16-byte aligned rows, skewed banks
#define GRAM_MAXV 64      // kx + ky (or kx when Y == X) staged vectors at most
#define GRAM_BPW 2        // 4x4 output blocks per warp
#define GRAM_THREADS 512  // 16 warps x 2 blocks: up to 32 blocks = 512 outputs (e.g. 20 x 20)

#ifndef KRY_EMUL   // (the CPU tier's execution emulator supplies host versions of the cp.async wrappers)
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? BYTES : 0;        // src-size 0: the destination is zero filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(s), "l"(gsrc), "n"(BYTES), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#endif

// rows e, e+1 of a staged vector as doubles
template <typename T> __device__ __forceinline__ double2 lds_pair(const T* p);
template <> __device__ __forceinline__ double2 lds_pair<double>(const double* p) {
    return *reinterpret_cast<const double2*>(p);
}
template <> __device__ __forceinline__ double2 lds_pair<float>(const float* p) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    return make_double2((double)v.x, (double)v.y);
}

// Shared memory: two buffers of [nvec][GRAM_LD] elements of T; the next 128-row chunk is fetched with
// cp.async while the current one is consumed.  same != 0 (X^H X): only the 4x4 blocks on and above the
// block diagonal are computed and mirrored at the end.
template <typename T>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_kernel(long long n, const T* __restrict__ X, long long ldx, int kx, const T* __restrict__ Y, long long ldy, int ky,
            int same, double* partials, unsigned int* ticket, double* out) {
#ifdef KRY_EMUL
    unsigned char* sh_raw = kry_emul_dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char sh_raw[];
#endif
    __shared__ bool last;
    __shared__ unsigned char blk_i[GRAM_BPW * (GRAM_THREADS / 32)], blk_j[GRAM_BPW * (GRAM_THREADS / 32)];
    T* sh = reinterpret_cast<T*>(sh_raw);
    const int nvec = same ? kx : kx + ky;
    const size_t bufsz = (size_t)nvec * GRAM_LD;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = GRAM_THREADS >> 5;
    const int nbx = (kx + 3) >> 2, nby = (ky + 3) >> 2;
    const int nb = same ? nbx * (nbx + 1) / 2 : nbx * nby;
    if (threadIdx.x == 0) {
        int p = 0;
        for (int i = 0; i < nbx; ++i)
            for (int j = same ? i : 0; j < nby; ++j) {
                blk_i[p] = (unsigned char)i;
                blk_j[p] = (unsigned char)j;
                ++p;
            }
    }
    double acc[GRAM_BPW][16];
#pragma unroll
    for (int b = 0; b < GRAM_BPW; ++b)
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[b][t] = 0.0;
    const long long nchunks = (n + GRAM_E - 1) / GRAM_E;

    auto stage = [&](long long c, int buf) {
        const long long r0 = c * GRAM_E;
        T* dst = sh + (size_t)buf * bufsz;
        for (int idx = threadIdx.x; idx < nvec * GRAM_E; idx += GRAM_THREADS) {
            const int v = idx / GRAM_E, e = idx - v * GRAM_E;
            const T* src = (v < kx) ? (X + (long long)v * ldx) : (Y + (long long)(v - kx) * ldy);
            const bool valid = (r0 + e) < n;
            cp_async_zfill<sizeof(T)>(dst + (size_t)v * GRAM_LD + e, valid ? (const void*)(src + r0 + e) : (const void*)X,
                                      valid);
        }
        cp_async_commit();
    };

    int buf = 0;
    long long c = blockIdx.x;
    if (c < nchunks) stage(c, 0);
    __syncthreads();                               // block table visible
    for (; c < nchunks; c += gridDim.x) {
        const long long nxt = c + gridDim.x;
        if (nxt < nchunks) {
            stage(nxt, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const T* sx = sh + (size_t)buf * bufsz;
        const T* sy = same ? sx : sx + (size_t)kx * GRAM_LD;
#pragma unroll
        for (int bi = 0; bi < GRAM_BPW; ++bi) {
            const int b = w + bi * nw;
            if (b < nb) {                         // uniform per warp
                const int a0 = (int)blk_i[b] << 2, b0 = (int)blk_j[b] << 2;
#pragma unroll
                for (int step = 0; step < GRAM_E / 64; ++step) {
                    const int e = (step << 6) + (lane << 1);
                    double2 xa[4], yb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int va = a0 + i < kx ? a0 + i : kx - 1;        // (clamped rows are discarded below)
                        const int vb = b0 + i < ky ? b0 + i : ky - 1;
                        xa[i] = lds_pair<T>(sx + (size_t)va * GRAM_LD + e);
                        yb[i] = lds_pair<T>(sy + (size_t)vb * GRAM_LD + e);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[bi][i * 4 + j] = fma(xa[i].x, yb[j].x, acc[bi][i * 4 + j]);
                            acc[bi][i * 4 + j] = fma(xa[i].y, yb[j].y, acc[bi][i * 4 + j]);
                        }
                }
            }
        }
        __syncthreads();                          // chunk consumed: its buffer may be refilled
        buf ^= 1;
    }
    // per-CTA partials: partials[blockIdx][kx*ky]
    double* mine = partials + (size_t)blockIdx.x * (size_t)(kx * ky);
#pragma unroll
    for (int bi = 0; bi < GRAM_BPW; ++bi) {
        const int b = w + bi * nw;
        if (b < nb) {
            const int a0 = (int)blk_i[b] << 2, b0 = (int)blk_j[b] << 2;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const double s = kry_warp_sum(acc[bi][t]);
                const int i = a0 + (t >> 2), j = b0 + (t & 3);
                if (lane == 0 && i < kx && j < ky) mine[i * ky + j] = s;
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        for (int o = threadIdx.x; o < kx * ky; o += GRAM_THREADS) {
            int i = o / ky, j = o - i * ky;
            // X^H X: the strictly lower 4x4 blocks were not computed; entry (i, j) equals entry (j, i)
            const int src = (same && (i >> 2) > (j >> 2)) ? j * ky + i : o;
            double s = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(partials + (size_t)b * (size_t)(kx * ky) + src);   // fixed order
            out[o] = s;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// Q = X R^-1 for upper triangular R: the inverse is formed once per CTA in shared memory (d <= 32), then
// every row of the (vector-major) block is a dense product with independent FMAs -- no division and no
// dependency chain per row.
template <typename T, int DMAX>
__global__ void __launch_bounds__(KRY_THREADS, (DMAX <= 24 ? 2 : 1))
block_trsm_kernel(long long n, const T* X, long long ldx, int d, const double* __restrict__ R, T* Q, long long ldq) {
    __shared__ double Rs[DMAX * DMAX];     // (leading dimension DMAX: compile-time offsets in the unrolled product)
    __shared__ double Ri[DMAX * DMAX];
    for (int idx = threadIdx.x; idx < DMAX * DMAX; idx += blockDim.x) {
        const int i = idx / DMAX, j = idx - i * DMAX;
        Rs[idx] = (i < d && j < d) ? R[i * d + j] : 0.0;
        Ri[idx] = 0.0;
    }
    __syncthreads();
    // column j of R^-1 by back substitution: R z = e_j  (thread j; d <= 32)
    if ((int)threadIdx.x < d) {
        const int j = threadIdx.x;
        for (int i = j; i >= 0; --i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int l = i + 1; l <= j; ++l) s = fma(-Rs[i * DMAX + l], Ri[l * DMAX + j], s);
            Ri[i * DMAX + j] = s / Rs[i * DMAX + i];
        }
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double x[DMAX];
#pragma unroll
        for (int j = 0; j < DMAX; ++j) x[j] = (j < d) ? (double)X[(long long)j * ldx + i] : 0.0;
#pragma unroll
        for (int j = 0; j < DMAX; ++j) {
            if (j < d) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l <= j; ++l) s = fma(x[l], Ri[l * DMAX + j], s);
                Q[(long long)j * ldq + i] = (T)s;
            }
        }
    }
}

