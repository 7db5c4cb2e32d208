// Device code of the N-sized streaming kernels (csrc/kry_vec.cu): elementwise updates, block dot / block axpy /
// combination, dense GEMV, diagonal operator, rot90.  A header of its own so that the CPU test tier can compile
// exactly these kernels for the host over the CUDA execution emulator (tests/csrc/cuda_emul, tests/test_vec_emul_cpu.py).
#pragma once
#include "kry_common.cuh"

#define KRY_JT 8   // basis vectors per register tile of the block dot

// z = a*x + b*y  (y may be NULL).  |a| == 1 or |b| == 1 keeps the reference's exact arithmetic (x + y, x - y: a
// single rounding -- the FMA form is also a single rounding then), so one kernel serves both.
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS) axpby_kernel(long long n, double a, const T* x, double b,
                                                           const T* y, T* z) {
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double xv[VEC], yv[VEC], zv[VEC];
        VecIO<T, VEC>::loadrw(x, i, xv);
        if (y) {
            VecIO<T, VEC>::loadrw(y, i, yv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) zv[u] = a * xv[u] + b * yv[u];
        } else {
#pragma unroll
            for (int u = 0; u < VEC; ++u) zv[u] = a * xv[u];
        }
        VecIO<T, VEC>::store(z, i, zv);
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
            double r = a * (double)x[i];
            if (y) r += b * (double)y[i];
            z[i] = (T)r;
        }
    }
}

// y += sign * coef[0] * x   (the coefficient lives on the device)
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS) axpy_dev_kernel(long long n, const double* coef, double sign,
                                                              const T* x, T* y) {
    const double c = sign * coef[0];
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double xv[VEC], yv[VEC];
        VecIO<T, VEC>::loadrw(x, i, xv);
        VecIO<T, VEC>::loadrw(y, i, yv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) yv[u] = fma(c, xv[u], yv[u]);
        VecIO<T, VEC>::store(y, i, yv);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x)
            y[i] = (T)fma(c, (double)x[i], (double)y[i]);
}

// out = mul * x / s[0]  (divide)  or  mul * x * s[0]
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS) scale_dev_kernel(long long n, const double* s, int divide,
                                                               double mul, const T* x, T* out) {
    const double sv = s[0];
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double xv[VEC];
        VecIO<T, VEC>::loadrw(x, i, xv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) xv[u] = divide ? (mul * xv[u]) / sv : (mul * xv[u]) * sv;
        VecIO<T, VEC>::store(out, i, xv);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
            double v = mul * (double)x[i];
            out[i] = (T)(divide ? v / sv : v * sv);
        }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS) diag_mul_kernel(long long n, const T* d, const T* x, T* y) {
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double dv[VEC], xv[VEC];
        VecIO<T, VEC>::load(d, i, dv);
        VecIO<T, VEC>::loadrw(x, i, xv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) xv[u] = dv[u] * xv[u];
        VecIO<T, VEC>::store(y, i, xv);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x)
            y[i] = (T)((double)d[i] * (double)x[i]);
}

// y = i * x for interleaved complex data (x, y: n complex numbers = 2n reals, may not alias):
// y[2k] = -x[2k+1], y[2k+1] = x[2k]  (twin storage of complex bases)
template <typename T>
__global__ void __launch_bounds__(KRY_THREADS) rot90_kernel(long long n, const T* __restrict__ x, T* y) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const T re = x[2 * k], im = x[2 * k + 1];
        y[2 * k] = -im;
        y[2 * k + 1] = re;
    }
}

// block dot: out[j] = <V_j, q>, j < nv.  KRY_JT basis vectors per register tile.
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2)
block_dot_kernel(long long n, const T* __restrict__ V, long long ldv, int nv, const T* q,
                 double* partials, unsigned int* ticket, double* out, int post, double* acc_out) {
    __shared__ double sm[32];
    __shared__ bool last;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int jb = 0; jb < nv; jb += KRY_JT) {
        double acc[KRY_JT];
#pragma unroll
        for (int t = 0; t < KRY_JT; ++t) acc[t] = 0.0;
        for (long long i = i0; i < nvec; i += stride) {
            double qv[VEC];
            VecIO<T, VEC>::loadrw(q, i, qv);
            double vv[KRY_JT][VEC];
#pragma unroll
            for (int t = 0; t < KRY_JT; ++t) {
                int j = jb + t;
                j = j < nv ? j : nv - 1;  // clamped duplicate loads hit L1; their sums are discarded
                VecIO<T, VEC>::load(V + (long long)j * ldv, i, vv[t]);
            }
#pragma unroll
            for (int t = 0; t < KRY_JT; ++t)
#pragma unroll
                for (int u = 0; u < VEC; ++u) acc[t] = fma(vv[t][u], qv[u], acc[t]);
        }
        if (blockIdx.x == 0) {  // scalar tail
            for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
                double qe = (double)q[i];
#pragma unroll
                for (int t = 0; t < KRY_JT; ++t) {
                    int j = jb + t;
                    if (j < nv) acc[t] = fma((double)V[(long long)j * ldv + i], qe, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < KRY_JT; ++t) {
            double s = kry_block_sum(acc[t], sm);
            if (threadIdx.x == 0 && jb + t < nv)
                partials[(long long)(jb + t) * KRY_MAX_PARTIAL_BLOCKS + blockIdx.x] = s;
        }
    }
    // last CTA to finish reduces all partials in a fixed order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        {   // one warp per basis vector, lanes stride over the CTAs (fixed order, parallel over j)
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int j = w; j < nv; j += nw) {
                double s = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32)
                    s += __ldcg(partials + (long long)j * KRY_MAX_PARTIAL_BLOCKS + b);
                s = kry_warp_sum(s);
                if (lane == 0) {
                    if (post == 1) s = sqrt(fabs(s));   // sqrt(||ip||_2) of a 1x1 matrix, utils.py:238
                    out[j] = s;
                    if (acc_out) acc_out[j] += s;
                }
            }
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// q += sign * sum_j coef[j] V_j   /   COMBINE: out = x0 + sum_j coef[j] V_j
template <typename T, int VEC, bool COMBINE>
__global__ void __launch_bounds__(KRY_THREADS, 2)
block_axpy_kernel(long long n, const T* __restrict__ V, long long ldv, int nv, const double* coef,
                  double sign, const T* x0, T* q) {
#ifdef KRY_EMUL
    double* sc = reinterpret_cast<double*>(kry_emul_dynamic_smem());
#else
    extern __shared__ double sc[];
#endif
    for (int j = threadIdx.x; j < nv; j += blockDim.x) sc[j] = sign * coef[j];
    __syncthreads();
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double qv[VEC];
        if (COMBINE) {
            if (x0) VecIO<T, VEC>::loadrw(x0, i, qv);
            else {
#pragma unroll
                for (int u = 0; u < VEC; ++u) qv[u] = 0.0;
            }
        } else {
            VecIO<T, VEC>::loadrw(q, i, qv);
        }
        if (COMBINE) {
            // reference order (linsys.py:947-948): yk = V.dot(yy) first, then x0 + yk
            double s[VEC];
#pragma unroll
            for (int u = 0; u < VEC; ++u) s[u] = 0.0;
            for (int jb = 0; jb < nv; jb += 4) {
                double vv[4][VEC];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    int j = jb + t < nv ? jb + t : nv - 1;
                    VecIO<T, VEC>::load(V + (long long)j * ldv, i, vv[t]);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (jb + t < nv) {
#pragma unroll
                        for (int u = 0; u < VEC; ++u) s[u] = fma(sc[jb + t], vv[t][u], s[u]);
                    }
            }
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[u] += s[u];
        } else {
            for (int jb = 0; jb < nv; jb += 4) {
                double vv[4][VEC];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    int j = jb + t < nv ? jb + t : nv - 1;
                    VecIO<T, VEC>::load(V + (long long)j * ldv, i, vv[t]);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (jb + t < nv) {
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[u] = fma(sc[jb + t], vv[t][u], qv[u]);
                    }
            }
        }
        VecIO<T, VEC>::store(q, i, qv);
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
            double s = 0.0, base;
            if (COMBINE) base = x0 ? (double)x0[i] : 0.0;
            else base = (double)q[i];
            if (COMBINE) {
                for (int j = 0; j < nv; ++j) s = fma(sc[j], (double)V[(long long)j * ldv + i], s);
                q[i] = (T)(base + s);
            } else {
                for (int j = 0; j < nv; ++j) base = fma(sc[j], (double)V[(long long)j * ldv + i], base);
                q[i] = (T)base;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(KRY_THREADS) gemv_kernel(long long m, long long n, const T* __restrict__ A,
                                                          long long lda, const T* __restrict__ x, T* y) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < m; r += nwarps) {
        const T* row = A + r * lda;
        double acc = 0.0;
        for (long long c = lane; c < n; c += 32) acc = fma((double)row[c], (double)x[c], acc);
        acc = kry_warp_sum(acc);
        if (lane == 0) y[r] = (T)acc;
    }
}
