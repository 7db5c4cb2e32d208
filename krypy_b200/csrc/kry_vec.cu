// N-sized streaming kernels: elementwise updates, tall-skinny block dot /
// block axpy / reconstruction, dense GEMV, diagonal operator.
// All HBM-bound: 16-byte vectorised accesses, grid sized in multiples of the SM
// count, deterministic two-stage reductions (fixed tree, last-CTA finish).
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

static inline int stream_grid(const kry_ctx* ctx, long long nvec, int per_sm) {
    long long need = (nvec + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------
// z = a*x + b*y
// ---------------------------------------------------------------------------
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

// a == 1 / b == +-1 special case keeps the reference's exact arithmetic
// (x + y, x - y: a single rounding) -- the FMA form above is also a single
// rounding for |a| == 1 or |b| == 1, so one kernel serves both.

template <typename T>
static int axpby_launch(kry_ctx* ctx, long long n, double a, const T* x, double b, const T* y, T* z) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(x) && kry_aligned16(z) && (!y || kry_aligned16(y));
    if (al) {
        int g = stream_grid(ctx, n / W, 8);
        axpby_kernel<T, W><<<g, KRY_THREADS, 0, ctx->stream>>>(n, a, x, b, y, z);
    } else {
        int g = stream_grid(ctx, n, 8);
        axpby_kernel<T, 1><<<g, KRY_THREADS, 0, ctx->stream>>>(n, a, x, b, y, z);
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

// ---------------------------------------------------------------------------
// y += sign*coef*x ; out = mul * x (/|*) s
// ---------------------------------------------------------------------------
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
// y[2k] = -x[2k+1], y[2k+1] = x[2k].  Complex systems run on the real kernels with every basis
// vector v stored next to its twin i*v ("complex by real embedding").
template <typename T>
__global__ void __launch_bounds__(KRY_THREADS) rot90_kernel(long long n, const T* __restrict__ x, T* y) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const T re = x[2 * k], im = x[2 * k + 1];
        y[2 * k] = -im;
        y[2 * k + 1] = re;
    }
}

// ---------------------------------------------------------------------------
// block dot: out[j] = <V_j, q>, j < nv.  JT basis vectors per register tile.
// ---------------------------------------------------------------------------
#define KRY_JT 8

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

// q += sign * sum_j coef[j] V_j   /   out = x0 + sum_j coef[j] V_j
template <typename T, int VEC, bool COMBINE>
__global__ void __launch_bounds__(KRY_THREADS, 2)
block_axpy_kernel(long long n, const T* __restrict__ V, long long ldv, int nv, const double* coef,
                  double sign, const T* x0, T* q) {
    extern __shared__ double sc[];
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

// ---------------------------------------------------------------------------
// dense row-major GEMV, warp per row
// ---------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------
template <typename T>
static int block_dot_launch(kry_ctx* ctx, long long n, const T* V, long long ldv, int nv, const T* q,
                            double* out, int post, double* acc) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(V) && kry_aligned16(q) && (ldv % W == 0);
    for (int j0 = 0; j0 < nv; j0 += KRY_MAX_SLOTS) {
        int cnt = nv - j0 < KRY_MAX_SLOTS ? nv - j0 : KRY_MAX_SLOTS;
        const T* Vj = V + (long long)j0 * ldv;
        if (al) {
            int g = stream_grid(ctx, n / W, 2);
            block_dot_kernel<T, W><<<g, KRY_THREADS, 0, ctx->stream>>>(
                n, Vj, ldv, cnt, q, ctx->d_partials, ctx->d_ticket, out + j0, post, acc ? acc + j0 : nullptr);
        } else {
            int g = stream_grid(ctx, n, 2);
            block_dot_kernel<T, 1><<<g, KRY_THREADS, 0, ctx->stream>>>(
                n, Vj, ldv, cnt, q, ctx->d_partials, ctx->d_ticket, out + j0, post, acc ? acc + j0 : nullptr);
        }
        KRY_LAUNCHED(ctx);
    }
    return KRY_OK;
}

template <typename T, bool COMBINE>
static int block_axpy_launch(kry_ctx* ctx, long long n, const T* V, long long ldv, int nv,
                             const double* coef, double sign, const T* x0, T* q) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(V) && kry_aligned16(q) && (ldv % W == 0) && (!x0 || kry_aligned16(x0));
    size_t smem = sizeof(double) * (size_t)(nv > 0 ? nv : 1);
    KRY_REQUIRE(smem <= 40000, "too many basis vectors in one block_axpy call");
    if (al) {
        int g = stream_grid(ctx, n / W, 4);
        block_axpy_kernel<T, W, COMBINE><<<g, KRY_THREADS, smem, ctx->stream>>>(n, V, ldv, nv, coef, sign, x0, q);
    } else {
        int g = stream_grid(ctx, n, 4);
        block_axpy_kernel<T, 1, COMBINE><<<g, KRY_THREADS, smem, ctx->stream>>>(n, V, ldv, nv, coef, sign, x0, q);
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

#define KRY_DISPATCH_VEC(T, KERNEL, n, al, ...)                                           \
    do {                                                                                  \
        const int W_ = VecWidth<T>::value;                                                \
        if (al) {                                                                         \
            int g_ = stream_grid(ctx, (n) / W_, 8);                                       \
            KERNEL<T, W_><<<g_, KRY_THREADS, 0, ctx->stream>>>(__VA_ARGS__);              \
        } else {                                                                          \
            int g_ = stream_grid(ctx, (n), 8);                                            \
            KERNEL<T, 1><<<g_, KRY_THREADS, 0, ctx->stream>>>(__VA_ARGS__);               \
        }                                                                                 \
        KRY_LAUNCHED(ctx);                                                                \
    } while (0)

extern "C" {

int kry_axpby(kry_ctx* ctx, int dtype, long long n, double a, const void* x, double b, const void* y,
              void* z) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && x && z, "bad arguments");
    KRY_REQUIRE(y || b == 0.0, "y is NULL but b != 0");
    if (n == 0) return KRY_OK;
    if (dtype == KRY_F64) return axpby_launch<double>(ctx, n, a, (const double*)x, b, (const double*)y, (double*)z);
    if (dtype == KRY_F32) return axpby_launch<float>(ctx, n, a, (const float*)x, b, (const float*)y, (float*)z);
    kry_set_error("kry_axpby: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_axpy_dev(kry_ctx* ctx, int dtype, long long n, const double* coef_dev, double sign, const void* x,
                 void* y) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && x && y && coef_dev, "bad arguments");
    if (n == 0) return KRY_OK;
    bool al = kry_aligned16(x) && kry_aligned16(y);
    if (dtype == KRY_F64) {
        KRY_DISPATCH_VEC(double, axpy_dev_kernel, n, al, n, coef_dev, sign, (const double*)x, (double*)y);
    } else if (dtype == KRY_F32) {
        KRY_DISPATCH_VEC(float, axpy_dev_kernel, n, al, n, coef_dev, sign, (const float*)x, (float*)y);
    } else {
        kry_set_error("kry_axpy_dev: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    return KRY_OK;
}

int kry_scale_dev(kry_ctx* ctx, int dtype, long long n, const double* s_dev, int divide, double mul,
                  const void* x, void* out) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && x && out && s_dev, "bad arguments");
    if (n == 0) return KRY_OK;
    bool al = kry_aligned16(x) && kry_aligned16(out);
    if (dtype == KRY_F64) {
        KRY_DISPATCH_VEC(double, scale_dev_kernel, n, al, n, s_dev, divide, mul, (const double*)x, (double*)out);
    } else if (dtype == KRY_F32) {
        KRY_DISPATCH_VEC(float, scale_dev_kernel, n, al, n, s_dev, divide, mul, (const float*)x, (float*)out);
    } else {
        kry_set_error("kry_scale_dev: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    return KRY_OK;
}

int kry_diag_mul(kry_ctx* ctx, int dtype, long long n, const void* d, const void* x, void* y) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && d && x && y, "bad arguments");
    if (n == 0) return KRY_OK;
    bool al = kry_aligned16(d) && kry_aligned16(x) && kry_aligned16(y);
    if (dtype == KRY_F64) {
        KRY_DISPATCH_VEC(double, diag_mul_kernel, n, al, n, (const double*)d, (const double*)x, (double*)y);
    } else if (dtype == KRY_F32) {
        KRY_DISPATCH_VEC(float, diag_mul_kernel, n, al, n, (const float*)d, (const float*)x, (float*)y);
    } else {
        kry_set_error("kry_diag_mul: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    return KRY_OK;
}

int kry_rot90(kry_ctx* ctx, int dtype, long long n, const void* x, void* y) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && x && y && x != y, "bad arguments (x and y must not alias)");
    if (n == 0) return KRY_OK;
    int g = stream_grid(ctx, n, 8);
    if (dtype == KRY_F64)
        rot90_kernel<double><<<g, KRY_THREADS, 0, ctx->stream>>>(n, (const double*)x, (double*)y);
    else if (dtype == KRY_F32)
        rot90_kernel<float><<<g, KRY_THREADS, 0, ctx->stream>>>(n, (const float*)x, (float*)y);
    else {
        kry_set_error("kry_rot90: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_block_dot(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, const void* q,
                  double* out_dev, int post, double* acc_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && nv >= 0 && out_dev, "bad arguments");
    if (nv == 0) return KRY_OK;
    KRY_REQUIRE(V && q, "NULL vector");
    if (dtype == KRY_F64)
        return block_dot_launch<double>(ctx, n, (const double*)V, ldv, nv, (const double*)q, out_dev, post, acc_dev);
    if (dtype == KRY_F32)
        return block_dot_launch<float>(ctx, n, (const float*)V, ldv, nv, (const float*)q, out_dev, post, acc_dev);
    kry_set_error("kry_block_dot: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_block_axpy(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv,
                   const double* coef_dev, double sign, void* q) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && nv >= 0 && q, "bad arguments");
    if (nv == 0 || n == 0) return KRY_OK;
    KRY_REQUIRE(V && coef_dev, "NULL argument");
    if (dtype == KRY_F64)
        return block_axpy_launch<double, false>(ctx, n, (const double*)V, ldv, nv, coef_dev, sign, nullptr, (double*)q);
    if (dtype == KRY_F32)
        return block_axpy_launch<float, false>(ctx, n, (const float*)V, ldv, nv, coef_dev, sign, nullptr, (float*)q);
    kry_set_error("kry_block_axpy: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_block_combine(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv,
                      const double* coef_dev, const void* x0, void* out) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && nv >= 0 && out, "bad arguments");
    if (n == 0) return KRY_OK;
    KRY_REQUIRE(nv == 0 || (V && coef_dev), "NULL argument");
    if (dtype == KRY_F64)
        return block_axpy_launch<double, true>(ctx, n, (const double*)V, ldv, nv, coef_dev, 1.0, (const double*)x0, (double*)out);
    if (dtype == KRY_F32)
        return block_axpy_launch<float, true>(ctx, n, (const float*)V, ldv, nv, coef_dev, 1.0, (const float*)x0, (float*)out);
    kry_set_error("kry_block_combine: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_gemv_dense(kry_ctx* ctx, int dtype, long long m, long long n, const void* A, long long lda,
                   const void* x, void* y) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(m >= 0 && n >= 0 && A && x && y && lda >= n, "bad arguments");
    if (m == 0) return KRY_OK;
    long long need = (m * 32 + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 8;
    int g = (int)(need < cap ? need : cap);
    if (dtype == KRY_F64)
        gemv_kernel<double><<<g, KRY_THREADS, 0, ctx->stream>>>(m, n, (const double*)A, lda, (const double*)x, (double*)y);
    else if (dtype == KRY_F32)
        gemv_kernel<float><<<g, KRY_THREADS, 0, ctx->stream>>>(m, n, (const float*)A, lda, (const float*)x, (float*)y);
    else {
        kry_set_error("kry_gemv_dense: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

}  // extern "C"
