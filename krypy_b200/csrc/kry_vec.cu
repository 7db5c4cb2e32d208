// N-sized streaming kernels: elementwise updates, tall-skinny block dot /
// block axpy / reconstruction, dense GEMV, diagonal operator.
// All HBM-bound: 16-byte vectorised accesses, grid sized in multiples of the SM
// count, deterministic two-stage reductions (fixed tree, last-CTA finish).
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_vec_kernels.cuh"     // the kernels (device code only); below: the launchers and the C ABI

static inline int stream_grid(const kry_ctx* ctx, long long nvec, int per_sm) {
    long long need = (nvec + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}


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
// dense row-major GEMV, warp per row
// ---------------------------------------------------------------------------

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
