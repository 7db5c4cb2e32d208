// Latency-bound device recurrences: GMRES Givens/Hessenberg update, triangular
// solve, MINRES sliding QR.  One small CTA each; inputs are staged into shared
// memory in parallel, one thread runs the (inherently serial) recurrence, the
// results go to device state and to the pinned host mailbox in parallel.
#include "kry_common.cuh"
#include "kry_small_core.h"
#include "kry_givens_dev.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_small_kernels.cuh"

extern "C" {

int kry_small_qr_apply(kry_ctx* ctx, int d, const double* Q_dev, const double* R_dev, const double* c_in_dev,
                       double* c_out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(d >= 0 && d <= 2048, "bad d");
    if (d == 0) return KRY_OK;
    KRY_REQUIRE(Q_dev && R_dev && c_in_dev && c_out_dev, "NULL argument");
    small_qr_apply_kernel<<<1, 128, sizeof(double) * 2 * (size_t)d, ctx->stream>>>(d, Q_dev, R_dev, c_in_dev, c_out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_givens_update(kry_ctx* ctx, int k, double* hcol_dev, double* rcol_dev, double* cs_dev, double* y_dev,
                      int mailbox_off) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && hcol_dev && rcol_dev && cs_dev && y_dev, "bad arguments");
    KRY_REQUIRE(k <= 2000, "k too large for the single-CTA Givens update (use restarts)");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 2 * k + 5 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    size_t smem = sizeof(double) * (size_t)(3 * k + 2);
    givens_kernel<<<1, 128, smem, ctx->stream>>>(k, hcol_dev, rcol_dev, cs_dev, y_dev, ctx->d_mailbox + mailbox_off);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_tri_solve(kry_ctx* ctx, int k, const double* R_dev, long long ldr, const double* y_dev, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && ldr >= k, "bad arguments");
    if (k == 0) return KRY_OK;
    KRY_REQUIRE(R_dev && y_dev && out_dev, "NULL argument");
    KRY_REQUIRE(k <= 6000, "k too large");
    tri_solve_kernel<false><<<1, 128, sizeof(double) * (size_t)k, ctx->stream>>>(k, R_dev, ldr, y_dev, out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_tri_solve_t(kry_ctx* ctx, int k, const double* Rt_dev, long long ldr, const double* y_dev, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && ldr >= k, "bad arguments");
    if (k == 0) return KRY_OK;
    KRY_REQUIRE(Rt_dev && y_dev && out_dev, "NULL argument");
    KRY_REQUIRE(k <= 6000, "k too large");
    tri_solve_kernel<true><<<1, 128, sizeof(double) * (size_t)k, ctx->stream>>>(k, Rt_dev, ldr, y_dev, out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_minres_recur(kry_ctx* ctx, int k, double* h3_dev, double* st_dev, int shift, int mailbox_off) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && h3_dev && st_dev, "bad arguments");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 8 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    minres_recur_kernel<<<1, 32, 0, ctx->stream>>>(k, h3_dev, st_dev, shift, ctx->d_mailbox + mailbox_off);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_givens_update_z(kry_ctx* ctx, int k, double* hcol_dev, double* rcol_dev, double* cs_dev, double* y_dev,
                        int mailbox_off) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && hcol_dev && rcol_dev && cs_dev && y_dev, "bad arguments");
    KRY_REQUIRE(k <= 1000, "k too large for the single-CTA Givens update (use restarts)");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 4 * k + 9 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    size_t smem = sizeof(double) * (size_t)(2 * (k + 2) + 4 * k);
    givens_z_kernel<<<1, 128, smem, ctx->stream>>>(k, hcol_dev, rcol_dev, cs_dev, y_dev, ctx->d_mailbox + mailbox_off);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_tri_solve_z(kry_ctx* ctx, int k, const double* R_dev, long long ldr, const double* y_dev, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && ldr >= k, "bad arguments");
    if (k == 0) return KRY_OK;
    KRY_REQUIRE(R_dev && y_dev && out_dev, "NULL argument");
    KRY_REQUIRE(k <= 3000, "k too large");
    tri_solve_z_kernel<<<1, 128, sizeof(double) * 2 * (size_t)k, ctx->stream>>>(k, R_dev, ldr, y_dev, out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

}  // extern "C"
