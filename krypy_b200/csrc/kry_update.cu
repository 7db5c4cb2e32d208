// Fused N-sized solver updates: CG (x, r, z, rho in one sweep) and MINRES
// (z, W shift, y in one sweep).  HBM-bound streaming kernels.
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_update_kernels.cuh"

static inline int upd_grid(const kry_ctx* ctx, long long nvec) {
    long long need = (nvec + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 4;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" {

static int cg_update_impl(kry_ctx* ctx, int dtype, long long n, const void* Ap, const void* p, void* yk, void* r,
                          void* z, const void* dinv, double rho, const double* pAp_dev, int mailbox_off, double* st) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && Ap && p && yk && r && (pAp_dev || st), "bad arguments");
    KRY_REQUIRE(!dinv || z, "dinv given without z");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 3 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    bool al = kry_aligned16(Ap) && kry_aligned16(p) && kry_aligned16(yk) && kry_aligned16(r) &&
              (!z || kry_aligned16(z)) && (!dinv || kry_aligned16(dinv));
    double* mb = ctx->d_mailbox + mailbox_off;
    if (dtype == KRY_F64) {
        if (al)
            cg_update_kernel<double, 2><<<upd_grid(ctx, n / 2), KRY_THREADS, 0, ctx->stream>>>(
                n, (const double*)Ap, (const double*)p, (double*)yk, (double*)r, (double*)z, (const double*)dinv, rho,
                pAp_dev, ctx->d_partials, ctx->d_ticket + 2, mb, st);
        else
            cg_update_kernel<double, 1><<<upd_grid(ctx, n), KRY_THREADS, 0, ctx->stream>>>(
                n, (const double*)Ap, (const double*)p, (double*)yk, (double*)r, (double*)z, (const double*)dinv, rho,
                pAp_dev, ctx->d_partials, ctx->d_ticket + 2, mb, st);
    } else if (dtype == KRY_F32) {
        if (al)
            cg_update_kernel<float, 4><<<upd_grid(ctx, n / 4), KRY_THREADS, 0, ctx->stream>>>(
                n, (const float*)Ap, (const float*)p, (float*)yk, (float*)r, (float*)z, (const float*)dinv, rho,
                pAp_dev, ctx->d_partials, ctx->d_ticket + 2, mb, st);
        else
            cg_update_kernel<float, 1><<<upd_grid(ctx, n), KRY_THREADS, 0, ctx->stream>>>(
                n, (const float*)Ap, (const float*)p, (float*)yk, (float*)r, (float*)z, (const float*)dinv, rho,
                pAp_dev, ctx->d_partials, ctx->d_ticket + 2, mb, st);
    } else {
        kry_set_error("kry_cg_update: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_cg_update(kry_ctx* ctx, int dtype, long long n, const void* Ap, const void* p, void* yk, void* r, void* z,
                  const void* dinv, double rho, const double* pAp_dev, int mailbox_off) {
    return cg_update_impl(ctx, dtype, n, Ap, p, yk, r, z, dinv, rho, pAp_dev, mailbox_off, nullptr);
}

int kry_cg_update_dev(kry_ctx* ctx, int dtype, long long n, const void* Ap, const void* p, void* yk, void* r, void* z,
                      const void* dinv, double* st_dev) {
    KRY_REQUIRE(st_dev != nullptr, "st_dev is NULL");
    return cg_update_impl(ctx, dtype, n, Ap, p, yk, r, z, dinv, 0.0, nullptr, 0, st_dev);
}

int kry_cg_scalars(kry_ctx* ctx, double* st_dev, int mailbox_off, int world, int rank, unsigned long long* epoch_dev,
                   double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(st_dev != nullptr, "st_dev is NULL");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 3 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(world == 1 || (epoch_dev && peer_slots_dev && peer_flags_dev), "NULL peer argument");
    PeerArgs pa;
    pa.world = world;
    pa.rank = rank;
    pa.epoch_dev = epoch_dev;
    pa.slots = peer_slots_dev;
    pa.flags = peer_flags_dev;
    cg_scalars_kernel<<<1, 64, 0, ctx->stream>>>(st_dev, ctx->d_mailbox + mailbox_off, pa);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_xpby_dev(kry_ctx* ctx, int dtype, long long n, const void* x, const double* beta_dev, const void* y,
                 void* out) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && x && beta_dev && y && out, "bad arguments");
    const bool al = kry_aligned16(x) && kry_aligned16(y) && kry_aligned16(out);
    if (dtype == KRY_F64) {
        if (al)
            xpby_dev_kernel<double, 2><<<upd_grid(ctx, n / 2), KRY_THREADS, 0, ctx->stream>>>(
                n, (const double*)x, beta_dev, (const double*)y, (double*)out);
        else
            xpby_dev_kernel<double, 1><<<upd_grid(ctx, n), KRY_THREADS, 0, ctx->stream>>>(
                n, (const double*)x, beta_dev, (const double*)y, (double*)out);
    } else if (dtype == KRY_F32) {
        if (al)
            xpby_dev_kernel<float, 4><<<upd_grid(ctx, n / 4), KRY_THREADS, 0, ctx->stream>>>(
                n, (const float*)x, beta_dev, (const float*)y, (float*)out);
        else
            xpby_dev_kernel<float, 1><<<upd_grid(ctx, n), KRY_THREADS, 0, ctx->stream>>>(
                n, (const float*)x, beta_dev, (const float*)y, (float*)out);
    } else {
        kry_set_error("kry_xpby_dev: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_minres_update(kry_ctx* ctx, int dtype, long long n, const void* v, void* w0, const void* w1, void* yk,
                      const double* st_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && v && w0 && w1 && yk && st_dev, "bad arguments");
    bool al = kry_aligned16(v) && kry_aligned16(w0) && kry_aligned16(w1) && kry_aligned16(yk);
    if (dtype == KRY_F64) {
        if (al)
            minres_update_kernel<double, 2><<<upd_grid(ctx, n / 2), KRY_THREADS, 0, ctx->stream>>>(
                n, (const double*)v, (double*)w0, (const double*)w1, (double*)yk, st_dev);
        else
            minres_update_kernel<double, 1><<<upd_grid(ctx, n), KRY_THREADS, 0, ctx->stream>>>(
                n, (const double*)v, (double*)w0, (const double*)w1, (double*)yk, st_dev);
    } else if (dtype == KRY_F32) {
        if (al)
            minres_update_kernel<float, 4><<<upd_grid(ctx, n / 4), KRY_THREADS, 0, ctx->stream>>>(
                n, (const float*)v, (float*)w0, (const float*)w1, (float*)yk, st_dev);
        else
            minres_update_kernel<float, 1><<<upd_grid(ctx, n), KRY_THREADS, 0, ctx->stream>>>(
                n, (const float*)v, (float*)w0, (const float*)w1, (float*)yk, st_dev);
    } else {
        kry_set_error("kry_minres_update: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

}  // extern "C"
