// CSR SpMV for sm_100a.
//
// Short-row matrices (stencils: 5/7 nnz per row) use a persistent, warp-specialised
// kernel: one producer warp stages each 256-row block's contiguous vals[] and colidx[]
// ranges into shared memory with 1-D TMA bulk copies (cp.async.bulk.shared::cluster.global
// + mbarrier complete_tx, SASS UBLKCP) through a STAGES-deep ring (full/empty mbarriers),
// so tens of KB per SM are in flight without holding registers.  Eight consumer warps run
// thread-per-row out of shared memory: for banded matrices the x gathers of neighbouring
// threads are contiguous (coalesced ld.global.nc), the next tile's row pointers are
// prefetched into registers while the current tile is computed, and each row is summed
// sequentially left to right with separately rounded multiply and add -- the same order and
// rounding as scipy's csr_matvec, so y is bit-identical to the reference's SpMV in fp64.
//
// Long-row matrices (or unaligned arrays) use a warp-per-row kernel with coalesced direct
// loads and a shuffle reduction.
#include <stdlib.h>
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_spmv_kernels.cuh"

static int spmv_stage_override() {   // tuning knob: KRY_SPMV_STAGES=2|3|4
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("KRY_SPMV_STAGES");
        v = s ? atoi(s) : 0;
        if (v != 2 && v != 3 && v != 4) v = 0;
    }
    return v;
}

template <typename T, int CPR, int STAGES, bool DOT, int NACC>
static int launch_staged_md(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                            const T* vals, const T* x, T* y, const T* w, double* dot_out, const MDotArgs<T>& md) {
    typedef SpmvCfg<T, CPR, STAGES> Cfg;
    auto kern = spmv_staged_kernel<T, CPR, STAGES, DOT, NACC>;
    static thread_local int occ[16] = {0};
    int& o = occ[ctx->device & 15];
    if (o == 0) {
        KRY_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        int nb = 0;
        KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, SPMV_THREADS, Cfg::SMEM_BYTES));
        KRY_REQUIRE(nb >= 1, "staged SpMV kernel does not fit on an SM");
        o = nb;
    }
    long long ntiles = (nrows + SPMV_R - 1) / SPMV_R;
    long long cap = (long long)ctx->sm_count * o;
    if (cap > KRY_MAX_PARTIAL_BLOCKS) cap = KRY_MAX_PARTIAL_BLOCKS;
    int g = (int)(ntiles < cap ? ntiles : cap);
    if (g < 1) g = 1;
    kern<<<g, SPMV_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(nrows, nnz, rowptr, colidx, vals, x, y, w,
                                                            ctx->d_partials, ctx->d_ticket + (NACC > 0 ? 13 : 1),
                                                            dot_out, md);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T, int CPR, int STAGES, bool DOT>
static int launch_staged(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                         const T* vals, const T* x, T* y, const T* w, double* dot_out) {
    MDotArgs<T> md;
    memset(&md, 0, sizeof(md));
    return launch_staged_md<T, CPR, STAGES, DOT, 0>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out, md);
}

// multi-dot epilogue: accumulator tile (8 / 16 / 32 registers) by the number of vectors; deeper TMA ring
// where the register tile leaves a single CTA per SM
template <typename T, int CPR>
static int launch_mdot_cpr(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                           const T* vals, const T* x, T* y, const MDotArgs<T>& md) {
    if (md.nb <= 8)
        return launch_staged_md<T, CPR, 2, false, 8>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, nullptr, md);
    if (md.nb <= 16)
        return launch_staged_md<T, CPR, 3, false, 16>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, nullptr, md);
    return launch_staged_md<T, CPR, 4, false, 32>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, nullptr, md);
}

template <typename T>
static int spmv_mdot_dispatch(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                              const T* vals, const T* x, T* y, const MDotArgs<T>& md) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    const bool al = kry_aligned16(vals) && kry_aligned16(colidx);
    if (al && avg <= 5.5) return launch_mdot_cpr<T, 6>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, md);
    if (al && avg <= 7.5) return launch_mdot_cpr<T, 8>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, md);
    if (al && avg <= 15.0) return launch_mdot_cpr<T, 16>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, md);
    kry_set_error("kry_spmv_csr_mdot: only the staged short-row path (<= 15 entries per row on average, 16-byte "
                  "aligned matrix arrays) carries the multi-dot epilogue");
    return KRY_ERR_UNSUPPORTED;
}

template <typename T, int CPR, bool DOT>
static int launch_staged_cpr(kry_ctx* ctx, int stages, long long nrows, long long nnz, const int* rowptr,
                             const int* colidx, const T* vals, const T* x, T* y, const T* w, double* dot_out) {
    if (stages == 2) return launch_staged<T, CPR, 2, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (stages == 4) return launch_staged<T, CPR, 4, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    return launch_staged<T, CPR, 3, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
}

template <typename T, bool DOT>
static int spmv_dispatch(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                         const T* vals, const T* x, T* y, const T* w, double* dot_out) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    const bool al = kry_aligned16(vals) && kry_aligned16(colidx);
    const int ov = spmv_stage_override();
    // stage capacity (entries per row) just above the mean row length; a 2-deep ring measured
    // fastest on B200 (118 us vs 121/133 us for 3/4 stages on config C2: occupancy beats depth)
    if (al && avg <= 5.5)
        return launch_staged_cpr<T, 6, DOT>(ctx, ov ? ov : 2, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (al && avg <= 7.5)
        return launch_staged_cpr<T, 8, DOT>(ctx, ov ? ov : 2, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (al && avg <= 15.0)
        return launch_staged_cpr<T, 16, DOT>(ctx, ov ? ov : 2, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (al && avg <= 30.0)
        return launch_staged<T, 32, 2, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    long long need = (nrows * 32 + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 8;
    int g = (int)(need < cap ? need : cap);
    if (g < 1) g = 1;
    spmv_warp_kernel<T, DOT><<<g, KRY_THREADS, 0, ctx->stream>>>(nrows, rowptr, colidx, vals, x, y, w,
                                                                 ctx->d_partials, ctx->d_ticket + 1, dot_out);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

extern "C" int kry_spmv_csr(kry_ctx* ctx, int dtype, long long nrows, long long ncols, long long nnz,
                            const int* rowptr, const int* colidx, const void* vals, const void* x, void* y,
                            const void* w_dev, double* dot_out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, "negative size");
    KRY_REQUIRE(nnz < 2147483647LL, "nnz must fit int32 row pointers");
    KRY_REQUIRE(rowptr && x, "NULL argument");
    KRY_REQUIRE(nnz == 0 || (colidx && vals), "NULL matrix arrays");
    KRY_REQUIRE(y || w_dev, "neither y nor a dot epilogue requested");
    KRY_REQUIRE(!w_dev || dot_out_dev, "w given without dot_out");
    if (nrows == 0) return KRY_OK;
    if (dtype == KRY_F64) {
        if (w_dev)
            return spmv_dispatch<double, true>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals,
                                               (const double*)x, (double*)y, (const double*)w_dev, dot_out_dev);
        return spmv_dispatch<double, false>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals, (const double*)x,
                                            (double*)y, nullptr, nullptr);
    }
    if (dtype == KRY_F32) {
        if (w_dev)
            return spmv_dispatch<float, true>(ctx, nrows, nnz, rowptr, colidx, (const float*)vals, (const float*)x,
                                              (float*)y, (const float*)w_dev, dot_out_dev);
        return spmv_dispatch<float, false>(ctx, nrows, nnz, rowptr, colidx, (const float*)vals, (const float*)x,
                                           (float*)y, nullptr, nullptr);
    }
    kry_set_error("kry_spmv_csr: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

// y = A x and c[j] = <B[j], y> (j < nb), c[nb] = <y, y> (want_sq) in ONE pass: the dots are taken while each
// row's result is still in the register of the thread that computed it.
//   world == 1: the sums go to out_dev[0 .. nb + want_sq)
//   world  > 1: row-partitioned run; the local sums are stored into every peer's slot array and this
//               rank's flag is released with epoch + 1 (no wait here: the consumer, kry_dist_update_scale
//               or kry_dist_update, acquires)
// Replaces utils.py:968 (A v_k) + the k+1 inner products of utils.py:1015 (block classical Gram-Schmidt),
// and deflation.py:135-143 / utils.py:604-627 (<W, A v> of the projector) when B = W.
extern "C" int kry_spmv_csr_mdot(kry_ctx* ctx, int dtype, long long nrows, long long ncols, long long nnz,
                                 const int* rowptr, const int* colidx, const void* vals, const void* x, void* y,
                                 const void* B, long long ldb, int nb, int want_sq, double* out_dev, int world,
                                 int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                                 unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nrows >= 1 && ncols >= 0 && nnz >= 0, "bad size");
    KRY_REQUIRE(nnz < 2147483647LL, "nnz must fit int32 row pointers");
    KRY_REQUIRE(rowptr && x && y, "NULL argument");
    KRY_REQUIRE(nnz == 0 || (colidx && vals), "NULL matrix arrays");
    KRY_REQUIRE(nb >= 0 && nb <= 32 && nb + (want_sq ? 1 : 0) >= 1, "0 <= nb <= 32 dot vectors (+ the square) per call");
    KRY_REQUIRE(nb == 0 || B, "NULL dot basis");
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(world > 1 ? (epoch_dev && peer_slots_dev && peer_flags_dev) : (out_dev != nullptr),
                "world > 1 needs the peer tables, world == 1 an output array");
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = world;
    pa.rank = rank;
    pa.epoch_dev = epoch_dev;
    pa.slots = peer_slots_dev;
    pa.flags = peer_flags_dev;
    if (dtype == KRY_F64) {
        MDotArgs<double> md = {(const double*)B, ldb, nb, want_sq ? 1 : 0, out_dev, pa};
        return spmv_mdot_dispatch<double>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals, (const double*)x,
                                          (double*)y, md);
    }
    if (dtype == KRY_F32) {
        MDotArgs<float> md = {(const float*)B, ldb, nb, want_sq ? 1 : 0, out_dev, pa};
        return spmv_mdot_dispatch<float>(ctx, nrows, nnz, rowptr, colidx, (const float*)vals, (const float*)x,
                                         (float*)y, md);
    }
    kry_set_error("kry_spmv_csr_mdot: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}
