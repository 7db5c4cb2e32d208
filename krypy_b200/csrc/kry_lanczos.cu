// Fused Lanczos step for a DIAGONAL inner-product matrix B = diag(b) (BASELINE config C5:
// MINRES with ip_B, krypy/utils.py:1000-1045 with inner(X, Y, ip_B) = X^H (B Y), utils.py:190-193):
//
//   q -= H[k-1,k] v_{k-1}                       (three-term recurrence, utils.py:1003-1009)
//   alpha = <v_k, q>_B = sum_i v_k[i] (b[i] q[i])  ; H[k,k] += alpha ; q -= alpha v_k
//   beta  = sqrt(<q, q>_B)                        (utils.py:1034, norm(Av, ip_B))
//   v_{k+1} = q / beta                            (utils.py:1045)
//
// as ONE cooperative kernel (two grid-wide reductions) instead of the seven launches of the
// generic-inner-product path (axpy, diag_mul, block_dot, axpy, diag_mul, block_dot, scale): 11
// instead of 21 vector passes.  Default for a diagonal ip_B since round 2 (KRY_LANCZOS_DIAGB=0 in the
// host layer selects the generic sequence).  Row-partitioned runs (kry_lanczos_diag_dist) complete
// both reductions over NVLink peer memory inside the kernel.
//
// Rounding mirrors the unfused path: b*q is rounded to the storage type before it enters the dot
// (kry_diag_mul stores B q), q is used as stored.  Reductions are deterministic for a fixed grid.
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_lanczos_kernels.cuh"

template <typename T>
static int lanczos_launch(kry_ctx* ctx, LanczosArgs<T>& a) {
    const int W = VecWidth<T>::value;
    const bool al = kry_aligned16(a.vk) && kry_aligned16(a.b) && kry_aligned16(a.q) &&
                    (!a.vprev || kry_aligned16(a.vprev)) && (!a.vnext || kry_aligned16(a.vnext));
    static int blocks_per_sm[2] = {0, 0};
    const int which = al ? 0 : 1;
    const bool peer = a.peer.world > 1;
    if (blocks_per_sm[which] == 0) {
        // (the PEER instantiation needs more shared memory: size the grid for it, it fits both)
        int nb = 0;
        if (al) KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lanczos_diag_kernel<T, W, true>, KRY_THREADS, 0));
        else KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lanczos_diag_kernel<T, 1, true>, KRY_THREADS, 0));
        KRY_REQUIRE(nb >= 1, "lanczos kernel does not fit");
        blocks_per_sm[which] = nb;
    }
    long long need = ((al ? a.n / W : a.n) + KRY_THREADS - 1) / KRY_THREADS;
    if (need < 1) need = 1;
    long long cap = (long long)blocks_per_sm[which] * ctx->sm_count;
    if (cap > KRY_MAX_PARTIAL_BLOCKS) cap = KRY_MAX_PARTIAL_BLOCKS;
    const int g = (int)(need < cap ? need : cap);
    void* args[] = {&a};
    void* k = peer ? (al ? (void*)lanczos_diag_kernel<T, W, true> : (void*)lanczos_diag_kernel<T, 1, true>)
                   : (al ? (void*)lanczos_diag_kernel<T, W, false> : (void*)lanczos_diag_kernel<T, 1, false>);
    KRY_CHECK_CUDA(cudaLaunchCooperativeKernel(k, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

extern "C" {

static int lanczos_impl(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk, const void* bdiag,
                        void* q, const double* pre_coef_dev, double* h3_dev, void* vnext, PeerArgs peer) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && vk && bdiag && q && h3_dev, "bad arguments");
    KRY_REQUIRE(!vprev || pre_coef_dev, "vprev without pre_coef_dev");
    if (dtype == KRY_F64) {
        LanczosArgs<double> a = {n, (const double*)vprev, (const double*)vk, (const double*)bdiag, (double*)q,
                                 pre_coef_dev, h3_dev, (double*)vnext, ctx->d_partials, peer};
        return lanczos_launch<double>(ctx, a);
    }
    if (dtype == KRY_F32) {
        LanczosArgs<float> a = {n, (const float*)vprev, (const float*)vk, (const float*)bdiag, (float*)q,
                                pre_coef_dev, h3_dev, (float*)vnext, ctx->d_partials, peer};
        return lanczos_launch<float>(ctx, a);
    }
    kry_set_error("kry_lanczos_diag: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_lanczos_diag(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk, const void* bdiag,
                     void* q, const double* pre_coef_dev, double* h3_dev, void* vnext) {
    PeerArgs peer;
    memset(&peer, 0, sizeof(peer));
    peer.world = 1;
    return lanczos_impl(ctx, dtype, n, vprev, vk, bdiag, q, pre_coef_dev, h3_dev, vnext, peer);
}

int kry_lanczos_diag_dist(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk, const void* bdiag,
                          void* q, const double* pre_coef_dev, double* h3_dev, void* vnext, int world, int rank,
                          unsigned long long* epoch_dev, double* const* peer_slots_dev,
                          unsigned long long* const* peer_flags_dev) {
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(epoch_dev && peer_slots_dev && peer_flags_dev, "NULL peer argument");
    PeerArgs peer;
    peer.world = world;
    peer.rank = rank;
    peer.epoch_dev = epoch_dev;
    peer.slots = peer_slots_dev;
    peer.flags = peer_flags_dev;
    return lanczos_impl(ctx, dtype, n, vprev, vk, bdiag, q, pre_coef_dev, h3_dev, vnext, peer);
}

}  // extern "C"
