// Fused Gram-Schmidt step and fused oblique (deflation) projection: ONE
// cooperative persistent kernel each.  HBM-bound tall-skinny sweeps:
//   phase A  c = Vdot^H q      (register tile of JT basis vectors per pass over q,
//                               warp-shuffle + shared-memory CTA reduction,
//                               per-CTA partials, grid.sync, fixed-order final sum
//                               recomputed identically by every CTA)
//   phase B  q -= Vsub c       (+ ||q||^2 partials in the epilogue)
//   phase C  vnext = q / ||q||
// Every thread owns the same q elements in every phase (identical grid-stride
// map), so the only grid-wide dependencies are the reductions themselves.
// Reductions are deterministic for a fixed grid size.
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_sweeps.cuh"
#include "kry_orth_kernels.cuh"

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <typename K>
static int max_blocks_of(K kern, int* out) {
    int nb = 0;
    KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, KRY_THREADS, 0));
    *out = nb;
    return KRY_OK;
}

int kry_orth_max_blocks(int dtype, int* out) {
    int a = 0, b = 0, rc;
    if (dtype == KRY_F64) {
        if ((rc = max_blocks_of(orth_kernel<double, 2, true>, &a))) return rc;
        if ((rc = max_blocks_of(orth_kernel<double, 1, true>, &b))) return rc;
    } else {
        if ((rc = max_blocks_of(orth_kernel<float, 4, true>, &a))) return rc;
        if ((rc = max_blocks_of(orth_kernel<float, 1, true>, &b))) return rc;
    }
    *out = a < b ? a : b;
    KRY_REQUIRE(*out >= 1, "orth kernel does not fit");
    return KRY_OK;
}

int kry_proj_max_blocks(int dtype, int* out) {
    int a = 0, b = 0, rc;
    if (dtype == KRY_F64) {
        if ((rc = max_blocks_of(proj_kernel<double, 2>, &a))) return rc;
        if ((rc = max_blocks_of(proj_kernel<double, 1>, &b))) return rc;
    } else {
        if ((rc = max_blocks_of(proj_kernel<float, 4>, &a))) return rc;
        if ((rc = max_blocks_of(proj_kernel<float, 1>, &b))) return rc;
    }
    *out = a < b ? a : b;
    KRY_REQUIRE(*out >= 1, "projection kernel does not fit");
    return KRY_OK;
}

static int coop_grid(long long work_items, int max_blocks) {
    long long need = (work_items + KRY_THREADS - 1) / KRY_THREADS;
    if (need < 1) need = 1;
    long long cap = max_blocks < KRY_MAX_PARTIAL_BLOCKS ? max_blocks : KRY_MAX_PARTIAL_BLOCKS;
    return (int)(need < cap ? need : cap);
}

template <typename T>
static int orth_launch(kry_ctx* ctx, OrthArgs<T>& a, int max_blocks) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(a.Vdot) && kry_aligned16(a.Vsub) && kry_aligned16(a.q) && (a.ldv % W == 0) &&
              (!a.pre_vec || kry_aligned16(a.pre_vec)) && (!a.vnext || kry_aligned16(a.vnext));
    void* args[] = {&a};
    const bool peer = a.peer.world > 1;
    if (al) {
        int g = coop_grid(a.n / W, max_blocks);
        void* k = peer ? (void*)orth_kernel<T, W, true> : (void*)orth_kernel<T, W, false>;
        KRY_CHECK_CUDA(cudaLaunchCooperativeKernel(k, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
    } else {
        int g = coop_grid(a.n, max_blocks);
        void* k = peer ? (void*)orth_kernel<T, 1, true> : (void*)orth_kernel<T, 1, false>;
        KRY_CHECK_CUDA(cudaLaunchCooperativeKernel(k, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T>
static int proj_launch(kry_ctx* ctx, ProjArgs<T>& p, int max_blocks) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(p.W) && kry_aligned16(p.V) && kry_aligned16(p.a) && (p.ldw % W == 0) && (p.ldv % W == 0);
    void* args[] = {&p};
    if (al) {
        int g = coop_grid(p.n / W, max_blocks);
        KRY_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)proj_kernel<T, W>, dim3(g), dim3(KRY_THREADS), args, 0,
                                                   ctx->stream));
    } else {
        int g = coop_grid(p.n, max_blocks);
        KRY_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)proj_kernel<T, 1>, dim3(g), dim3(KRY_THREADS), args, 0,
                                                   ctx->stream));
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

extern "C" {

static int orth_fused_impl(kry_ctx* ctx, int dtype, long long n, const void* Vdot, const void* Vsub, long long ldv,
                           int j0, int nv, void* q, int passes, int algo, const void* pre_vec,
                           const double* pre_coef_dev, double* h_dev, double* nrm_dev, void* vnext, PeerArgs peer) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && q, "bad arguments");
    KRY_REQUIRE(j0 >= 0 && nv >= j0, "bad basis range");
    KRY_REQUIRE(nv == j0 || (Vdot && Vsub && h_dev), "NULL basis / h");
    KRY_REQUIRE(passes == 1 || passes == 2, "passes must be 1 or 2");
    KRY_REQUIRE(algo == KRY_ORTH_CGS || algo == KRY_ORTH_MGS, "unknown algo");
    KRY_REQUIRE(!pre_vec || pre_coef_dev, "pre_vec without pre_coef_dev");
    KRY_REQUIRE(!vnext || nrm_dev, "vnext requires nrm_dev");
    KRY_REQUIRE(algo != KRY_ORTH_CGS || nv - j0 <= KRY_MAX_SLOTS, "CGS: too many vectors in one call");
    if (!Vdot) Vdot = q;   // never dereferenced when nv == j0; keeps alignment checks simple
    if (!Vsub) Vsub = q;
    if (dtype == KRY_F64) {
        OrthArgs<double> a = {n, (const double*)Vdot, (const double*)Vsub, ldv, j0, nv, passes, algo, (double*)q,
                              (const double*)pre_vec, pre_coef_dev, h_dev, nrm_dev, (double*)vnext, ctx->d_partials,
                              peer};
        return orth_launch<double>(ctx, a, ctx->orth_blocks_f64);
    }
    if (dtype == KRY_F32) {
        OrthArgs<float> a = {n, (const float*)Vdot, (const float*)Vsub, ldv, j0, nv, passes, algo, (float*)q,
                             (const float*)pre_vec, pre_coef_dev, h_dev, nrm_dev, (float*)vnext, ctx->d_partials,
                             peer};
        return orth_launch<float>(ctx, a, ctx->orth_blocks_f32);
    }
    kry_set_error("kry_orth_fused: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_orth_fused(kry_ctx* ctx, int dtype, long long n, const void* Vdot, const void* Vsub, long long ldv, int j0,
                   int nv, void* q, int passes, int algo, const void* pre_vec, const double* pre_coef_dev,
                   double* h_dev, double* nrm_dev, void* vnext) {
    PeerArgs peer;
    memset(&peer, 0, sizeof(peer));
    peer.world = 1;
    return orth_fused_impl(ctx, dtype, n, Vdot, Vsub, ldv, j0, nv, q, passes, algo, pre_vec, pre_coef_dev, h_dev,
                           nrm_dev, vnext, peer);
}

int kry_orth_fused_dist(kry_ctx* ctx, int dtype, long long n, const void* Vdot, const void* Vsub, long long ldv,
                        int j0, int nv, void* q, int passes, int algo, const void* pre_vec,
                        const double* pre_coef_dev, double* h_dev, double* nrm_dev, void* vnext, int world, int rank,
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
    return orth_fused_impl(ctx, dtype, n, Vdot, Vsub, ldv, j0, nv, q, passes, algo, pre_vec, pre_coef_dev, h_dev,
                           nrm_dev, vnext, peer);
}

int kry_project(kry_ctx* ctx, int dtype, long long n, const void* W, long long ldw, const void* V, long long ldv,
                int d, void* a, const double* Q_dev, const double* R_dev, int iterations, double* c_first_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && a && d >= 0 && iterations >= 1, "bad arguments");
    if (d == 0) return KRY_OK;
    KRY_REQUIRE(d <= KRY_MAX_SLOTS, "too many deflation vectors for one call");
    KRY_REQUIRE(W && V, "NULL basis");
    KRY_REQUIRE((Q_dev == nullptr) == (R_dev == nullptr), "Q and R must be given together");
    if (dtype == KRY_F64) {
        ProjArgs<double> p = {n, (const double*)W, ldw, (const double*)V, ldv, d, iterations, (double*)a,
                              Q_dev, R_dev, c_first_dev, ctx->d_partials};
        return proj_launch<double>(ctx, p, ctx->proj_blocks_f64);
    }
    if (dtype == KRY_F32) {
        ProjArgs<float> p = {n, (const float*)W, ldw, (const float*)V, ldv, d, iterations, (float*)a,
                             Q_dev, R_dev, c_first_dev, ctx->d_partials};
        return proj_launch<float>(ctx, p, ctx->proj_blocks_f32);
    }
    kry_set_error("kry_project: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

}  // extern "C"
