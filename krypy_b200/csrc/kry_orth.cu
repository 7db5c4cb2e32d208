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

template <typename T>
struct OrthArgs {
    long long n;
    const T* Vdot;
    const T* Vsub;
    long long ldv;
    int j0, nv, passes, algo;
    T* q;
    const T* pre_vec;
    const double* pre_coef;
    double* h;
    double* nrm;
    T* vnext;
    double* partials;   // [2][KRY_MAX_SLOTS][KRY_MAX_PARTIAL_BLOCKS]
    PeerArgs peer;      // world == 1: single GPU; otherwise the reductions are completed over NVLink
};

template <typename T, int VEC, bool PEER>
__global__ void __launch_bounds__(KRY_THREADS, 2) orth_kernel(OrthArgs<T> a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[32];
    __shared__ double red[ORTH_JT * 8];
    __shared__ double c_s[KRY_MAX_SLOTS];
    __shared__ double stage[PEER ? PEER_MAX_RANKS * PEER_SLOT : 1];
    __shared__ int okflag;
    unsigned long long epoch = PEER ? dld_volatile_u64(a.peer.epoch_dev) : 0ull;
    const long long n = a.n, ldv = a.ldv;
    T* q = a.q;
    int buf = 0;
    const int cnt = a.nv - a.j0;
    bool pre_pending = (a.pre_vec != nullptr);
    const double pre_c = pre_pending ? a.pre_coef[0] : 0.0;
    double nrm2_part = 0.0;

    if (a.algo == KRY_ORTH_CGS) {
        if (pre_pending) {
            // (Lanczos-style pre-subtraction with the block algorithm: not used by the solvers, kept
            // for the ABI) q -= pre_c * pre_vec as a sweep of its own
            mgs_pass<T, VEC>(nullptr, nullptr, 0.0, a.pre_vec, pre_c, q, n, false);
            pre_pending = false;
        }
        for (int pass = 0; pass < a.passes; ++pass) {
            // ---- phase A: block dots, up to 16 vectors per pass over q ----
            for (int jb = 0; jb < cnt; jb += ORTH_JT) {
                const int nt = cnt - jb < ORTH_JT ? cnt - jb : ORTH_JT;
                dots_dispatch<T, VEC>(nt, a.Vdot + (long long)(a.j0 + jb) * ldv, ldv, q, n, red, a.partials, buf, jb);
            }
            grid.sync();
            reduce_slots(a.partials, buf, cnt, c_s);
            if (PEER && cnt > 0) peer_exchange(a.peer, ++epoch, c_s, cnt, stage, &okflag);
            if (blockIdx.x == 0)
                for (int s = threadIdx.x; s < cnt; s += blockDim.x) a.h[a.j0 + s] += c_s[s];
            buf ^= 1;
            // ---- phase B: q -= Vsub c (+ ||q||^2 in the last pass) ----
            const bool want_nrm = (a.nrm != nullptr) && (pass == a.passes - 1);
            if (cnt > 0 || want_nrm)
                nrm2_part = update_dispatch<T, VEC, false>(a.Vsub + (long long)a.j0 * ldv, ldv, cnt, c_s, q, n, want_nrm);
            __syncthreads();   // c_s is rewritten by the next pass
        }
    } else {
        // ---- exact modified Gram-Schmidt: one dependent reduction per basis vector; the update with
        //      vector j-1 is fused into the sweep that computes <v_j, q> ----
        double c_prev = 0.0;
        int j_prev = -1;
        for (int pass = 0; pass < a.passes; ++pass) {
            for (int j = a.j0; j < a.nv; ++j) {
                const T* vj = a.Vdot + (long long)j * ldv;
                const T* vp = j_prev >= 0 ? a.Vsub + (long long)j_prev * ldv : nullptr;
                double acc = mgs_pass<T, VEC>(vj, vp, c_prev, pre_pending ? a.pre_vec : nullptr, pre_c, q, n, false);
                pre_pending = false;
                double s = kry_block_sum(acc, sm);
                if (threadIdx.x == 0) partial_slot(a.partials, buf, 0)[blockIdx.x] = s;
                grid.sync();
                c_prev = reduce_slot(a.partials, buf, 0, sm);
                if (PEER) {
                    __syncthreads();
                    if (threadIdx.x == 0) c_s[0] = c_prev;
                    __syncthreads();
                    peer_exchange(a.peer, ++epoch, c_s, 1, stage, &okflag);
                    c_prev = c_s[0];
                    __syncthreads();
                }
                j_prev = j;
                if (blockIdx.x == 0 && threadIdx.x == 0) a.h[j] += c_prev;
                buf ^= 1;
            }
        }
        // flush the pending subtraction (and a lone pre-subtraction when nv == j0)
        const T* vp = j_prev >= 0 ? a.Vsub + (long long)j_prev * ldv : nullptr;
        const bool want_nrm = (a.nrm != nullptr);
        if (vp || pre_pending || want_nrm)
            nrm2_part = mgs_pass<T, VEC>(nullptr, vp, c_prev, pre_pending ? a.pre_vec : nullptr, pre_c, q, n, want_nrm);
    }

    // ---- norm and phase C ----
    if (a.nrm != nullptr) {
        double s = kry_block_sum(nrm2_part, sm);
        if (threadIdx.x == 0) partial_slot(a.partials, buf, 0)[blockIdx.x] = s;
        grid.sync();
        double nrm2 = reduce_slot(a.partials, buf, 0, sm);
        if (PEER) {
            __syncthreads();
            if (threadIdx.x == 0) c_s[0] = nrm2;
            __syncthreads();
            peer_exchange(a.peer, ++epoch, c_s, 1, stage, &okflag);
            nrm2 = c_s[0];
        }
        const double nrm = sqrt(nrm2);
        if (blockIdx.x == 0 && threadIdx.x == 0) a.nrm[0] = nrm;
        if (a.vnext != nullptr) scale_pass<T, VEC>(q, a.vnext, n, nrm);
    }
    if (PEER) {
        // every CTA read epoch_dev before the first grid.sync; one more grid-wide sync orders the
        // write-back after all of those reads (also when no reduction was needed)
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.peer.epoch_dev = epoch;
    }
}

// ---------------------------------------------------------------------------
// oblique projection  a <- (I - V R^-1 Q^H W^H)^iterations a
// ---------------------------------------------------------------------------
template <typename T>
struct ProjArgs {
    long long n;
    const T* W;
    long long ldw;
    const T* V;
    long long ldv;
    int d, iterations;
    T* a;
    const double* Q;
    const double* R;
    double* c_first;
    double* partials;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2) proj_kernel(ProjArgs<T> p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[ORTH_JT * 8];
    __shared__ double c_s[KRY_MAX_SLOTS];
    __shared__ double t_s[KRY_MAX_SLOTS];
    const long long n = p.n;
    const int d = p.d;
    T* av = p.a;
    int buf = 0;
    for (int iter = 0; iter < p.iterations; ++iter) {
        for (int jb = 0; jb < d; jb += ORTH_JT) {
            const int nt = d - jb < ORTH_JT ? d - jb : ORTH_JT;
            dots_dispatch<T, VEC>(nt, p.W + (long long)jb * p.ldw, p.ldw, av, n, red, p.partials, buf, jb);
        }
        grid.sync();
        reduce_slots(p.partials, buf, d, c_s);
        buf ^= 1;
        if (iter == 0 && p.c_first && blockIdx.x == 0)
            for (int s = threadIdx.x; s < d; s += blockDim.x) p.c_first[s] = c_s[s];
        // x = R^{-1} Q^H c   (utils.py:547-548), every CTA redundantly and identically
        if (p.Q != nullptr) {
            for (int i = threadIdx.x; i < d; i += blockDim.x) {
                double t = 0.0;
                for (int j = 0; j < d; ++j) t = fma(__ldg(p.Q + (long long)j * d + i), c_s[j], t);
                t_s[i] = t;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int j = d - 1; j >= 0; --j) {   // column-oriented back substitution
                    const double xj = t_s[j] / __ldg(p.R + (long long)j * d + j);
                    t_s[j] = xj;
                    for (int i = 0; i < j; ++i) t_s[i] = fma(-xj, __ldg(p.R + (long long)i * d + j), t_s[i]);
                }
            }
            __syncthreads();
        } else {
            for (int i = threadIdx.x; i < d; i += blockDim.x) t_s[i] = c_s[i];
            __syncthreads();
        }
        // a -= V x  (the reference forms Pa = V.dot(x) first and then subtracts, utils.py:549, 621)
        update_dispatch<T, VEC, true>(p.V, p.ldv, d, t_s, av, n, false);
        __syncthreads();
    }
}

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
