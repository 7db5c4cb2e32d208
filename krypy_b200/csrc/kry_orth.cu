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
#include <stdlib.h>

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#define ORTH_JT 16

// value as it reads back after being stored as T (identity for double)
template <typename T> __device__ __forceinline__ double round_as(double v) { return (double)(T)v; }

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

__device__ __forceinline__ double* partial_slot(double* partials, int buf, int slot) {
    return partials + ((size_t)buf * KRY_MAX_SLOTS + (size_t)slot) * KRY_MAX_PARTIAL_BLOCKS;
}

// fixed-order sum of one slot's per-CTA partials; identical in every CTA
__device__ __forceinline__ double reduce_slot(double* partials, int buf, int slot, double* sm) {
    const double* p = partial_slot(partials, buf, slot);
    double v = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) v += __ldcg(p + b);
    return kry_block_sum(v, sm);
}

// c[slot] for slots [0, cnt): warps split the slots, lanes stride over CTAs
__device__ __forceinline__ void reduce_slots(double* partials, int buf, int cnt, double* c_s) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int s = w; s < cnt; s += nw) {
        const double* p = partial_slot(partials, buf, s);
        double v = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(p + b);
        v = kry_warp_sum(v);
        if (lane == 0) c_s[s] = v;
    }
    __syncthreads();
}

// Row-partitioned run: turn the local sums c_s[0..cnt) (identical in every CTA) into global sums.
// CTA 0 stores them into every peer's slot array and releases its flag; every CTA acquires all
// flags and sums the per-rank partials in rank order (bitwise identical on all ranks).
__device__ __forceinline__ void peer_exchange(const PeerArgs& pa, unsigned long long epoch, double* c_s, int cnt,
                                              double* stage, int* okflag) {
    if (blockIdx.x == 0) peer_publish(pa, epoch, c_s, cnt);
    const bool ok = peer_wait(pa, epoch, okflag);
    const double* mine = pa.slots[pa.rank] + (size_t)(epoch & 1ull) * (size_t)pa.world * PEER_SLOT;
    for (int idx = threadIdx.x; idx < pa.world * cnt; idx += blockDim.x) {
        const int r = idx / cnt, j = idx - r * cnt;
        stage[r * PEER_SLOT + j] = dld_volatile_f64(mine + (size_t)r * PEER_SLOT + j);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        double sum = 0.0;
        for (int r = 0; r < pa.world; ++r) sum += stage[r * PEER_SLOT + j];
        c_s[j] = ok ? sum : nan_f64();
    }
    __syncthreads();
}

// JT: basis vectors per register tile.  JT = 16 (128 registers, 2 CTAs/SM) is sized for the long
// sweeps of GMRES(30); the JT = 4 instantiation (<= 64 registers, 4 CTAs/SM, phase C unrolled by
// hand) keeps more loads in flight when only a few vectors are involved -- small k, Lanczos, exact
// MGS.  It is an opt-in measurement variant (KRY_ORTH_SMALLK=1): the JT = 16 code is unchanged.
// CU: phase C (the normalised store) unrolled by hand, four loads in flight per thread.  Always on
// for JT = 4; KRY_ORTH_CUNROLL=1 selects it for the JT = 16 kernel as a second measurement variant.
template <typename T, int VEC, bool PEER, int JT = 16, bool CU = (JT < 16)>
__global__ void __launch_bounds__(KRY_THREADS, (JT >= 16 ? 2 : 4)) orth_kernel(OrthArgs<T> a) {
    constexpr int TB = JT < 8 ? JT : 8;     // vectors loaded per inner tile
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[32];
    __shared__ double c_s[KRY_MAX_SLOTS];
    __shared__ double stage[PEER ? PEER_MAX_RANKS * PEER_SLOT : 1];
    __shared__ int okflag;
    unsigned long long epoch = PEER ? dld_volatile_u64(a.peer.epoch_dev) : 0ull;
    const long long n = a.n, ldv = a.ldv;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tail0 = nvec * VEC + threadIdx.x;   // scalar tail handled by CTA 0
    const bool tail_cta = (blockIdx.x == 0);
    T* q = a.q;
    int buf = 0;
    const int cnt = a.nv - a.j0;
    bool pre_pending = (a.pre_vec != nullptr);
    const double pre_c = pre_pending ? a.pre_coef[0] : 0.0;
    double nrm2_part = 0.0;

    if (a.algo == KRY_ORTH_CGS) {
        for (int pass = 0; pass < a.passes; ++pass) {
            // ---- phase A: block dots ----
            for (int jb = 0; jb < cnt || (jb == 0 && pre_pending); jb += JT) {
                double acc[JT];
#pragma unroll
                for (int t = 0; t < JT; ++t) acc[t] = 0.0;
                for (long long i = i0; i < nvec; i += stride) {
                    double qv[VEC];
                    VecIO<T, VEC>::loadrw(q, i, qv);
                    if (pre_pending) {
                        double pv[VEC];
                        VecIO<T, VEC>::load(a.pre_vec, i, pv);
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[u] = fma(-pre_c, pv[u], qv[u]);
                        VecIO<T, VEC>::store(q, i, qv);
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[u] = round_as<T>(qv[u]);   // value as stored
                    }
                    if (cnt > 0) {
#pragma unroll
                        for (int tb = 0; tb < JT; tb += TB) {
                            if (jb + tb < cnt) {
                                double vv[TB][VEC];
#pragma unroll
                                for (int t = 0; t < TB; ++t) {
                                    int j = jb + tb + t;
                                    j = j < cnt ? j : cnt - 1;
                                    VecIO<T, VEC>::load(a.Vdot + (long long)(a.j0 + j) * ldv, i, vv[t]);
                                }
#pragma unroll
                                for (int t = 0; t < TB; ++t)
#pragma unroll
                                    for (int u = 0; u < VEC; ++u) acc[tb + t] = fma(vv[t][u], qv[u], acc[tb + t]);
                            }
                        }
                    }
                }
                if (tail_cta) {
                    for (long long i = tail0; i < n; i += blockDim.x) {
                        double qe = (double)q[i];
                        if (pre_pending) {
                            qe = fma(-pre_c, (double)a.pre_vec[i], qe);
                            q[i] = (T)qe;
                            qe = (double)q[i];
                        }
#pragma unroll
                        for (int t = 0; t < JT; ++t)
                            if (jb + t < cnt)
                                acc[t] = fma((double)a.Vdot[(long long)(a.j0 + jb + t) * ldv + i], qe, acc[t]);
                    }
                }
                pre_pending = false;
#pragma unroll
                for (int t = 0; t < JT; ++t) {
                    if (jb + t < cnt) {   // uniform across the CTA
                        double s = kry_block_sum(acc[t], sm);
                        if (threadIdx.x == 0) partial_slot(a.partials, buf, jb + t)[blockIdx.x] = s;
                    }
                }
            }
            grid.sync();
            reduce_slots(a.partials, buf, cnt, c_s);
            if (PEER && cnt > 0) peer_exchange(a.peer, ++epoch, c_s, cnt, stage, &okflag);
            if (blockIdx.x == 0)
                for (int s = threadIdx.x; s < cnt; s += blockDim.x) a.h[a.j0 + s] += c_s[s];
            buf ^= 1;
            // ---- phase B: q -= Vsub c ----
            const bool want_nrm = (a.nrm != nullptr) && (pass == a.passes - 1);
            for (long long i = i0; i < nvec; i += stride) {
                double qv[VEC];
                VecIO<T, VEC>::loadrw(q, i, qv);
                for (int jb = 0; jb < cnt; jb += TB) {
                    double vv[TB][VEC];
#pragma unroll
                    for (int t = 0; t < TB; ++t) {
                        int j = jb + t < cnt ? jb + t : cnt - 1;
                        VecIO<T, VEC>::load(a.Vsub + (long long)(a.j0 + j) * ldv, i, vv[t]);
                    }
#pragma unroll
                    for (int t = 0; t < TB; ++t)
                        if (jb + t < cnt) {
                            const double c = c_s[jb + t];
#pragma unroll
                            for (int u = 0; u < VEC; ++u) qv[u] = fma(-c, vv[t][u], qv[u]);
                        }
                }
                VecIO<T, VEC>::store(q, i, qv);
                if (want_nrm) {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) {
                        const double r = round_as<T>(qv[u]);
                        nrm2_part = fma(r, r, nrm2_part);
                    }
                }
            }
            if (tail_cta) {
                for (long long i = tail0; i < n; i += blockDim.x) {
                    double qe = (double)q[i];
                    for (int j = 0; j < cnt; ++j)
                        qe = fma(-c_s[j], (double)a.Vsub[(long long)(a.j0 + j) * ldv + i], qe);
                    q[i] = (T)qe;
                    qe = (double)q[i];
                    if (want_nrm) nrm2_part = fma(qe, qe, nrm2_part);
                }
            }
            __syncthreads();   // c_s is rewritten by the next pass
        }
    } else {
        // ---- exact modified Gram-Schmidt: one dependent reduction per basis vector ----
        double c_prev = 0.0;
        int j_prev = -1;
        for (int pass = 0; pass < a.passes; ++pass) {
            for (int j = a.j0; j < a.nv; ++j) {
                double acc = 0.0;
                const T* vj = a.Vdot + (long long)j * ldv;
                const T* vp = j_prev >= 0 ? a.Vsub + (long long)j_prev * ldv : nullptr;
                const bool modify = (j_prev >= 0) || pre_pending;
                for (long long i = i0; i < nvec; i += stride) {
                    double qv[VEC], vv[VEC];
                    VecIO<T, VEC>::loadrw(q, i, qv);
                    VecIO<T, VEC>::load(vj, i, vv);
                    if (pre_pending) {
                        double pv[VEC];
                        VecIO<T, VEC>::load(a.pre_vec, i, pv);
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[u] = fma(-pre_c, pv[u], qv[u]);
                    }
                    if (vp) {
                        double pv[VEC];
                        VecIO<T, VEC>::load(vp, i, pv);
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[u] = fma(-c_prev, pv[u], qv[u]);
                    }
                    if (modify) {
                        VecIO<T, VEC>::store(q, i, qv);
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[u] = round_as<T>(qv[u]);
                    }
#pragma unroll
                    for (int u = 0; u < VEC; ++u) acc = fma(vv[u], qv[u], acc);
                }
                if (tail_cta) {
                    for (long long i = tail0; i < n; i += blockDim.x) {
                        double qe = (double)q[i];
                        if (pre_pending) qe = fma(-pre_c, (double)a.pre_vec[i], qe);
                        if (vp) qe = fma(-c_prev, (double)vp[i], qe);
                        if (modify) {
                            q[i] = (T)qe;
                            qe = (double)q[i];
                        }
                        acc = fma((double)vj[i], qe, acc);
                    }
                }
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
        if (vp || pre_pending || want_nrm) {
            for (long long i = i0; i < nvec; i += stride) {
                double qv[VEC];
                VecIO<T, VEC>::loadrw(q, i, qv);
                if (pre_pending) {
                    double pv[VEC];
                    VecIO<T, VEC>::load(a.pre_vec, i, pv);
#pragma unroll
                    for (int u = 0; u < VEC; ++u) qv[u] = fma(-pre_c, pv[u], qv[u]);
                }
                if (vp) {
                    double pv[VEC];
                    VecIO<T, VEC>::load(vp, i, pv);
#pragma unroll
                    for (int u = 0; u < VEC; ++u) qv[u] = fma(-c_prev, pv[u], qv[u]);
                }
                if (vp || pre_pending) {
                    VecIO<T, VEC>::store(q, i, qv);
#pragma unroll
                    for (int u = 0; u < VEC; ++u) qv[u] = round_as<T>(qv[u]);
                }
                if (want_nrm) {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) nrm2_part = fma(qv[u], qv[u], nrm2_part);
                }
            }
            if (tail_cta) {
                for (long long i = tail0; i < n; i += blockDim.x) {
                    double qe = (double)q[i];
                    if (pre_pending) qe = fma(-pre_c, (double)a.pre_vec[i], qe);
                    if (vp) qe = fma(-c_prev, (double)vp[i], qe);
                    if (vp || pre_pending) {
                        q[i] = (T)qe;
                        qe = (double)q[i];
                    }
                    if (want_nrm) nrm2_part = fma(qe, qe, nrm2_part);
                }
            }
        }
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
        if (a.vnext != nullptr) {
            long long i = i0;
            if (CU) {
                // four independent loads in flight per thread before the first store
                for (; i + 3 * stride < nvec; i += 4 * stride) {
                    double qv[4][VEC];
#pragma unroll
                    for (int r = 0; r < 4; ++r) VecIO<T, VEC>::loadrw(q, i + r * stride, qv[r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
#pragma unroll
                        for (int u = 0; u < VEC; ++u) qv[r][u] = nrm > 0.0 ? qv[r][u] / nrm : 0.0;
                        VecIO<T, VEC>::store(a.vnext, i + r * stride, qv[r]);
                    }
                }
            }
            for (; i < nvec; i += stride) {
                double qv[VEC];
                VecIO<T, VEC>::loadrw(q, i, qv);
#pragma unroll
                for (int u = 0; u < VEC; ++u) qv[u] = nrm > 0.0 ? qv[u] / nrm : 0.0;
                VecIO<T, VEC>::store(a.vnext, i, qv);
            }
            if (tail_cta)
                for (long long i = tail0; i < n; i += blockDim.x)
                    a.vnext[i] = (T)(nrm > 0.0 ? (double)q[i] / nrm : 0.0);
        }
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
    __shared__ double sm[32];
    __shared__ double c_s[KRY_MAX_SLOTS];
    __shared__ double t_s[KRY_MAX_SLOTS];
    const long long n = p.n;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tail0 = nvec * VEC + threadIdx.x;
    const bool tail_cta = (blockIdx.x == 0);
    const int d = p.d;
    T* av = p.a;
    int buf = 0;
    for (int iter = 0; iter < p.iterations; ++iter) {
        for (int jb = 0; jb < d; jb += ORTH_JT) {
            double acc[ORTH_JT];
#pragma unroll
            for (int t = 0; t < ORTH_JT; ++t) acc[t] = 0.0;
            for (long long i = i0; i < nvec; i += stride) {
                double qv[VEC];
                VecIO<T, VEC>::loadrw(av, i, qv);
#pragma unroll
                for (int tb = 0; tb < ORTH_JT; tb += 8) {
                    if (jb + tb < d) {
                        double vv[8][VEC];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            int j = jb + tb + t;
                            j = j < d ? j : d - 1;
                            VecIO<T, VEC>::load(p.W + (long long)j * p.ldw, i, vv[t]);
                        }
#pragma unroll
                        for (int t = 0; t < 8; ++t)
#pragma unroll
                            for (int u = 0; u < VEC; ++u) acc[tb + t] = fma(vv[t][u], qv[u], acc[tb + t]);
                    }
                }
            }
            if (tail_cta) {
                for (long long i = tail0; i < n; i += blockDim.x) {
                    const double qe = (double)av[i];
#pragma unroll
                    for (int t = 0; t < ORTH_JT; ++t)
                        if (jb + t < d) acc[t] = fma((double)p.W[(long long)(jb + t) * p.ldw + i], qe, acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < ORTH_JT; ++t) {
                if (jb + t < d) {
                    double s = kry_block_sum(acc[t], sm);
                    if (threadIdx.x == 0) partial_slot(p.partials, buf, jb + t)[blockIdx.x] = s;
                }
            }
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
        // a -= V x
        for (long long i = i0; i < nvec; i += stride) {
            double qv[VEC];
            VecIO<T, VEC>::loadrw(av, i, qv);
            // reference forms Pa = V.dot(x) first and then subtracts (utils.py:549, 621)
            double pa[VEC];
#pragma unroll
            for (int u = 0; u < VEC; ++u) pa[u] = 0.0;
            for (int jb = 0; jb < d; jb += 8) {
                double vv[8][VEC];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    int j = jb + t < d ? jb + t : d - 1;
                    VecIO<T, VEC>::load(p.V + (long long)j * p.ldv, i, vv[t]);
                }
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (jb + t < d) {
                        const double c = t_s[jb + t];
#pragma unroll
                        for (int u = 0; u < VEC; ++u) pa[u] = fma(c, vv[t][u], pa[u]);
                    }
            }
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[u] -= pa[u];
            VecIO<T, VEC>::store(av, i, qv);
        }
        if (tail_cta) {
            for (long long i = tail0; i < n; i += blockDim.x) {
                double pa = 0.0;
                for (int j = 0; j < d; ++j) pa = fma(t_s[j], (double)p.V[(long long)j * p.ldv + i], pa);
                av[i] = (T)((double)av[i] - pa);
            }
        }
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

// KRY_ORTH_SMALLK=1 (measurement switch, default off): calls that involve few basis vectors per
// sweep -- block CGS against <= 4 vectors, every exact-MGS / Lanczos call -- use the JT = 4
// instantiation (higher occupancy, see orth_kernel).  Single-GPU only.
// Value: 1 = threshold 4 (one register tile); any other n > 1 = use the variant up to n vectors
// (q is then re-read once per 4-vector tile in the dot phase: a bandwidth-for-occupancy trade to measure).
static int orth_smallk_threshold() {
    static int state = -1;
    if (state < 0) {
        const char* e = getenv("KRY_ORTH_SMALLK");
        int v = e ? atoi(e) : 0;
        state = v <= 0 ? 0 : (v == 1 ? 4 : (v > KRY_MAX_SLOTS ? KRY_MAX_SLOTS : v));
    }
    return state;
}

static bool orth_cunroll_enabled() {
    static int state = -1;
    if (state < 0) {
        const char* e = getenv("KRY_ORTH_CUNROLL");
        state = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return state == 1;
}

template <typename T>
static int orth_launch_small(kry_ctx* ctx, OrthArgs<T>& a, bool al) {
    const int W = VecWidth<T>::value;
    static int blocks_per_sm[2] = {0, 0};        // [aligned, unaligned] instantiation
    const int which = al ? 0 : 1;
    if (blocks_per_sm[which] == 0) {
        int nb = 0, rc;
        if (al) rc = max_blocks_of(orth_kernel<T, W, false, 4>, &nb);
        else rc = max_blocks_of(orth_kernel<T, 1, false, 4>, &nb);
        if (rc) return rc;
        KRY_REQUIRE(nb >= 1, "small-tile orth kernel does not fit");
        blocks_per_sm[which] = nb;
    }
    void* args[] = {&a};
    const int max_blocks = blocks_per_sm[which] * ctx->sm_count;
    const int g = coop_grid(al ? a.n / W : a.n, max_blocks);
    void* k = al ? (void*)orth_kernel<T, W, false, 4> : (void*)orth_kernel<T, 1, false, 4>;
    KRY_CHECK_CUDA(cudaLaunchCooperativeKernel(k, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T>
static int orth_launch(kry_ctx* ctx, OrthArgs<T>& a, int max_blocks) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(a.Vdot) && kry_aligned16(a.Vsub) && kry_aligned16(a.q) && (a.ldv % W == 0) &&
              (!a.pre_vec || kry_aligned16(a.pre_vec)) && (!a.vnext || kry_aligned16(a.vnext));
    void* args[] = {&a};
    const bool peer = a.peer.world > 1;
    const int small_thr = orth_smallk_threshold();
    if (!peer && small_thr > 0 && (a.algo == KRY_ORTH_MGS || a.nv - a.j0 <= small_thr))
        return orth_launch_small<T>(ctx, a, al);
    if (!peer && orth_cunroll_enabled()) {
        // same kernel, same grid; only phase C differs
        const int g = coop_grid(al ? a.n / W : a.n, max_blocks);
        void* k = al ? (void*)orth_kernel<T, W, false, 16, true> : (void*)orth_kernel<T, 1, false, 16, true>;
        KRY_CHECK_CUDA(cudaLaunchCooperativeKernel(k, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
        KRY_LAUNCHED(ctx);
        return KRY_OK;
    }
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
