// Sweep building blocks of the tall-skinny kernels (fused Gram-Schmidt, deflation projector,
// row-partitioned Gram-Schmidt): block dots, block updates, normalised store, exact-MGS sweep.
// Shared by kry_orth.cu and kry_dist.cu.
#pragma once
#include "kry_common.cuh"

#define ORTH_JT 16

// value as it reads back after being stored as T (identity for double)
template <typename T> __device__ __forceinline__ double round_as(double v) { return (double)(T)v; }


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

// ---------------------------------------------------------------------------
// Sweep building blocks.  Every thread owns the same elements (grid-stride map over 16-byte
// packs) in every phase.  The loops are specialised on the EXACT number of basis vectors they
// touch (no clamped duplicate loads for remainder tiles) and unrolled over the grid stride (U)
// when only a few vectors are involved, so that every thread keeps >= 8 independent 16-byte
// loads in flight: with 2 CTAs/SM x 256 threads that is what it takes to cover the HBM latency
// (measured: 2 loads in flight per thread = 45 % of the copy bandwidth at nv = 1).
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int orth_unroll(int nt) { return nt <= 1 ? 4 : (nt == 2 ? 3 : (nt <= 4 ? 2 : 1)); }

// CTA reduction of NT accumulators with ONE barrier pair; per-CTA partials to slots [slot0, slot0+NT)
template <int NT>
__device__ __forceinline__ void reduce_store(double (&acc)[NT], double* red /*[16*8]*/, double* partials, int buf,
                                             int slot0) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = kry_warp_sum(acc[t]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NT; ++t) red[t * 8 + w] = acc[t];
    }
    __syncthreads();
    if (threadIdx.x < NT) {
        double s = 0.0;
        for (int ww = 0; ww < nw; ++ww) s += red[threadIdx.x * 8 + ww];     // fixed warp order
        partial_slot(partials, buf, slot0 + threadIdx.x)[blockIdx.x] = s;
    }
}

// acc[t] += <V[t], q> over this thread's elements, t < NT (1 <= NT <= 16), one pass over q.
// SQ: <q, q> as one more sum in slot slot0 + NT (row-partitioned one-wait step: the norm after the update
// follows from it); red then holds (NT + 1) * 8 doubles.
template <typename T, int VEC, int NT, bool SQ = false>
__device__ __forceinline__ void dots_pass(const T* __restrict__ V, long long ldv, const T* q, long long n,
                                          double* red, double* partials, int buf, int slot0) {
    constexpr int U = orth_unroll(NT);
    constexpr int B0 = NT < 8 ? NT : 8, B1 = NT - B0;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[NT + (SQ ? 1 : 0)];
#pragma unroll
    for (int t = 0; t < NT + (SQ ? 1 : 0); ++t) acc[t] = 0.0;
    if (U > 1) {
        for (; i + (U - 1) * stride < nvec; i += U * stride) {
            double qv[U][VEC], vv[U][B0][VEC];
#pragma unroll
            for (int r = 0; r < U; ++r) {
                VecIO<T, VEC>::loadrw(q, i + r * stride, qv[r]);
#pragma unroll
                for (int t = 0; t < B0; ++t) VecIO<T, VEC>::load(V + (long long)t * ldv, i + r * stride, vv[r][t]);
            }
#pragma unroll
            for (int r = 0; r < U; ++r) {
#pragma unroll
                for (int t = 0; t < B0; ++t)
#pragma unroll
                    for (int u = 0; u < VEC; ++u) acc[t] = fma(vv[r][t][u], qv[r][u], acc[t]);
                if (SQ) {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) acc[NT] = fma(qv[r][u], qv[r][u], acc[NT]);
                }
            }
        }
    }
    for (; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
        {
            double vv[B0][VEC];
#pragma unroll
            for (int t = 0; t < B0; ++t) VecIO<T, VEC>::load(V + (long long)t * ldv, i, vv[t]);
#pragma unroll
            for (int t = 0; t < B0; ++t)
#pragma unroll
                for (int u = 0; u < VEC; ++u) acc[t] = fma(vv[t][u], qv[u], acc[t]);
        }
        if (B1 > 0) {
            double vv[B1 > 0 ? B1 : 1][VEC];
#pragma unroll
            for (int t = 0; t < B1; ++t) VecIO<T, VEC>::load(V + (long long)(B0 + t) * ldv, i, vv[t]);
#pragma unroll
            for (int t = 0; t < B1; ++t)
#pragma unroll
                for (int u = 0; u < VEC; ++u) acc[B0 + t] = fma(vv[t][u], qv[u], acc[B0 + t]);
        }
        if (SQ) {
#pragma unroll
            for (int u = 0; u < VEC; ++u) acc[NT] = fma(qv[u], qv[u], acc[NT]);
        }
    }
    if (blockIdx.x == 0) {   // scalar tail
        for (long long e = nvec * VEC + threadIdx.x; e < n; e += blockDim.x) {
            const double qe = (double)q[e];
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t] = fma((double)V[(long long)t * ldv + e], qe, acc[t]);
            if (SQ) acc[NT] = fma(qe, qe, acc[NT]);
        }
    }
    reduce_store<NT + (SQ ? 1 : 0)>(acc, red, partials, buf, slot0);
}

template <typename T, int VEC, bool SQ = false>
__device__ __forceinline__ void dots_dispatch(int nt, const T* V, long long ldv, const T* q, long long n, double* red,
                                              double* partials, int buf, int slot0) {
    switch (nt) {
#define KRY_DOTS_CASE(NT) case NT: dots_pass<T, VEC, NT, SQ>(V, ldv, q, n, red, partials, buf, slot0); break;
        KRY_DOTS_CASE(1) KRY_DOTS_CASE(2) KRY_DOTS_CASE(3) KRY_DOTS_CASE(4) KRY_DOTS_CASE(5) KRY_DOTS_CASE(6)
        KRY_DOTS_CASE(7) KRY_DOTS_CASE(8) KRY_DOTS_CASE(9) KRY_DOTS_CASE(10) KRY_DOTS_CASE(11) KRY_DOTS_CASE(12)
        KRY_DOTS_CASE(13) KRY_DOTS_CASE(14) KRY_DOTS_CASE(15) KRY_DOTS_CASE(16)
#undef KRY_DOTS_CASE
        default: break;
    }
}

// one element pack: q -= sum_j c[j] V[j] over full blocks of 8 and an exact remainder block of R vectors.
// FORM_FIRST: accumulate Pa = sum_j c[j] V[j] first and subtract once (the projector's rounding,
// krypy/utils.py:549, 621) instead of updating q vector by vector (Gram-Schmidt's).
template <typename T, int VEC, int R, bool FORM_FIRST>
__device__ __forceinline__ void update_pack(const T* __restrict__ V, long long ldv, int nfull, const double* c_s,
                                            long long i, double (&qv)[VEC]) {
    double pa[VEC];
#pragma unroll
    for (int u = 0; u < VEC; ++u) pa[u] = 0.0;
    for (int jb = 0; jb < nfull; jb += 8) {
        double vv[8][VEC];
#pragma unroll
        for (int t = 0; t < 8; ++t) VecIO<T, VEC>::load(V + (long long)(jb + t) * ldv, i, vv[t]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const double c = c_s[jb + t];
#pragma unroll
            for (int u = 0; u < VEC; ++u) {
                if (FORM_FIRST) pa[u] = fma(c, vv[t][u], pa[u]);
                else qv[u] = fma(-c, vv[t][u], qv[u]);
            }
        }
    }
    if (R > 0) {
        double vv[R > 0 ? R : 1][VEC];
#pragma unroll
        for (int t = 0; t < R; ++t) VecIO<T, VEC>::load(V + (long long)(nfull + t) * ldv, i, vv[t]);
#pragma unroll
        for (int t = 0; t < R; ++t) {
            const double c = c_s[nfull + t];
#pragma unroll
            for (int u = 0; u < VEC; ++u) {
                if (FORM_FIRST) pa[u] = fma(c, vv[t][u], pa[u]);
                else qv[u] = fma(-c, vv[t][u], qv[u]);
            }
        }
    }
    if (FORM_FIRST) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) qv[u] -= pa[u];
    }
}

// q -= V c for cnt vectors (cnt = nfull + R, nfull a multiple of 8, 0 <= R < 8); returns this thread's
// share of ||q||^2 (of the values as stored) when want_nrm
template <typename T, int VEC, int R, bool FORM_FIRST>
__device__ __forceinline__ double update_pass(const T* __restrict__ V, long long ldv, int cnt, const double* c_s, T* q,
                                              long long n, bool want_nrm) {
    const int nfull = cnt - R;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double nrm2 = 0.0;
    // few vectors: unroll over the stride to keep enough loads in flight
    constexpr int U = (R >= 1 && R <= 4) ? orth_unroll(R) : 1;
    if (U > 1 && nfull == 0) {
        for (; i + (U - 1) * stride < nvec; i += U * stride) {
            double qv[U][VEC], vv[U][R > 0 ? R : 1][VEC];
#pragma unroll
            for (int r = 0; r < U; ++r) {
                VecIO<T, VEC>::loadrw(q, i + r * stride, qv[r]);
#pragma unroll
                for (int t = 0; t < R; ++t) VecIO<T, VEC>::load(V + (long long)t * ldv, i + r * stride, vv[r][t]);
            }
#pragma unroll
            for (int r = 0; r < U; ++r) {
                double pa[VEC];
#pragma unroll
                for (int u = 0; u < VEC; ++u) pa[u] = 0.0;
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const double c = c_s[t];
#pragma unroll
                    for (int u = 0; u < VEC; ++u) {
                        if (FORM_FIRST) pa[u] = fma(c, vv[r][t][u], pa[u]);
                        else qv[r][u] = fma(-c, vv[r][t][u], qv[r][u]);
                    }
                }
                if (FORM_FIRST) {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) qv[r][u] -= pa[u];
                }
                VecIO<T, VEC>::store(q, i + r * stride, qv[r]);
                if (want_nrm) {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) {
                        const double v = round_as<T>(qv[r][u]);
                        nrm2 = fma(v, v, nrm2);
                    }
                }
            }
        }
    }
    for (; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
        update_pack<T, VEC, R, FORM_FIRST>(V, ldv, nfull, c_s, i, qv);
        VecIO<T, VEC>::store(q, i, qv);
        if (want_nrm) {
#pragma unroll
            for (int u = 0; u < VEC; ++u) {
                const double v = round_as<T>(qv[u]);
                nrm2 = fma(v, v, nrm2);
            }
        }
    }
    if (blockIdx.x == 0) {   // scalar tail
        for (long long e = nvec * VEC + threadIdx.x; e < n; e += blockDim.x) {
            double qe = (double)q[e];
            if (FORM_FIRST) {
                double pa = 0.0;
                for (int j = 0; j < cnt; ++j) pa = fma(c_s[j], (double)V[(long long)j * ldv + e], pa);
                qe -= pa;
            } else {
                for (int j = 0; j < cnt; ++j) qe = fma(-c_s[j], (double)V[(long long)j * ldv + e], qe);
            }
            q[e] = (T)qe;
            qe = (double)q[e];
            if (want_nrm) nrm2 = fma(qe, qe, nrm2);
        }
    }
    return nrm2;
}

template <typename T, int VEC, bool FORM_FIRST>
__device__ __forceinline__ double update_dispatch(const T* V, long long ldv, int cnt, const double* c_s, T* q,
                                                  long long n, bool want_nrm) {
    switch (cnt & 7) {
        case 1: return update_pass<T, VEC, 1, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 2: return update_pass<T, VEC, 2, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 3: return update_pass<T, VEC, 3, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 4: return update_pass<T, VEC, 4, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 5: return update_pass<T, VEC, 5, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 6: return update_pass<T, VEC, 6, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 7: return update_pass<T, VEC, 7, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
        default: return update_pass<T, VEC, 0, FORM_FIRST>(V, ldv, cnt, c_s, q, n, want_nrm);
    }
}

// vnext = (q - V c) / nrm in ONE sweep (q is only read): the Gram-Schmidt update of update_pass followed by
// the normalised store of scale_pass, element by element with the same roundings (q - V c is rounded to T
// before the division, as if it had been stored).  Returns this thread's share of ||q - V c||^2 (of the
// rounded values).  store == false: the norm only.  CTAs [0, nblk) take part (grid-stride map over nblk CTAs).
template <typename T, int VEC, int R>
__device__ __forceinline__ double update_scale_pass(const T* __restrict__ V, long long ldv, int cnt, const double* c_s,
                                                    const T* q, T* vnext, long long n, double nrm, bool store, int nblk) {
    const int nfull = cnt - R;
    const long long nvec = n / VEC;
    const long long stride = (long long)nblk * blockDim.x;      // CTAs [0, nblk) sweep
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double nrm2 = 0.0;
    constexpr int U = (R >= 1 && R <= 4) ? orth_unroll(R) : 1;
    if (U > 1 && nfull == 0) {
        for (; i + (U - 1) * stride < nvec; i += U * stride) {
            double qv[U][VEC], vv[U][R > 0 ? R : 1][VEC];
#pragma unroll
            for (int r = 0; r < U; ++r) {
                VecIO<T, VEC>::loadrw(q, i + r * stride, qv[r]);
#pragma unroll
                for (int t = 0; t < R; ++t) VecIO<T, VEC>::load(V + (long long)t * ldv, i + r * stride, vv[r][t]);
            }
#pragma unroll
            for (int r = 0; r < U; ++r) {
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const double c = c_s[t];
#pragma unroll
                    for (int u = 0; u < VEC; ++u) qv[r][u] = fma(-c, vv[r][t][u], qv[r][u]);
                }
#pragma unroll
                for (int u = 0; u < VEC; ++u) {
                    const double v = round_as<T>(qv[r][u]);
                    nrm2 = fma(v, v, nrm2);
                    qv[r][u] = nrm > 0.0 ? v / nrm : 0.0;
                }
                if (store) VecIO<T, VEC>::store(vnext, i + r * stride, qv[r]);
            }
        }
    }
    for (; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
        update_pack<T, VEC, R, false>(V, ldv, nfull, c_s, i, qv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            const double v = round_as<T>(qv[u]);
            nrm2 = fma(v, v, nrm2);
            qv[u] = nrm > 0.0 ? v / nrm : 0.0;
        }
        if (store) VecIO<T, VEC>::store(vnext, i, qv);
    }
    if (blockIdx.x == 0) {   // scalar tail
        for (long long e = nvec * VEC + threadIdx.x; e < n; e += blockDim.x) {
            double qe = (double)q[e];
            for (int j = 0; j < cnt; ++j) qe = fma(-c_s[j], (double)V[(long long)j * ldv + e], qe);
            qe = round_as<T>(qe);
            nrm2 = fma(qe, qe, nrm2);
            if (store) vnext[e] = (T)(nrm > 0.0 ? qe / nrm : 0.0);
        }
    }
    return nrm2;
}

template <typename T, int VEC>
__device__ __forceinline__ double update_scale_dispatch(const T* V, long long ldv, int cnt, const double* c_s,
                                                        const T* q, T* vnext, long long n, double nrm, bool store,
                                                        int nblk) {
    switch (cnt & 7) {
        case 1: return update_scale_pass<T, VEC, 1>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        case 2: return update_scale_pass<T, VEC, 2>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        case 3: return update_scale_pass<T, VEC, 3>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        case 4: return update_scale_pass<T, VEC, 4>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        case 5: return update_scale_pass<T, VEC, 5>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        case 6: return update_scale_pass<T, VEC, 6>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        case 7: return update_scale_pass<T, VEC, 7>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
        default: return update_scale_pass<T, VEC, 0>(V, ldv, cnt, c_s, q, vnext, n, nrm, store, nblk);
    }
}

// vnext = q / nrm (0 when nrm == 0), four loads in flight per thread
template <typename T, int VEC>
__device__ __forceinline__ void scale_pass(const T* q, T* vnext, long long n, double nrm) {
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        double qv[4][VEC];
#pragma unroll
        for (int r = 0; r < 4; ++r) VecIO<T, VEC>::loadrw(q, i + r * stride, qv[r]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[r][u] = nrm > 0.0 ? qv[r][u] / nrm : 0.0;
            VecIO<T, VEC>::store(vnext, i + r * stride, qv[r]);
        }
    }
    for (; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) qv[u] = nrm > 0.0 ? qv[u] / nrm : 0.0;
        VecIO<T, VEC>::store(vnext, i, qv);
    }
    if (blockIdx.x == 0)
        for (long long e = nvec * VEC + threadIdx.x; e < n; e += blockDim.x)
            vnext[e] = (T)(nrm > 0.0 ? (double)q[e] / nrm : 0.0);
}

// One sweep of exact modified Gram-Schmidt: q -= pre_c*pre (optional), q -= c_prev*vp (optional, the
// pending update of the previous vector), store q if modified, return this thread's share of
// <vj, q> (vj == nullptr: of ||q||^2 when want_nrm, else 0).  Unrolled 2x over the stride.
template <typename T, int VEC>
__device__ __forceinline__ double mgs_pass(const T* __restrict__ vj, const T* __restrict__ vp, double c_prev,
                                           const T* __restrict__ pre, double pre_c, T* q, long long n, bool want_nrm) {
    constexpr int U = 2;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool modify = (vp != nullptr) || (pre != nullptr);
    double acc = 0.0;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        double qv[U][VEC], vv[U][VEC], pv[U][VEC], wv[U][VEC];
#pragma unroll
        for (int r = 0; r < U; ++r) {
            VecIO<T, VEC>::loadrw(q, i + r * stride, qv[r]);
            if (vj) VecIO<T, VEC>::load(vj, i + r * stride, vv[r]);
            if (pre) VecIO<T, VEC>::load(pre, i + r * stride, wv[r]);
            if (vp) VecIO<T, VEC>::load(vp, i + r * stride, pv[r]);
        }
#pragma unroll
        for (int r = 0; r < U; ++r) {
            if (pre) {
#pragma unroll
                for (int u = 0; u < VEC; ++u) qv[r][u] = fma(-pre_c, wv[r][u], qv[r][u]);
            }
            if (vp) {
#pragma unroll
                for (int u = 0; u < VEC; ++u) qv[r][u] = fma(-c_prev, pv[r][u], qv[r][u]);
            }
            if (modify) {
                VecIO<T, VEC>::store(q, i + r * stride, qv[r]);
#pragma unroll
                for (int u = 0; u < VEC; ++u) qv[r][u] = round_as<T>(qv[r][u]);
            }
            if (vj) {
#pragma unroll
                for (int u = 0; u < VEC; ++u) acc = fma(vv[r][u], qv[r][u], acc);
            } else if (want_nrm) {
#pragma unroll
                for (int u = 0; u < VEC; ++u) acc = fma(qv[r][u], qv[r][u], acc);
            }
        }
    }
    for (; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
        if (pre) {
            double wv[VEC];
            VecIO<T, VEC>::load(pre, i, wv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[u] = fma(-pre_c, wv[u], qv[u]);
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
        if (vj) {
            double vv[VEC];
            VecIO<T, VEC>::load(vj, i, vv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) acc = fma(vv[u], qv[u], acc);
        } else if (want_nrm) {
#pragma unroll
            for (int u = 0; u < VEC; ++u) acc = fma(qv[u], qv[u], acc);
        }
    }
    if (blockIdx.x == 0) {   // scalar tail
        for (long long e = nvec * VEC + threadIdx.x; e < n; e += blockDim.x) {
            double qe = (double)q[e];
            if (pre) qe = fma(-pre_c, (double)pre[e], qe);
            if (vp) qe = fma(-c_prev, (double)vp[e], qe);
            if (modify) {
                q[e] = (T)qe;
                qe = (double)q[e];
            }
            if (vj) acc = fma((double)vj[e], qe, acc);
            else if (want_nrm) acc = fma(qe, qe, acc);
        }
    }
    return acc;
}

