// Device code of kry_orth_fused_z (csrc/kry_cplx.cu): complex sweeps and the cooperative kernel.  A header of
// its own so that the CPU test tier can compile exactly this code for the host over a small CUDA execution
// emulator (tests/csrc/cuda_emul, tests/test_cplx_emul_cpu.py).
#pragma once
#include "kry_common.cuh"
#include "kry_sweeps.cuh"

#ifndef KRY_ZTYPE
#define KRY_ZTYPE
typedef double2 Z;
#endif

#define ZJT 8   // complex vectors per dots tile: 16 accumulators, the register tile of ORTH_JT real rows

__device__ __forceinline__ Z zld(const Z* __restrict__ p, long long i) { return __ldg(p + i); }
__device__ __forceinline__ Z zldrw(const Z* p, long long i) { return p[i]; }

// (re, im) += conj(v) * q
__device__ __forceinline__ void zdot_acc(const Z v, const Z q, double& re, double& im) {
    re = fma(v.x, q.x, re);
    re = fma(v.y, q.y, re);
    im = fma(v.x, q.y, im);
    im = fma(-v.y, q.x, im);
}

// q -= (cr + i ci) * v
__device__ __forceinline__ void zupd(Z& q, const double cr, const double ci, const Z v) {
    q.x = fma(-cr, v.x, q.x);
    q.x = fma(ci, v.y, q.x);
    q.y = fma(-cr, v.y, q.y);
    q.y = fma(-ci, v.x, q.y);
}

// loads in flight per thread: (NT + 1) * U 16-byte loads, >= 8 wherever the registers allow (kry_sweeps.cuh)
__host__ __device__ constexpr int z_unroll(int nt) { return nt <= 1 ? 4 : (nt == 2 ? 3 : (nt <= 4 ? 2 : 1)); }

// slots [slot0, slot0 + 2 NT): Re, Im of <V[t], q> = sum_i conj(V[t][i]) q[i], t < NT (1 <= NT <= 8), one pass over q
template <int NT>
__device__ __forceinline__ void zdots_pass(const Z* __restrict__ V, long long ldv, const Z* q, long long n, double* red,
                                           double* partials, int buf, int slot0) {
    constexpr int U = z_unroll(NT);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[2 * NT];
#pragma unroll
    for (int t = 0; t < 2 * NT; ++t) acc[t] = 0.0;
    if (U > 1) {
        for (; i + (U - 1) * stride < n; i += U * stride) {
            Z qv[U], vv[U][NT];
#pragma unroll
            for (int r = 0; r < U; ++r) {
                qv[r] = zldrw(q, i + r * stride);
#pragma unroll
                for (int t = 0; t < NT; ++t) vv[r][t] = zld(V + (long long)t * ldv, i + r * stride);
            }
#pragma unroll
            for (int r = 0; r < U; ++r)
#pragma unroll
                for (int t = 0; t < NT; ++t) zdot_acc(vv[r][t], qv[r], acc[2 * t], acc[2 * t + 1]);
        }
    }
    for (; i < n; i += stride) {
        const Z qv = zldrw(q, i);
        Z vv[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) vv[t] = zld(V + (long long)t * ldv, i);
#pragma unroll
        for (int t = 0; t < NT; ++t) zdot_acc(vv[t], qv, acc[2 * t], acc[2 * t + 1]);
    }
    reduce_store<2 * NT>(acc, red, partials, buf, slot0);
}

__device__ __forceinline__ void zdots_dispatch(int nt, const Z* V, long long ldv, const Z* q, long long n, double* red,
                                               double* partials, int buf, int slot0) {
    switch (nt) {
#define KRY_ZDOTS_CASE(NT) case NT: zdots_pass<NT>(V, ldv, q, n, red, partials, buf, slot0); break;
        KRY_ZDOTS_CASE(1) KRY_ZDOTS_CASE(2) KRY_ZDOTS_CASE(3) KRY_ZDOTS_CASE(4)
        KRY_ZDOTS_CASE(5) KRY_ZDOTS_CASE(6) KRY_ZDOTS_CASE(7) KRY_ZDOTS_CASE(8)
#undef KRY_ZDOTS_CASE
        default: break;
    }
}

// one element: q -= sum_j c[j] V[j] over full blocks of 8 vectors and an exact remainder block of R
template <int R>
__device__ __forceinline__ void zupdate_pack(const Z* __restrict__ V, long long ldv, int nfull, const double* c_s,
                                             long long i, Z& qv) {
    for (int jb = 0; jb < nfull; jb += 8) {
        Z vv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) vv[t] = zld(V + (long long)(jb + t) * ldv, i);
#pragma unroll
        for (int t = 0; t < 8; ++t) zupd(qv, c_s[2 * (jb + t)], c_s[2 * (jb + t) + 1], vv[t]);
    }
    if (R > 0) {
        Z vv[R > 0 ? R : 1];
#pragma unroll
        for (int t = 0; t < R; ++t) vv[t] = zld(V + (long long)(nfull + t) * ldv, i);
#pragma unroll
        for (int t = 0; t < R; ++t) zupd(qv, c_s[2 * (nfull + t)], c_s[2 * (nfull + t) + 1], vv[t]);
    }
}

// q -= V c for cnt vectors (cnt = nfull + R, nfull a multiple of 8, 0 <= R < 8); returns this thread's share of
// ||q||^2 when want_nrm
template <int R>
__device__ __forceinline__ double zupdate_pass(const Z* __restrict__ V, long long ldv, int cnt, const double* c_s, Z* q,
                                               long long n, bool want_nrm) {
    const int nfull = cnt - R;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double nrm2 = 0.0;
    constexpr int U = (R >= 1 && R <= 4) ? z_unroll(R) : 1;
    if (U > 1 && nfull == 0) {
        for (; i + (U - 1) * stride < n; i += U * stride) {
            Z qv[U], vv[U][R > 0 ? R : 1];
#pragma unroll
            for (int r = 0; r < U; ++r) {
                qv[r] = zldrw(q, i + r * stride);
#pragma unroll
                for (int t = 0; t < R; ++t) vv[r][t] = zld(V + (long long)t * ldv, i + r * stride);
            }
#pragma unroll
            for (int r = 0; r < U; ++r) {
#pragma unroll
                for (int t = 0; t < R; ++t) zupd(qv[r], c_s[2 * t], c_s[2 * t + 1], vv[r][t]);
                q[i + r * stride] = qv[r];
                if (want_nrm) {
                    nrm2 = fma(qv[r].x, qv[r].x, nrm2);
                    nrm2 = fma(qv[r].y, qv[r].y, nrm2);
                }
            }
        }
    }
    for (; i < n; i += stride) {
        Z qv = zldrw(q, i);
        zupdate_pack<R>(V, ldv, nfull, c_s, i, qv);
        q[i] = qv;
        if (want_nrm) {
            nrm2 = fma(qv.x, qv.x, nrm2);
            nrm2 = fma(qv.y, qv.y, nrm2);
        }
    }
    return nrm2;
}

__device__ __forceinline__ double zupdate_dispatch(const Z* V, long long ldv, int cnt, const double* c_s, Z* q,
                                                   long long n, bool want_nrm) {
    switch (cnt & 7) {
        case 1: return zupdate_pass<1>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 2: return zupdate_pass<2>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 3: return zupdate_pass<3>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 4: return zupdate_pass<4>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 5: return zupdate_pass<5>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 6: return zupdate_pass<6>(V, ldv, cnt, c_s, q, n, want_nrm);
        case 7: return zupdate_pass<7>(V, ldv, cnt, c_s, q, n, want_nrm);
        default: return zupdate_pass<0>(V, ldv, cnt, c_s, q, n, want_nrm);
    }
}

// One sweep of exact modified Gram-Schmidt: q -= c_prev * vp (optional, the pending update of the previous
// vector), store q if modified; (are, aim) = this thread's share of <vj, q> (vj == nullptr: are = its share of
// ||q||^2 when want_nrm).  Unrolled 2x over the stride.
__device__ __forceinline__ void zmgs_pass(const Z* __restrict__ vj, const Z* __restrict__ vp, double cr, double ci, Z* q,
                                          long long n, bool want_nrm, double& are, double& aim) {
    constexpr int U = 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double re = 0.0, im = 0.0;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        Z qv[U], vv[U], pv[U];
#pragma unroll
        for (int r = 0; r < U; ++r) {
            qv[r] = zldrw(q, i + r * stride);
            if (vj) vv[r] = zld(vj, i + r * stride);
            if (vp) pv[r] = zld(vp, i + r * stride);
        }
#pragma unroll
        for (int r = 0; r < U; ++r) {
            if (vp) {
                zupd(qv[r], cr, ci, pv[r]);
                q[i + r * stride] = qv[r];
            }
            if (vj) {
                zdot_acc(vv[r], qv[r], re, im);
            } else if (want_nrm) {
                re = fma(qv[r].x, qv[r].x, re);
                re = fma(qv[r].y, qv[r].y, re);
            }
        }
    }
    for (; i < n; i += stride) {
        Z qv = zldrw(q, i);
        if (vp) {
            const Z pv = zld(vp, i);
            zupd(qv, cr, ci, pv);
            q[i] = qv;
        }
        if (vj) {
            const Z vv = zld(vj, i);
            zdot_acc(vv, qv, re, im);
        } else if (want_nrm) {
            re = fma(qv.x, qv.x, re);
            re = fma(qv.y, qv.y, re);
        }
    }
    are = re;
    aim = im;
}

struct ZOrthArgs {
    long long n;
    const Z* Vdot;
    const Z* Vsub;
    long long ldv;
    int j0, nv, passes, algo;
    Z* q;
    double* h;
    double* nrm;
    Z* vnext;
    double* partials;   // [2][KRY_MAX_SLOTS][KRY_MAX_PARTIAL_BLOCKS]
};

// Same phases and the same grid-wide dependencies as orth_kernel (kry_orth.cu): dots, grid.sync, fixed-order
// final sums recomputed identically by every CTA, update, norm, normalised store.
__global__ void __launch_bounds__(KRY_THREADS, 2) zorth_kernel(ZOrthArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[32];
    __shared__ double red[ORTH_JT * 8];
    __shared__ double c_s[KRY_MAX_SLOTS];
    const long long n = a.n, ldv = a.ldv;
    Z* q = a.q;
    int buf = 0;
    const int cnt = a.nv - a.j0;
    double nrm2_part = 0.0;

    if (a.algo == KRY_ORTH_CGS) {
        for (int pass = 0; pass < a.passes; ++pass) {
            // ---- phase A: block dots, up to 8 complex vectors per pass over q ----
            for (int jb = 0; jb < cnt; jb += ZJT) {
                const int nt = cnt - jb < ZJT ? cnt - jb : ZJT;
                zdots_dispatch(nt, a.Vdot + (long long)(a.j0 + jb) * ldv, ldv, q, n, red, a.partials, buf, 2 * jb);
            }
            grid.sync();
            reduce_slots(a.partials, buf, 2 * cnt, c_s);
            if (blockIdx.x == 0)
                for (int s = threadIdx.x; s < 2 * cnt; s += blockDim.x) a.h[2 * a.j0 + s] += c_s[s];
            buf ^= 1;
            // ---- phase B: q -= Vsub c (+ ||q||^2 in the last pass) ----
            const bool want_nrm = (a.nrm != nullptr) && (pass == a.passes - 1);
            if (cnt > 0 || want_nrm)
                nrm2_part = zupdate_dispatch(a.Vsub + (long long)a.j0 * ldv, ldv, cnt, c_s, q, n, want_nrm);
            __syncthreads();   // c_s is rewritten by the next pass
        }
    } else {
        // ---- exact modified Gram-Schmidt with COMPLEX coefficients (krypy/utils.py:1012-1029): one dependent
        //      reduction per basis vector; the update with vector j-1 is fused into the sweep of <v_j, q> ----
        double cr = 0.0, ci = 0.0;
        int j_prev = -1;
        for (int pass = 0; pass < a.passes; ++pass) {
            for (int j = a.j0; j < a.nv; ++j) {
                const Z* vj = a.Vdot + (long long)j * ldv;
                const Z* vp = j_prev >= 0 ? a.Vsub + (long long)j_prev * ldv : nullptr;
                double acc[2];
                zmgs_pass(vj, vp, cr, ci, q, n, false, acc[0], acc[1]);
                reduce_store<2>(acc, red, a.partials, buf, 0);
                grid.sync();
                reduce_slots(a.partials, buf, 2, c_s);
                cr = c_s[0];
                ci = c_s[1];
                __syncthreads();
                j_prev = j;
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    a.h[2 * j] += cr;
                    a.h[2 * j + 1] += ci;
                }
                buf ^= 1;
            }
        }
        const Z* vp = j_prev >= 0 ? a.Vsub + (long long)j_prev * ldv : nullptr;
        const bool want_nrm = (a.nrm != nullptr);
        if (vp || want_nrm) {
            double dummy;
            zmgs_pass(nullptr, vp, cr, ci, q, n, want_nrm, nrm2_part, dummy);
        }
    }

    // ---- norm and normalised store ----
    if (a.nrm != nullptr) {
        double s = kry_block_sum(nrm2_part, sm);
        if (threadIdx.x == 0) partial_slot(a.partials, buf, 0)[blockIdx.x] = s;
        grid.sync();
        const double nrm2 = reduce_slot(a.partials, buf, 0, sm);
        const double nrm = sqrt(nrm2);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.nrm[0] = nrm;
        }
        if (a.vnext != nullptr)
            scale_pass<double, 2>(reinterpret_cast<const double*>(q), reinterpret_cast<double*>(a.vnext), 2 * n, nrm);
    }
}

