// Device code of kry_orth_fused / kry_orth_fused_dist / kry_project (csrc/kry_orth.cu): the two cooperative kernels.
// A header of its own so that the CPU test tier can compile exactly this code for the host over the CUDA execution
// emulator (tests/csrc/cuda_emul, tests/test_orth_emul_cpu.py).
#pragma once
#include "kry_common.cuh"
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

