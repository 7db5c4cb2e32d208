// Device code of the fused solver updates (csrc/kry_update.cu): CG (kry_cg_update / kry_cg_update_dev,
// kry_cg_scalars, kry_xpby_dev) and MINRES (kry_minres_update).  A header of its own so that the CPU test tier can
// run a CG iteration chain with exactly these kernels over the CUDA execution emulator
// (tests/csrc/cuda_emul, tests/test_cg_emul_cpu.py).
#pragma once
#include "kry_common.cuh"

#ifndef KRY_ROUND_AS_DEFINED
#define KRY_ROUND_AS_DEFINED
template <typename T> __device__ __forceinline__ double round_as(double v) { return (double)(T)v; }
#endif

// krypy/linsys.py:634 (alpha), :655 (yk += alpha p), :658 (Mlrk -= alpha Ap),
// :661 (MMlrk = M Mlrk, diagonal M), :664-665 (rho = <Mlrk, MMlrk>)
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2)
cg_update_kernel(long long n, const T* __restrict__ Ap, const T* __restrict__ p, T* yk, T* r, T* z,
                 const T* __restrict__ dinv, double rho, const double* pAp, double* partials,
                 unsigned int* ticket, double* mailbox, double* st) {
    __shared__ double sm[32];
    __shared__ bool last;
    // st != NULL: the scalars of the recurrence live in device memory (kry_cg_update_dev):
    // st[1] = rho, st[2] = <p,Ap>; out: st[3] = alpha, st[5] = this device's share of the new rho
    const double pap = st ? st[2] : pAp[0];
    if (st) rho = st[1];
    const double alpha = rho / pap;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double pv[VEC], av[VEC], yv[VEC], rv[VEC], zv[VEC];
        VecIO<T, VEC>::load(p, i, pv);
        VecIO<T, VEC>::load(Ap, i, av);
        VecIO<T, VEC>::loadrw(yk, i, yv);
        VecIO<T, VEC>::loadrw(r, i, rv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            yv[u] = fma(alpha, pv[u], yv[u]);
            rv[u] = round_as<T>(fma(-alpha, av[u], rv[u]));
        }
        VecIO<T, VEC>::store(yk, i, yv);
        VecIO<T, VEC>::store(r, i, rv);
        if (dinv) {
            double dv[VEC];
            VecIO<T, VEC>::load(dinv, i, dv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) zv[u] = round_as<T>(dv[u] * rv[u]);
            VecIO<T, VEC>::store(z, i, zv);
        } else {
#pragma unroll
            for (int u = 0; u < VEC; ++u) zv[u] = rv[u];
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) acc = fma(rv[u], zv[u], acc);
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
            yk[i] = (T)fma(alpha, (double)p[i], (double)yk[i]);
            double rv = round_as<T>(fma(-alpha, (double)Ap[i], (double)r[i]));
            r[i] = (T)rv;
            double zv = rv;
            if (dinv) {
                zv = round_as<T>((double)dinv[i] * rv);
                z[i] = (T)zv;
            }
            acc = fma(rv, zv, acc);
        }
    }
    double s = kry_block_sum(acc, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double v = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) v += __ldcg(partials + b);
        double rr = kry_block_sum(v, sm);
        if (threadIdx.x == 0) {
            if (st) {
                st[3] = alpha;
                st[5] = rr;
            } else {
                mailbox[0] = rr;
                mailbox[1] = alpha;
                mailbox[2] = pap;
            }
            *ticket = 0u;
        }
    }
}

// krypy/linsys.py:844-846
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2)
minres_update_kernel(long long n, const T* __restrict__ v, T* w0, const T* __restrict__ w1, T* yk,
                     const double* st) {
    const double R0 = st[8], R1 = st[9], R2 = st[10], yc = st[11];
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double vv[VEC], a0[VEC], a1[VEC], yv[VEC], zv[VEC];
        VecIO<T, VEC>::load(v, i, vv);
        VecIO<T, VEC>::loadrw(w0, i, a0);
        VecIO<T, VEC>::load(w1, i, a1);
        VecIO<T, VEC>::loadrw(yk, i, yv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            zv[u] = round_as<T>(fma(-R1, a1[u], fma(-R0, a0[u], vv[u])) / R2);
            yv[u] = fma(yc, zv[u], yv[u]);
        }
        VecIO<T, VEC>::store(w0, i, zv);
        VecIO<T, VEC>::store(yk, i, yv);
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
            double zv = round_as<T>(fma(-R1, (double)w1[i], fma(-R0, (double)w0[i], (double)v[i])) / R2);
            w0[i] = (T)zv;
            yk[i] = (T)fma(yc, zv, (double)yk[i]);
        }
    }
}

// The scalar recurrence of CG on the device (one small CTA): completes the new rho (row-partitioned runs:
// global sum over NVLink peer memory), shifts rho, forms beta and publishes to the pinned mailbox.
//   st: [0] rho_{k-1}  [1] rho_k  [2] <p,Ap>  [3] alpha  [4] beta = rho_k / rho_{k-1}  [5] local share of the new rho
// rho_k is stored as sqrt(|sum|)^2 -- the reference squares the NORM it computed (linsys.py:664-665).
__global__ void __launch_bounds__(64) cg_scalars_kernel(double* st, double* mailbox, PeerArgs pa) {
    __shared__ int okflag;
    __shared__ double v[1];
    double sum = st[5];
    if (pa.world > 1) {
        const unsigned long long E = dld_volatile_u64(pa.epoch_dev) + 1ull;
        if (threadIdx.x == 0) v[0] = sum;
        __syncthreads();
        peer_publish(pa, E, v, 1);
        const bool ok = peer_wait(pa, E, &okflag);
        sum = ok ? peer_sum(pa, E, 0) : nan_f64();
        __syncthreads();
        if (threadIdx.x == 0) *pa.epoch_dev = E;
    }
    if (threadIdx.x == 0) {
        const double nrm = sqrt(fabs(sum));
        const double rho_new = __dmul_rn(nrm, nrm);
        const double prev = st[1];
        st[0] = prev;
        st[1] = rho_new;
        st[4] = rho_new / prev;
        mailbox[0] = sum;
        mailbox[1] = st[3];
        mailbox[2] = st[2];
    }
}

// out = x + beta_dev[0] * y   (CG direction update p_k = z + beta p_{k-1}, linsys.py:627, beta on the device)
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 4)
xpby_dev_kernel(long long n, const T* __restrict__ x, const double* beta_dev, const T* y, T* out) {
    const double beta = beta_dev[0];
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double xv[VEC], yv[VEC];
        VecIO<T, VEC>::load(x, i, xv);
        VecIO<T, VEC>::loadrw(y, i, yv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) yv[u] = __dadd_rn(xv[u], __dmul_rn(beta, yv[u]));   // numpy: z + (beta*p)
        VecIO<T, VEC>::store(out, i, yv);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x)
            out[i] = (T)__dadd_rn((double)x[i], __dmul_rn(beta, (double)y[i]));
}

