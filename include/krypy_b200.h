/* krypy_b200 -- C ABI of the B200 (sm_100a) Krylov hot path.
 *
 * The reference (andrenarchy/krypy v2.2.0) is pure Python and has no FFI; its
 * plugin surface is the Python operator/solver API.  This
 * header is the boundary a maintainer binds with ctypes/cffi:
 * every entry point replaces the NumPy/SciPy call(s) cited next to it
 * (paths relative to the reference repository root).
 *
 * Conventions
 *  - plain C types only; every pointer named *_dev / V / q / x / y ... is a DEVICE
 *    pointer owned by the caller (e.g. torch.Tensor.data_ptr()); nothing here
 *    allocates caller-visible memory.
 *  - dtype: KRY_F32 / KRY_F64 selects the storage type of the N-sized vectors
 *    and matrix values.  All small quantities (inner products, Hessenberg
 *    entries, Givens rotations, coefficients) are double on both paths.
 *  - a basis is stored VECTOR-MAJOR: vector j starts at V + j*ldv elements.
 *  - all calls are asynchronous on the context's stream; kry_sync() is the only
 *    blocking call.  Results the host needs for convergence control are written
 *    by the kernels into a pinned "mailbox" (kry_mailbox_host) and are valid
 *    after kry_sync().
 *  - return value: 0 on success, negative error class otherwise; the message is
 *    available from kry_last_error() (thread-local).  Never throws.
 */
#ifndef KRYPY_B200_H
#define KRYPY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define KRY_ABI_VERSION 1

#define KRY_OK               0
#define KRY_ERR_ARG         -1
#define KRY_ERR_CUDA        -2
#define KRY_ERR_UNSUPPORTED -3

#define KRY_F32 0
#define KRY_F64 1

#define KRY_MAILBOX_DOUBLES 16384

typedef struct kry_ctx kry_ctx;

/* ---- context ---------------------------------------------------------- */
int         kry_version(void);
const char* kry_last_error(void);
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0 = default */
int         kry_ctx_create(int device, void* stream, kry_ctx** out);
int         kry_ctx_destroy(kry_ctx* ctx);
int         kry_ctx_set_stream(kry_ctx* ctx, void* stream);
/* info[0]=SM count, [1]=10*major+minor, [2]=L2 bytes, [3]=max opt-in smem/block,
 * [4]=cooperative launch supported, [5]=max co-resident CTAs of the fused
 * orthogonalisation kernel (f64) */
int         kry_device_info(kry_ctx* ctx, long long info[8]);
double*     kry_mailbox_host(kry_ctx* ctx);  /* pinned host view, KRY_MAILBOX_DOUBLES */
double*     kry_mailbox_dev(kry_ctx* ctx);   /* device alias of the same memory      */
int         kry_sync(kry_ctx* ctx);          /* cudaStreamSynchronize(stream)        */
long long   kry_launch_count(kry_ctx* ctx);  /* kernels launched through this ctx    */
void        kry_reset_launch_count(kry_ctx* ctx);

/* L2 residency window (cudaStreamAttributeAccessPolicyWindow on the context's stream): kernels launched
 * afterwards keep [base, base + bytes) in the L2 set-aside ("persisting"; the part of the window beyond the
 * set-aside is "streaming").  Used for the vector an Arnoldi step re-reads four times, w = A v_k of
 * krypy/utils.py:968 that utils.py:1012-1045 orthogonalises and normalises: its passes after the first are
 * served by the 126 MB L2 instead of HBM.  base == NULL or bytes <= 0 removes the window.  A pure performance
 * hint: results are unchanged.  info (may be NULL): [0] max set-aside, [1] max window, [2] set-aside in
 * effect, [3] window bytes, [4] hit ratio * 1e6.  KRY_ERR_UNSUPPORTED when the device or the stream does not
 * take the hint (never an error of the solve). */
int         kry_l2_window(kry_ctx* ctx, const void* base, long long bytes, long long info[5]);

/* ---- operators -------------------------------------------------------- */
/* y = A x for CSR A (int32 indices).  Replaces scipy csr_matvec behind
 * krypy/utils.py:1593-1594 (MatrixLinearOperator._dot) as called from
 * utils.py:968 (Arnoldi.advance), linsys.py:631 (Cg), linsys.py:156.
 * Optional fused epilogue (w_dev != NULL): dot_out_dev[0] = sum_i w[i]*y[i]
 * (linsys.py:634, <p,Ap>).  y may be NULL when only the dot is wanted. */
int kry_spmv_csr(kry_ctx* ctx, int dtype, long long nrows, long long ncols, long long nnz,
                 const int* rowptr, const int* colidx, const void* vals,
                 const void* x, void* y, const void* w_dev, double* dot_out_dev);
/* y = A x for CSR A on COMPLEX128 vectors x, y (re, im interleaved), natively: vals is complex128
 * (vals_complex != 0, 20 bytes per entry) or float64 (a real matrix applied to complex vectors, 12 bytes
 * per entry) -- scipy csr_matvec on complex data behind krypy/utils.py:1593-1594.  The real-embedding
 * path moves 48 bytes per entry.  Per row the products are summed in storage order, each product and sum
 * rounded separately. */
int kry_spmv_csr_z(kry_ctx* ctx, int vals_complex, long long nrows, long long ncols, long long nnz,
                   const int* rowptr, const int* colidx, const void* vals, const void* x, void* y);
/* y = A x for a dense row-major m x n matrix (numpy.ndarray.dot,
 * krypy/utils.py:1593-1594; BASELINE config 1 and the reference's N<=100 tests) */
int kry_gemv_dense(kry_ctx* ctx, int dtype, long long m, long long n,
                   const void* A, long long lda, const void* x, void* y);
/* y = d .* x  (a diagonal operator such as the Jacobi M of config C3,
 * applied at linsys.py:661 / utils.py:1031) */
int kry_diag_mul(kry_ctx* ctx, int dtype, long long n, const void* d, const void* x, void* y);

/* ---- N-sized elementwise updates ---------------------------------------- */
/* z = a*x + b*y (y may be NULL if b == 0; z may alias x or y).
 * linsys.py:156 (b - A z), :427 (x0 + Mr yk), :627 (p = z + beta p). */
int kry_axpby(kry_ctx* ctx, int dtype, long long n, double a, const void* x,
              double b, const void* y, void* z);
/* y += sign * coef_dev[0] * x   (utils.py:1007-1009, 1027-1029) */
int kry_axpy_dev(kry_ctx* ctx, int dtype, long long n, const double* coef_dev, double sign,
                 const void* x, void* y);
/* out = mul * x / s_dev[0]   (divide != 0)   or   out = mul * x * s_dev[0]
 * (utils.py:938, 950, 1042-1045; linsys.py:614-618, 669-671) */
int kry_scale_dev(kry_ctx* ctx, int dtype, long long n, const double* s_dev, int divide,
                  double mul, const void* x, void* out);

/* y = i*x for n interleaved complex numbers (2n reals; x != y).  Complex systems are run on
 * the real kernels by real embedding: a complex vector is 2n reals, every basis vector v is
 * stored next to its twin i*v, and <v,q>_C = <v,q>_R + i<iv,q>_R (krypy/utils.py:157). */
int kry_rot90(kry_ctx* ctx, int dtype, long long n, const void* x, void* y);

/* ---- tall-skinny reductions / updates ----------------------------------- */
/* out_dev[j] = sum_i V_j[i]*q[i], j < nv (deterministic two-stage reduction).
 * post: 0 none, 1 out = sqrt(out).  acc_dev (may be NULL): acc_dev[j] += out[j].
 * utils.py:183,191,193 (inner), :226-238 (norm), :540 (Projection._apply). */
int kry_block_dot(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv,
                  const void* q, double* out_dev, int post, double* acc_dev);
/* q += sign * sum_j coef_dev[j] * V_j   (utils.py:549 + :621-624) */
int kry_block_axpy(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv,
                   const double* coef_dev, double sign, void* q);
/* out = x0 + sum_j coef_dev[j] * V_j  (x0 may be NULL)
 * (linsys.py:947-948 Gmres._get_xk; deflation.py:68) */
int kry_block_combine(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv,
                      const double* coef_dev, const void* x0, void* out);

/* ---- fused Gram-Schmidt step (one cooperative kernel) ------------------- */
/* Orthogonalise q against basis vectors j0..nv-1 (krypy/utils.py:996-1045):
 *   algo KRY_ORTH_CGS: per pass c = Vdot^H q ; q -= Vsub c       (block, fused)
 *   algo KRY_ORTH_MGS: per pass, for j: c_j = <Vdot_j,q>; q -= c_j Vsub_j
 *                      (the reference's exact MGS order, utils.py:1012-1029)
 * passes = 1 (mgs/cgs) or 2 (dmgs/cgs2).  h_dev[j] += c_j.
 * pre_vec/pre_coef_dev (may be NULL): q -= pre_coef_dev[0]*pre_vec first
 *   (Lanczos three-term recurrence, utils.py:1000-1009).
 * nrm_dev (may be NULL): nrm_dev[0] = ||q||_2 afterwards (utils.py:1034).
 * vnext (may be NULL; needs nrm_dev): vnext = q / nrm (utils.py:1045), zeros if
 *   nrm == 0.
 * Vdot/Vsub are the V and P bases of utils.py:1015,1027 (equal when M is None). */
#define KRY_ORTH_CGS 0
#define KRY_ORTH_MGS 1
int kry_orth_fused(kry_ctx* ctx, int dtype, long long n, const void* Vdot, const void* Vsub,
                   long long ldv, int j0, int nv, void* q, int passes, int algo,
                   const void* pre_vec, const double* pre_coef_dev,
                   double* h_dev, double* nrm_dev, void* vnext);

/* The same step on COMPLEX128 vectors, natively (csrc/kry_cplx.cu): n complex elements per vector
 * (re, im interleaved, 16-byte aligned), ldv in complex elements, h_dev[2j], h_dev[2j+1] += Re, Im of
 * c_j = <Vdot_j, q> = sum_i conj(Vdot_j[i]) q[i] (the reference's inner product, krypy/utils.py:183
 * X.T.conj() @ Y, as used by Arnoldi.advance utils.py:1012-1029 on complex data -- half of the
 * reference's own test matrix, test/test_linsys.py:118-141), q -= c_j Vsub_j with complex c_j.
 * Every basis vector is read once per sweep (the real-embedding path reads v_j and its twin i v_j).
 * KRY_ORTH_CGS: at most 32 vectors per call.  No pre-subtraction argument: the Lanczos recurrence has
 * real coefficients (utils.py:1003-1009) and runs on the real kernels. */
int kry_orth_fused_z(kry_ctx* ctx, long long n, const void* Vdot, const void* Vsub, long long ldv,
                     int j0, int nv, void* q, int passes, int algo,
                     double* h_dev, double* nrm_dev, void* vnext);

/* Fused Lanczos step for a DIAGONAL inner-product matrix B = diag(bdiag) (BASELINE config C5),
 * krypy/utils.py:1000-1045 with inner(X, Y, ip_B) = X^H (B Y) (utils.py:190-193), ONE cooperative
 * kernel:  q -= pre_coef_dev[0]*vprev (vprev may be NULL: k == 0);  alpha = <vk, q>_B;
 * h3_dev[1] += alpha;  q -= alpha*vk;  h3_dev[2] = beta = sqrt(<q, q>_B);  vnext = q/beta
 * (vnext may be NULL).  h3_dev is the [H[k-1,k], H[k,k], H[k+1,k]] triple kry_minres_recur reads.
 * Default for a real positive diagonal ip_B (host switch KRY_LANCZOS_DIAGB=0: the generic sequence
 * kry_axpy_dev / kry_diag_mul / kry_block_dot / kry_scale_dev). */
int kry_lanczos_diag(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk,
                     const void* bdiag, void* q, const double* pre_coef_dev, double* h3_dev,
                     void* vnext);
/* The same step on a row-partitioned run: both reductions are completed over NVLink peer memory
 * inside the kernel (peer arguments as for kry_peer_allreduce). */
int kry_lanczos_diag_dist(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk,
                          const void* bdiag, void* q, const double* pre_coef_dev, double* h3_dev,
                          void* vnext, int world, int rank, unsigned long long* epoch_dev,
                          double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev);

/* ---- block kernels for the projector set-up (CholQR2) ------------------------------------ */
/* kry_gram: out_dev[i*ky + j] = <X_i, Y_j> (Euclidean, fp64 accumulation) for a block X of kx and a
 * block Y of ky vectors in ONE pass over both (krypy/utils.py:160-193 inner() with m, n > 1, as used
 * by qr / Projection.__init__ / Ritz); kx + ky <= 64 (kx when Y == X), ceil(kx/4)*ceil(ky/4) <= 32.
 * kry_block_trsm: Q = X R^-1 for an upper triangular d x d R_dev (row-major, d <= 32), vector-major
 * blocks, Q may alias X: the second half of a Cholesky-QR round (replaces the column-by-column
 * Gram-Schmidt of krypy/utils.py:698-706 for full-rank blocks). */
int kry_gram(kry_ctx* ctx, int dtype, long long n, const void* X, long long ldx, int kx, const void* Y,
             long long ldy, int ky, double* out_dev);
int kry_block_trsm(kry_ctx* ctx, int dtype, long long n, const void* X, long long ldx, int d,
                   const double* R_dev, void* Q, long long ldq);

/* ---- oblique projection for deflation (one cooperative kernel) ---------- */
/* a <- (I - V (R^-1 Q^H) W^H)^iterations a  (krypy/utils.py:604-627 with
 * :522-552; called from deflation.py:135-143).  W, V: d vectors each; Q, R:
 * d x d row-major device matrices (Q may be NULL: identity transform, the
 * orthogonal-projection case utils.py:510-512).  c_first_dev[0..d) receives the
 * raw W^H a of the first application (Ya = WR^H c, utils.py:542-545). */
int kry_project(kry_ctx* ctx, int dtype, long long n, const void* W, long long ldw,
                const void* V, long long ldv, int d, void* a,
                const double* Q_dev, const double* R_dev, int iterations,
                double* c_first_dev);

/* ---- small (latency-bound) device recurrences --------------------------- */
/* GMRES Hessenberg update (krypy/linsys.py:982-993 + utils.py:405-436, drotg
 * semantics): hcol_dev[0..k+1] is column k of H; applies the k stored
 * rotations cs_dev[2i],cs_dev[2i+1], builds rotation k, rotates y_dev[k..k+1].
 * rcol_dev[0..k+1] receives column k of R; hcol_dev is zeroed afterwards (the
 * orthogonalisation kernels accumulate into it with +=).  Mailbox layout at
 * mailbox[off..]: [ |y[k+1]|, H[0..k+1,k], R[0..k+1,k] ]  (2k+5 doubles). */
int kry_givens_update(kry_ctx* ctx, int k, double* hcol_dev, double* rcol_dev,
                      double* cs_dev, double* y_dev, int mailbox_off);
/* out_dev[0..k) = R[:k,:k]^{-1} y[:k], R row-major with leading dim ldr
 * (scipy.linalg.solve_triangular, linsys.py:946) */
int kry_tri_solve(kry_ctx* ctx, int k, const double* R_dev, long long ldr, const double* y_dev,
                  double* out_dev);
/* the same with R stored column after column (entry (i, j) at Rt_dev[j * ldr + i]): the layout the Givens
 * update leaves on the device when step j's rcol_dev is row j of one array -- the solution update of a
 * restart cycle (linsys.py:941-949) then needs no host copy of R */
int kry_tri_solve_t(kry_ctx* ctx, int k, const double* Rt_dev, long long ldr, const double* y_dev,
                    double* out_dev);
/* Complex twins (complex numbers interleaved re/im in double arrays; krypy/utils.py:419-427:
 * drotg when both entries are real-valued, zrotg otherwise).  hcol_dev, rcol_dev, y_dev hold
 * k+2 complex numbers; cs_dev 4 doubles per rotation [c, flag, s_re, s_im].
 * Mailbox at off: [ |y[k+1]|, H[0..k+1,k] (2(k+2) doubles), R[0..k+1,k] (2(k+2) doubles) ].
 * kry_tri_solve_z: R complex row-major, ldr in complex elements. */
int kry_givens_update_z(kry_ctx* ctx, int k, double* hcol_dev, double* rcol_dev,
                        double* cs_dev, double* y_dev, int mailbox_off);
int kry_tri_solve_z(kry_ctx* ctx, int k, const double* R_dev, long long ldr, const double* y_dev,
                    double* out_dev);
/* MINRES sliding QR (krypy/linsys.py:827-847).  st_dev: 16 doubles of state
 * [G1c,G1s,G1valid,G2c,G2s,G2valid,y0,-, R0,R1,R2,ycoef, ...] (zero-initialised,
 * st[6] = ||r0||); h3_dev: [H[k-1,k], H[k,k], H[k+1,k]] as left by
 * kry_orth_fused (h_dev -> &h3[1]-k, nrm_dev -> &h3[2]); with shift != 0, on return
 * h3[0] = H[k+1,k] and h3[1] = 0 for the next step (utils.py:1003).
 * Mailbox at off: [ |y_next|, R0, R1, R2, ycoef, H[k-1,k], H[k,k], H[k+1,k] ]. */
int kry_minres_recur(kry_ctx* ctx, int k, double* h3_dev, double* st_dev, int shift, int mailbox_off);
/* CG with the scalars of the recurrence resident on the device (krypy/linsys.py:627-665), so that the
 * direction update and the operator apply of iteration k+1 can be enqueued before the host has read
 * iteration k's residual (look-ahead) and, on row-partitioned runs, without a host round trip per
 * reduction.  st_dev: [0] rho_{k-1}  [1] rho_k  [2] <p,Ap>  [3] alpha  [4] beta  [5] local share of
 * the new rho.
 *   kry_cg_update_dev : kry_cg_update with rho = st[1], <p,Ap> = st[2]; writes st[3] = alpha, st[5]
 *   kry_cg_scalars    : new rho (global sum over the peers when world > 1; stored as sqrt(|sum|)^2 like
 *                       the reference, which squares the norm), shift, beta; mailbox[off..off+2] =
 *                       (raw sum, alpha, <p,Ap>)
 *   kry_xpby_dev      : out = x + beta_dev[0]*y   (p_k = z + beta p_{k-1}) */
int kry_cg_update_dev(kry_ctx* ctx, int dtype, long long n, const void* Ap, const void* p, void* yk, void* r,
                      void* z, const void* dinv, double* st_dev);
int kry_cg_scalars(kry_ctx* ctx, double* st_dev, int mailbox_off, int world, int rank,
                   unsigned long long* epoch_dev, double* const* peer_slots_dev,
                   unsigned long long* const* peer_flags_dev);
int kry_xpby_dev(kry_ctx* ctx, int dtype, long long n, const void* x, const double* beta_dev, const void* y,
                 void* out);

/* z = (v - R0*W0 - R1*W1)/R2 ; W0 <- W1 ; W1 <- z ; yk += ycoef*z, scalars from
 * st_dev[8..11] (krypy/linsys.py:844-846). w0/w1 are swapped by the caller. */
int kry_minres_update(kry_ctx* ctx, int dtype, long long n, const void* v, void* w0, const void* w1,
                      void* yk, const double* st_dev);

/* c_out[0..d) = R^{-1} Q^H c_in (Q, R: d x d row-major; krypy/utils.py:547-548), the small
 * transform of the deflation projector when its block dot and block update run as separate
 * kernels (row-partitioned multi-GPU runs). */
int kry_small_qr_apply(kry_ctx* ctx, int d, const double* Q_dev, const double* R_dev,
                       const double* c_in_dev, double* c_out_dev);

/* ---- CG fused update (one streaming kernel) ------------------------------ */
/* krypy/linsys.py:634-665 with a diagonal (Jacobi) or identity M:
 *   alpha = rho / pAp_dev[0];  yk += alpha p;  r -= alpha Ap;
 *   z = dinv .* r (dinv NULL: M is the identity, z is not written);
 *   rho_new = <r, z>.
 * mailbox[off..] = [rho_new, alpha, pAp].  The search-direction update
 * p = z + beta p (linsys.py:627) stays a separate kry_axpby because the host may
 * replace rho by the explicit residual first (linsys.py:681-683). */
int kry_cg_update(kry_ctx* ctx, int dtype, long long n, const void* Ap, const void* p, void* yk,
                  void* r, void* z, const void* dinv, double rho, const double* pAp_dev,
                  int mailbox_off);

/* ---- multi-GPU exchange over NVLink peer memory (one process per GPU) ------- */
/* The reference has no distributed path; these entry points carry the
 * exchange steps of the row-partitioned solve: the remote entries of x that
 * the local rows of A reference (krypy/utils.py:1593-1594 applied to a row block) and the
 * global sums behind utils.inner/norm (utils.py:183, 226).
 * kry_peer_alloc: cudaMalloc'ed, zeroed, IPC-exportable buffer.  kry_ipc_export/open:
 * 64-byte cudaIpcMemHandle_t passed between ranks by the host (e.g. all_gather_object). */
int kry_peer_alloc(kry_ctx* ctx, long long bytes, void** out);
int kry_peer_free(kry_ctx* ctx, void* p);
int kry_ipc_export(kry_ctx* ctx, const void* p, unsigned char handle[64]);
int kry_ipc_open(kry_ctx* ctx, const unsigned char handle[64], void** out);
int kry_ipc_close(kry_ctx* ctx, void* p);
/* dst[i] = peer_bases_dev[halo_peer[i]][elem_offset + halo_off[i]]  (/ div_dev[0] if given):
 * P2P loads of the remote vector entries the local CSR rows need. */
int kry_halo_gather(kry_ctx* ctx, int dtype, long long nhalo, const void* const* peer_bases_dev,
                    long long elem_offset, const int* halo_peer, const int* halo_off,
                    const double* div_dev, void* dst);
/* inout_dev[0..n) <- sum over ranks (n <= 64), deterministic rank-order sum, bitwise identical
 * on every rank.  peer_slots_dev[r]: rank r's slot array (2*world*64 doubles, zeroed);
 * peer_flags_dev[r]: rank r's flag array (world u64, zeroed).  epoch_dev: this rank's operation
 * counter in device memory (zeroed at set-up, incremented by every peer operation, so the
 * launch arguments are replayable from a CUDA graph); all ranks issue the same sequence.  post/acc_dev as in kry_block_dot
 * (applied to the global sum). */
int kry_peer_allreduce(kry_ctx* ctx, int world, int rank, unsigned long long* epoch_dev, int n,
                       double* inout_dev, double* const* peer_slots_dev,
                       unsigned long long* const* peer_flags_dev, int post, double* acc_dev);
int kry_peer_barrier(kry_ctx* ctx, int world, int rank, unsigned long long* epoch_dev,
                     double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev);

/* ---- row-partitioned Gram-Schmidt step with the exchange fused into the kernels ------- */
/* The phases of kry_orth_fused (krypy/utils.py:1012-1045) for a row-partitioned basis; the
 * global sums travel inside the producing/consuming kernels (P2P stores + release flag in the
 * producer's last CTA, acquire + rank-order sum in every CTA of the consumer):
 *   kry_dist_dot    c_local = V^H q (and, want_sq, <q, q> as sum number nv; nv + want_sq <= 64), published
 *                   to all peers
 *   kry_dist_update c = global sum; h_acc_dev[j] += c_j; q -= V c; want_nrm: publishes ||q||^2
 *   kry_dist_scale  nrm = sqrt(global sum) -> nrm_out_dev[0]; vnext = q / nrm (vnext may be NULL)
 *   kry_dist_halo   handshake (every rank's vector is complete) + kry_halo_gather
 * peer arguments as for kry_peer_allreduce. */
/* kry_orth_fused for a row-partitioned basis as ONE cooperative kernel: after each grid-wide
 * reduction CTA 0 publishes the local sums to all peers and every CTA completes the global sum
 * (same arguments as kry_orth_fused + the peer arguments of kry_peer_allreduce; CGS: nv-j0 <= 64). */
int kry_orth_fused_dist(kry_ctx* ctx, int dtype, long long n, const void* Vdot, const void* Vsub,
                        long long ldv, int j0, int nv, void* q, int passes, int algo,
                        const void* pre_vec, const double* pre_coef_dev, double* h_dev, double* nrm_dev,
                        void* vnext, int world, int rank, unsigned long long* epoch_dev,
                        double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev);
int kry_dist_dot(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, const void* q,
                 int want_sq, int world, int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                 unsigned long long* const* peer_flags_dev);
int kry_dist_update(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, void* q,
                    double* h_acc_dev, int want_nrm, int world, int rank, unsigned long long* epoch_dev,
                    double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev);
int kry_dist_scale(kry_ctx* ctx, int dtype, long long n, const void* q, void* vnext, double* nrm_out_dev,
                   int world, int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                   unsigned long long* const* peer_flags_dev);
/* kry_dist_scale + "my segment of vnext is complete" handshake + halo gather of vnext in ONE
 * kernel: the next SpMV on vnext needs no further exchange (halo_dst = vnext + block). */
int kry_dist_scale_halo(kry_ctx* ctx, int dtype, long long n, const void* q, void* vnext,
                        double* nrm_out_dev, long long nhalo, const void* const* peer_bases_dev,
                        long long elem_offset, const int* halo_peer, const int* halo_off, void* halo_dst,
                        int world, int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                        unsigned long long* const* peer_flags_dev);
/* kry_dist_scale_haloq: like kry_dist_scale_halo, but the halo of v_next is gathered from the peers'
 * un-normalised q (element offset q_elem_offset inside the peer-mapped regions) and divided by the norm
 * here, so the norm's flag doubles as the "segment complete" handshake (one cross-GPU wait per call
 * instead of two).  The caller alternates between two q buffers from step to step. */
int kry_dist_scale_haloq(kry_ctx* ctx, int dtype, long long n, const void* q, void* vnext, double* nrm_out_dev,
                         long long nhalo, const void* const* peer_bases_dev, long long q_elem_offset,
                         const int* halo_peer, const int* halo_off, void* halo_dst, int world, int rank,
                         unsigned long long* epoch_dev, double* const* peer_slots_dev,
                         unsigned long long* const* peer_flags_dev);
/* ONE-WAIT Arnoldi step for row-partitioned block classical Gram-Schmidt: kry_spmv_csr, kry_dist_dot with
 * want_sq (publishes the local V^H w and <w, w>, w = A v_k), kry_dist_update_scale -- three kernels, one
 * cross-GPU wait, q written once and never rewritten.
 *   kry_dist_update_scale acquires the peers' partials (the step's only cross-GPU wait), c = rank-order sums,
 *                        nrm^2 = <w,w> - sum c_j^2 (exact norm through a second exchange inside the kernel when
 *                        that difference cancels below 1e-3 <w,w>), vnext = (q - V c) / nrm in one sweep
 *                        (utils.py:1026-1045; q is not modified), h_acc[0..nv) += c, nrm_out[0] = nrm, the halo
 *                        of vnext from the peers' q and the halo entries of V this rank holds at
 *                        V[j * ldv + halo_base + i] (bitwise the owners' values), and -- k_givens >= 0 -- the
 *                        GMRES Givens / Hessenberg update of kry_givens_update (linsys.py:982-993) in one extra
 *                        CTA beside the sweep (requires k_givens + 1 == nv and nrm_out == h_acc + nv).
 *   kry_spmv_csr_mdot    y = A x (utils.py:968) with c[j] = <B[j], y> (j < nb) and c[nb] = <y, y> (want_sq) taken
 *                        in the SpMV epilogue while each row's result is in a register; world > 1: the local sums
 *                        are published like kry_dist_dot does; world == 1: they go to out_dev (the deflation
 *                        projector's <W, A v>, deflation.py:135-143).  Staged short-row path only (<= 15 entries
 *                        per row on average, 16-byte aligned arrays), nb <= 32; otherwise KRY_ERR_UNSUPPORTED.
 *                        Measured slower than SpMV + block dot below 16 vectors (profiles/r2_mdot_kernel.txt):
 *                        not used by the solvers. */
int kry_spmv_csr_mdot(kry_ctx* ctx, int dtype, long long nrows, long long ncols, long long nnz, const int* rowptr,
                      const int* colidx, const void* vals, const void* x, void* y, const void* B, long long ldb,
                      int nb, int want_sq, double* out_dev, int world, int rank, unsigned long long* epoch_dev,
                      double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev);
int kry_dist_update_scale(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, const void* q,
                          void* vnext, double* h_acc_dev, double* nrm_out_dev, long long nhalo,
                          const void* const* peer_q_dev, long long q_elem_offset, const int* halo_peer,
                          const int* halo_off, long long halo_base, void* halo_dst, int k_givens, double* rcol_dev,
                          double* cs_dev, double* y_dev, long long mailbox_off, int world, int rank,
                          unsigned long long* epoch_dev, double* const* peer_slots_dev,
                          unsigned long long* const* peer_flags_dev);
int kry_dist_halo(kry_ctx* ctx, int dtype, long long nhalo, const void* const* peer_bases_dev,
                  long long elem_offset, const int* halo_peer, const int* halo_off, void* dst, int world,
                  int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                  unsigned long long* const* peer_flags_dev);

#ifdef __cplusplus
}
#endif
#endif /* KRYPY_B200_H */
