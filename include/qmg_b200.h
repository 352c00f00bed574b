/* qmg_b200.h -- C ABI of libqmg_b200.so: the sm_100a kernels behind quantum-mg's
 * data-parallel hot path (2D stencil apply, transfer, BLAS-1 reductions).
 *
 * Every vector/matrix argument is a DEVICE pointer to interleaved
 * complex<double> (re,im pairs, 16-byte aligned) in the reference's even-odd
 * layouts (/root/reference/lattice/lattice.h:75-182):
 *   site      i = (y + parity*Y)*X/2 + x/2,  parity = (x+y)%2  (all even, then all odd)
 *   cv        i*nc + c
 *   cm        (i*nc + c1)*nc + c2            (row major)
 *   hopping   mu*size_cm + cm,  mu in {+x,+y,-x,-y}
 *   gauge     mu*V + i,         mu in {x,y}  (nc = 1 lattice)
 * All functions return 0 on success, non-zero on failure; qmg_last_error()
 * gives the message.  One host thread drives one device; work is issued on
 * the stream set by qmg_set_stream (default: the legacy default stream).
 * There is no CPU fallback: without a CUDA device every compute entry fails.
 *
 * Each entry names the reference interface it replaces (paths relative to
 * /root/reference; "qlinalg" = the un-vendored quantum-linalg dependency,
 * whose contract is fixed by the cited call site).
 */
#ifndef QMG_B200_H
#define QMG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double qmg_cplx; /* pointer arithmetic is in doubles: element k is at p[2k], p[2k+1] */

/* ------------------------------------------------------------ runtime ---- */
int qmg_init(int device);                 /* cudaSetDevice + scratch; idempotent */
int qmg_finalize(void);
int qmg_set_stream(void* cuda_stream);    /* cudaStream_t (e.g. torch's current stream) */
void* qmg_get_stream(void);
int qmg_sync(void);                       /* cudaStreamSynchronize on the active stream */
const char* qmg_last_error(void);
int qmg_device_count(void);
int qmg_sm_count(void);
long qmg_kernel_launches(void);           /* kernels launched by this library so far */

/* Per-entry-point timing (also QMG_PROFILE=1): each call is bracketed by stream synchronisations and accumulated by name. */
int qmg_profile_enable(int on);
int qmg_profile_reset(void);
double qmg_profile_report(void);          /* prints the table, returns total seconds */

/* allocate_vector / deallocate_vector (qlinalg; stencil/stencil_2d.h:220,299) */
int qmg_malloc(void** dptr, size_t bytes);
int qmg_free(void* dptr);                 /* parks the block in a size-keyed cache; qmg_trim returns parked blocks to the driver */
int qmg_trim(void);
size_t qmg_cached_bytes(void);
int qmg_memcpy_h2d(void* dst, const void* src, size_t bytes);
int qmg_memcpy_d2h(void* dst, const void* src, size_t bytes);
int qmg_memcpy_d2d(void* dst, const void* src, size_t bytes);
/* Allocation mode of qmg_malloc: 0 = cudaMalloc (default), 1 = cudaMallocManaged with the device as preferred
 * location, so that UNMODIFIED reference drivers, which index vectors from host code
 * (e.g. tests/n11_wilson_test/wilson_test.cpp:96), still run.  Also selectable with the environment
 * variable QMG_MANAGED=1 read at qmg_init. */
int qmg_set_alloc_mode(int managed);
int qmg_get_alloc_mode(void);
int qmg_malloc_host(void** hptr, size_t bytes);   /* pinned host staging, allocated and first touched on the GPU's NUMA node (QMG_NUMA=0: wherever the caller runs) */
int qmg_device_numa_node(void);                   /* NUMA node of the active GPU (sysfs numa_node of its PCI function), -1 if unknown */
int qmg_free_host(void* hptr);

/* --------------------------------------------------------------- sharding -- */
/* y-slab sharding over the GPUs of one node, one process per GPU (the "Becomes MPI" periodic-boundary loops of
 * cshift/cshift_2d.h:72-205).  After qmg_comm_init with nranks > 1 every lattice handed to this library is the LOCAL slab
 * (X, Y/nranks) of rank `rank`, and
 *   - qmg_stencil_apply exchanges one boundary row of rhs with the two ring neighbours over NCCL, overlapped with the
 *     interior rows (when the descriptor's halo pointers are NULL);
 *   - the link fills, variant builders and the coarse build exchange the boundary rows of their inputs once;
 *   - every reduction is all-reduced over the ranks before it is returned;
 *   - prolong / restrict / block-orthonormalise stay local (slab heights must be multiples of the block size).
 * unique_id128: the 128-byte NCCL id from qmg_comm_unique_id on rank 0, handed to every rank by the caller (any out-of-band channel). */
int qmg_comm_unique_id(void* out128);
int qmg_comm_init(int nranks, int rank, const void* unique_id128);
int qmg_comm_finalize(void);
int qmg_comm_size(void);
int qmg_comm_rank(void);
int qmg_comm_active(void);
/* Single-rank loopback: the slab is the whole periodic lattice, but every access across y = 0 / Y-1 goes through the
 * pack / exchange / halo-row path of the sharded build (device copies instead of NCCL).  For single-GPU testing. */
int qmg_comm_set_loopback(int on);
long qmg_comm_halo_exchanges(void);
long qmg_comm_allreduces(void);     /* ncclAllReduce calls so far (0 when the reductions all-reduce in-kernel, see qmg_comm_p2p) */
/* 1 when every reducing kernel all-reduces its result itself through NVLink peer memory (IPC-mapped mailboxes, set up by
 * qmg_comm_init unless QMG_P2P=0 or a peer cannot be mapped), so that no collective is launched per dot / norm */
int qmg_comm_p2p(void);
long qmg_comm_p2p_halo_exchanges(void);   /* halo exchanges done by peer stores into the neighbours' mailboxes (the rest went through ncclSend/Recv) */
int qmg_halo_exchange(const qmg_cplx* field, int X, int Y, int dof, qmg_cplx* out_ym, qmg_cplx* out_yp);

/* ------------------------------------------------------------ stencil ---- */
/* One stored stencil = one pointer set of Stencil2D (stencil/stencil_2d.h:148-210). */
typedef struct qmg_stencil_desc
{
  int X, Y, nc;                /* Lattice2D dims and dof per site (lattice.h:29) */
  const qmg_cplx* clover;      /* V*nc*nc or NULL (stencil_2d.h:155) */
  const qmg_cplx* hopping;     /* 4*V*nc*nc or NULL (stencil_2d.h:158) */
  double shift[2];             /* stencil_2d.h:170 */
  double eo_shift[2];          /* stencil_2d.h:173: + on even, - on odd sites */
  double dof_shift[2];         /* stencil_2d.h:177: + on top half of dof, - on bottom half */
  /* y-slab sharding: rows y=-1 and y=Y of the INPUT vector, as received from
   * the neighbouring ranks, each X*nc elements laid out (parity, x/2, c) with
   * parity that of the halo row's sites; NULL = periodic wrap inside this
   * lattice (cshift/cshift_2d.h:101,114 "Becomes MPI"). */
  const qmg_cplx* halo_ym;     /* row below y=0   */
  const qmg_cplx* halo_yp;     /* row above y=Y-1 */
  /* Opt-in traffic saving for operators with D^dag = gamma5 D gamma5 (Wilson, and its Galerkin coarsenings under
   * chirality-preserving transfers): the stored backward blocks satisfy
   *   hopping_{-mu}(x)[a][b] = s_a s_b conj(hopping_{+mu}(x - mu)[b][a]),  s = +1 / -1 on the top / bottom half of the dof,
   * so the apply reads clover, +x and +y blocks only (3 of 5) and takes the backward hops from the neighbours' forward
   * blocks.  The backward blocks stay stored (the layout contract is unchanged) but are NOT read: set this only after
   * qmg_stencil_gamma5_deviation reports they obey the relation, and clear it when blocks are edited.  nc must be even. */
  int gamma5_hermitian;
  const qmg_cplx* hop_halo_ym; /* gamma5_hermitian on a y-slab: row -1 of hopping_{+y} (X*nc*nc, layout (parity, x/2, nc*nc)) */
  /* Opt-in matrix-free apply for Wilson2D (nc = 2): the stored blocks are the Wilson blocks of this U(1) gauge field
   * (operators/wilson.h:153-209), so the WHOLE-operator apply (pieces = QMG_APPLY_ALL, all directions) reads the links -- 96
   * instead of 384 bytes per site -- and rebuilds the block elements with the arithmetic of qmg_fill_wilson: same bits as the
   * stored-block apply.  Set it only after qmg_wilson_mf_deviation returned 0, and clear it when blocks are edited.  Every other
   * apply (pieces, variants, fused reductions) keeps reading the stored blocks.  NULL: off. */
  const qmg_cplx* wilson_gauge;         /* 2 V complex: [mu V + site], mu = x, y (the nc = 1 lattice of Wilson2D's gauge_links) */
  double wilson_w;                      /* Wilson parameter the blocks were filled with */
  const qmg_cplx* wilson_gauge_halo_ym; /* y-slab: row -1 of U_y (X complex, layout (parity, x/2)), from qmg_halo_exchange */
} qmg_stencil_desc;

/* pieces bitmask (mirrors apply_M_clover/_eo/_oe/_shift, stencil_2d.h:694-909) */
enum { QMG_APPLY_CLOVER = 1, QMG_APPLY_HOP_TO_EVEN = 2 /* apply_M_eo */, QMG_APPLY_HOP_TO_ODD = 4 /* apply_M_oe */,
       QMG_APPLY_SHIFT = 8, QMG_APPLY_ALL = 15,
       QMG_APPLY_IDENTITY_CLOVER = 16 /* rbjacobi: clover is 1 and is not read, stencil_2d.h:1685 */,
       QMG_APPLY_ACCUMULATE = 32 /* lhs += ...  (Stencil2D::apply_M accumulates, stencil_2d.h:912) */,
       QMG_APPLY_EVEN_ROWS_ONLY = 64 /* write only the even half of lhs */,
       QMG_APPLY_ODD_ROWS_ONLY = 128 /* write only the odd half of lhs */ };
/* dir_mask: bit mu selects hopping direction mu (stencil_dir_index, stencil_2d.h:25-31); 15 = all */

/* lhs (=|+=) [clover + shifts] rhs + sum_mu hopping_mu(x) rhs(x+mu).
 * Replaces Stencil2D::apply_M and its pieces (stencil_2d.h:666-936) and the
 * zero_vector + apply of the apply_stencil_2D_* wrappers (:2571-2716). */
int qmg_stencil_apply(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs, const qmg_cplx* rhs);

/* The same apply for vectors in HOST memory (pinned for full overlap; pageable works, slower): rhs is uploaded in row
 * chunks on one copy stream, every chunk of output rows is computed once the rows it reads have arrived and is
 * downloaded on a second copy stream, so both PCIe directions and the SMs are busy together.  Returns when lhs_host is
 * complete.  dev_lhs / dev_rhs: device staging vectors of size_cv elements, or NULL (taken from the block cache);
 * rows_per_chunk <= 0 picks ~64 MB chunks.  This is the entry a driver that keeps the reference's host-resident vectors
 * (tests/n11_wilson_test/wilson_test.cpp:96-104) calls in place of apply_stencil_2D_M. */
int qmg_stencil_apply_host(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs_host, const qmg_cplx* rhs_host,
                           qmg_cplx* dev_lhs, qmg_cplx* dev_rhs, int rows_per_chunk);

/* lhs = b - A rhs in the same single pass (the residual every smoother and K-cycle stage forms as apply + caxpbyz,
 * multigrid/stateful_multigrid.h:863-866,1023-1029): b is read once, A rhs never touches memory.  lhs may alias b, not rhs;
 * QMG_APPLY_ACCUMULATE is not allowed.  Bit-identical to qmg_stencil_apply followed by qmg_caxpbyz(1, b, -1, A rhs, lhs). */
int qmg_stencil_apply_residual(const qmg_stencil_desc* st, int pieces, int dir_mask, qmg_cplx* lhs, const qmg_cplx* rhs, const qmg_cplx* b);

/* Which kernel serves gamma5-hermitian (link-compressed) applies at nc = 8 (also the environment variable QMG_TILE at
 * qmg_init): 0 streaming kernel reading the neighbours' forward blocks through L2; 1 (default) the persistent ring kernel
 * (one CTA per SM, a producer warp filling a two-stage ring of 32-site patches with cp.async.bulk copies on mbarriers, 512
 * consumer threads) on lattices that give every SM at least 8 patches, the one-patch cp.async kernel on smaller ones;
 * 3 always the one-patch cp.async kernel (two threads per (site, column)); 2 the same with one thread per column; 4 the
 * one-patch kernel staged by cp.async.bulk; 5-11 experimental patch / ring shapes (csrc/qmg_stencil.cu launch_tile8).
 * All produce the same result up to the order of the sums. */
int qmg_set_tile_kernel(int mode);
int qmg_get_tile_kernel(void);

/* result2 = { sum |hopping_{-mu}(x) - s s conj(hopping_{+mu}(x-mu))^T|^2 over sites and mu, sum |hopping|^2 }: how far the
 * stored backward blocks are from the gamma5-hermitian relation (see qmg_stencil_desc.gamma5_hermitian). */
int qmg_stencil_gamma5_deviation(const qmg_stencil_desc* st, double* result2);
/* { sum |stored - regenerated|^2, sum |stored|^2 } of an nc = 2 set against the Wilson blocks of st->wilson_gauge / wilson_w
 * (operators/wilson.h:153-209): 0 exactly is the licence for the matrix-free apply (qmg_stencil_desc.wilson_gauge) */
int qmg_wilson_mf_deviation(const qmg_stencil_desc* st, double* result2);

/* Fused apply + reductions for the Krylov updates: out3 = { <lhs|rhs_dot>, |lhs|^2 } after lhs = A rhs.
 * (MR step: alpha = <Ar|r>/<Ar|Ar>, stateful_multigrid.h:860 via qlinalg minres.)
 * dot_with may be NULL (then only the norm is produced).  result: 3 doubles (re, im, norm2). */
int qmg_stencil_apply_dot(const qmg_stencil_desc* st, int pieces, qmg_cplx* lhs, const qmg_cplx* rhs,
                          const qmg_cplx* dot_with, double* result3);

/* Operator fills from U(1) links (gauge: 2*V complex on the nc=1 lattice). */
int qmg_fill_wilson(int X, int Y, double wilson_coeff, const qmg_cplx* gauge, qmg_cplx* clover, qmg_cplx* hopping);      /* operators/wilson.h:153-209 */
int qmg_fill_staggered(int X, int Y, const qmg_cplx* gauge, qmg_cplx* hopping);                                         /* operators/staggered.h:50-72 */
int qmg_fill_laplace(int X, int Y, const qmg_cplx* gauge, qmg_cplx* clover, qmg_cplx* hopping);                         /* operators/gaugedlaplace.h:45-68 */
int qmg_fill_dwf(int X, int Y, int Ls, double wilson_coeff, double mass_re, double mass_im, const qmg_cplx* gauge, qmg_cplx* clover, qmg_cplx* hopping); /* operators/dwf.h:154-255 */

/* Stencil-variant builders. */
int qmg_build_dagger(int X, int Y, int nc, const qmg_cplx* clover, const qmg_cplx* hopping,
                     qmg_cplx* dagger_clover, qmg_cplx* dagger_hopping);                       /* stencil_2d.h:1080-1139 */
int qmg_build_rbjacobi(const qmg_stencil_desc* st, qmg_cplx* cinv, qmg_cplx* rbj_clover, qmg_cplx* rbj_hopping); /* stencil_2d.h:1452-1601 */

/* cshift (cshift/cshift_2d.h:225): lhs(x) = rhs(x + dir) for the source parities in eo (1 even, 2 odd, 3 both). */
int qmg_cshift(qmg_cplx* lhs, const qmg_cplx* rhs, int cdir, int eo, int dof_per_site, int X, int Y);

/* ------------------------------------------------- batched site matrices -- */
int qmg_cmat_xy(const qmg_cplx* M, const qmg_cplx* x, qmg_cplx* y, long nsites, int nc, int accumulate);  /* qlinalg cMATxy / cMATxpy (stencil_2d.h:675,1055) */
int qmg_cmat_single_xy(const qmg_cplx* M, const qmg_cplx* x, qmg_cplx* y, long nsites, int nc);           /* qlinalg cMAT_single_xy (dwf.h:106) */
int qmg_cmat_conjtrans(const qmg_cplx* in, qmg_cplx* out, long nsites, int nc);                           /* qlinalg cMATcopy_conjtrans_square (stencil_2d.h:1097); in == out allowed */
int qmg_cmat_mul(const qmg_cplx* Xm, const qmg_cplx* Ym, qmg_cplx* Zm, long nsites, int nc);                /* qlinalg cMATxtMATyMATz_square (stencil_2d.h:1564) */
int qmg_cmat_inverse(const qmg_cplx* M, qmg_cplx* Minv, long nsites, int nc);                             /* qlinalg cMATx_do_qr_square + cMATqr_do_xinv_square (stencil_2d.h:1536-1537) */
int qmg_cmat_qr(const qmg_cplx* M, qmg_cplx* Q, qmg_cplx* R, long nsites, int nc);                      /* qlinalg cMATx_do_qr_square (stencil_2d.h:1536): M = Q R per site, modified Gram-Schmidt */
int qmg_cmat_qr_inverse(const qmg_cplx* Q, const qmg_cplx* R, qmg_cplx* Minv, long nsites, int nc);     /* qlinalg cMATqr_do_xinv_square (stencil_2d.h:1537): Minv = R^-1 Q^dag */
int qmg_cmat_add_pattern(const double* pattern_host, int len, qmg_cplx* v, long nrepeat);                  /* qlinalg capx_pattern (stencil_2d.h:1526); pattern: len complex on HOST */

/* ------------------------------------------------------------- BLAS-1 ---- */
/* qlinalg blas/generic_vector.h; n counts complex elements; scalars are (re,im). */
int qmg_zero(qmg_cplx* x, long n);                                                     /* zero_vector */
int qmg_copy(qmg_cplx* dst, const qmg_cplx* src, long n);                              /* copy_vector */
int qmg_constant(qmg_cplx* x, double re, double im, long n);                           /* constant_vector */
int qmg_cax(double ar, double ai, qmg_cplx* x, long n);                                /* cax   x *= a */
int qmg_caxy(double ar, double ai, const qmg_cplx* x, qmg_cplx* y, long n);            /* caxy  y = a x */
int qmg_caxpy(double ar, double ai, const qmg_cplx* x, qmg_cplx* y, long n);           /* caxpy y += a x */
int qmg_caxpby(double ar, double ai, const qmg_cplx* x, double br, double bi, qmg_cplx* y, long n);                         /* y = a x + b y */
int qmg_caxpbyz(double ar, double ai, const qmg_cplx* x, double br, double bi, const qmg_cplx* y, qmg_cplx* z, long n);     /* z = a x + b y */
int qmg_caxpbypz(double ar, double ai, const qmg_cplx* x, double br, double bi, const qmg_cplx* y, qmg_cplx* z, long n);    /* z += a x + b y */
int qmg_cxty(const qmg_cplx* x, qmg_cplx* y, long n);                                  /* y *= x elementwise */
int qmg_conj(qmg_cplx* x, long n);                                                     /* conj_vector */
int qmg_cinvx(qmg_cplx* x, long n);                                                    /* cinvx */
int qmg_polar(qmg_cplx* x, long n);                                                    /* polar: x = exp(i Re x) */
/* per-element maps used by the block (bi-)orthonormalisation (transfer/transfer.h:360-376,583,736):
 * op 0: x -> |x|            (abs_vector)
 *    1: x -> 1/sqrt(Re x)   (inv_real_sqrt)
 *    2: x -> 1/|x|          (inv_abs_sqrt)
 *    3: x -> e^{i arg x}/sqrt|x|  (inv_phase_abs_sqrt)
 *    4: x -> arg x          (arg_vector) */
int qmg_elementwise(int op, qmg_cplx* x, long n);
/* strided flavours (the *_blas family, operators/wilson.h:79-125): element k at x[k*stride] */
int qmg_zero_strided(qmg_cplx* x, long stride, long n);
int qmg_constant_strided(qmg_cplx* x, long stride, double re, double im, long n);
int qmg_caxy_strided(double ar, double ai, const qmg_cplx* x, long xs, qmg_cplx* y, long ys, long n, int accumulate); /* caxy_blas / caxpy_blas / copy_vector_blas */
int qmg_cax_strided(double ar, double ai, qmg_cplx* x, long stride, long n);
/* per-site dof pattern: out[s*nc+i] = scale[i]*in[s*nc+shuffle[i]] (caxy_shuffle_pattern, wilson.h:132); scale/shuffle on HOST */
int qmg_shuffle_pattern(const double* scale_host, const int* shuffle_host, int nc, const qmg_cplx* in, qmg_cplx* out, long nsites);

/* reductions: result written to HOST (synchronises the stream) */
int qmg_dot(const qmg_cplx* x, const qmg_cplx* y, long n, double* result2);            /* dot: sum conj(x) y */
int qmg_norm2sq(const qmg_cplx* x, long n, double* result);                            /* norm2sq */
int qmg_diffnorm2sq(const qmg_cplx* x, const qmg_cplx* y, long n, double* result);     /* diffnorm2sq */
int qmg_norminf(const qmg_cplx* x, long n, double* result);                            /* norminf */
/* <x|y> and |x|^2 in one pass: result3 = { Re<x|y>, Im<x|y>, |x|^2 }  (MR / GCR: alpha = <Ap|r>/<Ap|Ap>) */
int qmg_dot_norm(const qmg_cplx* x, const qmg_cplx* y, long n, double* result3);
/* k dot products against one vector in one pass: result[2j..] = <xs[j]|y> (GCR orthogonalisation). xs: HOST array of k device pointers. */
int qmg_multi_dot(const qmg_cplx* const* xs_host, int k, const qmg_cplx* y, long n, double* result2k);
/* fused Krylov updates */
/* x += a p ; r -= a q ; result = |r|^2   (MR / GCR step) */
int qmg_update_xr_norm(double ar, double ai, const qmg_cplx* p, const qmg_cplx* q, qmg_cplx* x, qmg_cplx* r, long n, double* result);
/* One MR / GCR step with a single host wait: alpha = omega <q|r>/<q|q> formed on the device, x += alpha p, r -= alpha q;
 * result4 = { |r|^2, Re<q|r>, Im<q|r>, <q|q> }.  Bit-identical to qmg_dot_norm(q, r) + qmg_update_xr_norm(alpha, p, q, x, r). */
int qmg_step_xr_norm(double omega, const qmg_cplx* p, const qmg_cplx* q, qmg_cplx* x, qmg_cplx* r, long n, double* result4);
/* The general MR / GCR step (qmg_step_xr_norm with the first and last steps of a solve folded in):
 *   alpha = omega <q|r_in>/<q|q> ;  t = (x_in ? x_in : 0) + alpha p ;  x_out = acc ? acc + t : t ;
 *   r_out = r_in - alpha q ;  |r_out|^2.
 * x_in NULL: first step from a zero start (x_out is written, never read).  r_in != r_out: the same first step reading the
 * right-hand side itself.  acc: folds the "lhs += z" after a smoother (stateful_multigrid.h:1050) into its last step.
 * flags: QMG_STEP_WANT_RNORM also returns |r_in|^2 (read anyway); QMG_STEP_X_ONLY skips r_out and every reduction
 * read-back (the last step of a smoother whose residual nobody reads) -- result5 is then untouched.
 * result5 = { |r_out|^2, Re<q|r_in>, Im<q|r_in>, <q|q>, |r_in|^2 }.  Without flags bit-identical to qmg_step_xr_norm.
 * QMG_STEP_DOTS_READY: <q|r_in> and <q|q> were left on the device by qmg_gcr_orthogonalize -- the dot pass is skipped.
 * qq_dev (device, may be NULL): receives <q|q>, the |Ap_k|^2 later GCR orthogonalisations divide by. */
enum { QMG_STEP_WANT_RNORM = 1, QMG_STEP_X_ONLY = 2, QMG_STEP_DOTS_READY = 4,
       QMG_STEP_R_ONLY = 8 /* r_out = r_in - alpha q and |r_out|^2 only: p, x_in, x_out, acc are not touched (GCR forming x at the end) */,
       QMG_STEP_NO_NORM = 16 /* x_out and r_out as without flags, but |r_out|^2 is not formed: no reduction, no host wait, nothing returned
                                (a smoother's last step whose residual the caller takes over but whose norm nobody reads) */ };
int qmg_krylov_step(double omega, const qmg_cplx* p, const qmg_cplx* q, const qmg_cplx* x_in, qmg_cplx* x_out,
                    const qmg_cplx* r_in, qmg_cplx* r_out, const qmg_cplx* acc, long n, int flags, double* result5, double* qq_dev);
/* GCR: orthogonalise the new direction against the k stored ones with the coefficients formed ON THE DEVICE and prepare the
 * step, no host round trip (quantum-linalg's minv_vector_gcr / _gcr_var_precond loop body between the operator apply and
 * the x / r update):  dots_dev[2j..] = <Ap[j]|Ap_k> ;  beta_j = -dots_j / apn_dev[j] ;  Ap_k += sum beta_j Ap[j] ;
 * p_k = dir + sum beta_j p[j] (dir == p_k allowed) ;  { <Ap_k|r>, |Ap_k|^2 } stay on the device for qmg_krylov_step with
 * QMG_STEP_DOTS_READY.  Ap_host / p_host: HOST arrays of k device pointers; dots_dev: 2k doubles of device scratch;
 * apn_dev: k doubles on the device (filled by qmg_krylov_step's qq_dev).  k >= 1.  Same sums as qmg_multi_dot + host
 * division + 2 qmg_multi_axpyz + the dot pass of the step.  p_host NULL: only the A p basis is orthogonalised (dir, pk
 * unused) -- the solver keeps the raw directions and forms x from them once, at the end of the solve. */
int qmg_gcr_orthogonalize(const qmg_cplx* const* Ap_host, const qmg_cplx* const* p_host, int k, qmg_cplx* Apk, const qmg_cplx* dir, qmg_cplx* pk,
                          const qmg_cplx* r, long n, double* dots_dev, const double* apn_dev);
/* y += sum_j a_j xs[j]   (GCR: p_k += sum beta_i p_i) ; a: HOST 2k doubles; xs: HOST array of k device pointers */
int qmg_multi_axpy(const double* a_host, const qmg_cplx* const* xs_host, int k, qmg_cplx* y, long n);
/* y = x0 + sum_j a_j xs[j]   (GCR: p_k = r + sum beta_i p_i without a separate copy); x0 == y allowed */
int qmg_multi_axpyz(const double* a_host, const qmg_cplx* const* xs_host, int k, const qmg_cplx* x0, qmg_cplx* y, long n);
/* Two MR steps from a zero start (the K-cycle's smoother, /root/reference/multigrid/stateful_multigrid.h:860,1046) in two vector
 * passes: with q1 = A r0 and p2 = A q1 (instead of A r1 = q1 - a1 p2) both step lengths follow from one pass of dot products,
 * out9 = { <q1|r0> (re, im), <q1|q1>, <p2|r0> (re, im), <p2|q1> (re, im), <p2|p2>, <r0|r0> }, and x, r2 from one update:
 * x_out = (acc ? acc : 0) + cx0 r0 + cx1 q1 ; r_out = r0 + cr1 q1 + cr2 p2 (r_out NULL: not formed; may alias r0). */
int qmg_mr2_gram(const qmg_cplx* r0, const qmg_cplx* q1, const qmg_cplx* p2, long n, double* out9);
int qmg_mr2_update(const double* cx0, const double* cx1, const double* cr1, const double* cr2, const qmg_cplx* r0, const qmg_cplx* q1, const qmg_cplx* p2,
                   const qmg_cplx* acc, qmg_cplx* x_out, qmg_cplx* r_out, long n);
/* BiCGstab(L) sweeps (quantum-linalg minv_vector_bicgstab_l as the K-cycle set-up calls it,
 * /root/reference/tests/n13_wilson_kcycle/wilson_kcycle.cpp:359) with the BLAS-1 traffic of a sweep cut from 280 to 148 vector
 * passes; every element sees the floating-point operations of the call-by-call sequence in the same order.  L <= qmg_bicgstab_max_l().
 * replay: the updates of the LOWER vectors of the BiCG part (u_i = r_i - beta_j u_i, r_i -= alpha_j u_{i+1}, i < j; x += alpha_j u_0)
 *   for all L steps in one pass, after the solver did the top ones (i = j) step by step.  r_host / u_host: HOST arrays of L device
 *   pointers r_0..r_{L-1}, u_0..u_{L-1}; alpha_host / beta_host: 2 L doubles.
 * mgs: modified Gram-Schmidt of r_1..r_L in place (right-looking order, coefficients formed on the device, no host wait between
 *   the L passes); sums_host: L rows of 2 L + 2 doubles, row i-1 = { |r_i|^2, <r_i|r_0>, <r_i|r_{i+1}>, ..., <r_i|r_L> } (complex as re, im).
 * finish: x += sum_{j<L} cx_j r_j ; r_0 += sum_{j=1..L} cr_j r_j ; result = |r_0|^2 -- one pass over r_0..r_L (r_host: L + 1 pointers). */
int qmg_bicgstab_max_l(void);
int qmg_set_bicgstab_fused(int on);   /* 0: the call-by-call sequence (default 1; QMG_BICGSTAB_FUSED=0 at qmg_init) */
int qmg_get_bicgstab_fused(void);
int qmg_bicgstab_replay(int L, qmg_cplx* const* r_host, qmg_cplx* const* u_host, qmg_cplx* x, const double* alpha_host, const double* beta_host, long n);
int qmg_bicgstab_mgs(int L, qmg_cplx* const* r_host, long n, double* sums_host);
int qmg_bicgstab_finish(int L, qmg_cplx* const* r_host, qmg_cplx* x, const double* cx_host, const double* cr_host, long n, double* result);
/* gaussian fill, counter-based (Philox) so results do not depend on the launch shape; when sharded the counter is the
 * GLOBAL element index of an even-odd field (the ranks together draw what one GPU draws for the whole lattice) */
int qmg_gaussian(qmg_cplx* x, long n, uint64_t seed, uint64_t stream_id, double dev);

/* per-row (time-slice) reductions of colour vectors: op 0 norm2sq_cv_timeslice(a), 1 redot_cv_timeslice(a, b),
 * 2 dot_cv_timeslice(a, b) (reductions/reductions.h:24-92); host_out: Y doubles (op 2: 2 Y, interleaved re, im) */
int qmg_timeslice_reduce(int op, const qmg_cplx* a, const qmg_cplx* b, int X, int Y, int nc, double* host_out);

/* ------------------------------------------------------- U(1) gauge fields -- */
/* u1/u1_utils.h on the nc = 1 lattice: gauge = 2 V complex links, phases = 2 V real angles, both [mu * V + site]. */
int qmg_zero_bytes(void* dptr, size_t bytes);                                               /* zero_vector on a double field (tests/n13_wilson_kcycle/wilson_kcycle.cpp:203) */
int qmg_polar_vector(const double* phases, qmg_cplx* gauge, long n);                       /* qlinalg polar_vector (n13 :212) */
/* result4 = { Re <plaq>, Im <plaq>, topological charge, 0 } in one pass (get_plaquette_u1 / get_topo_u1, u1_utils.h:424-508) */
int qmg_u1_plaquette(const qmg_cplx* gauge, int X, int Y, double* result4);
int qmg_u1_noncompact_action(const double* phases, int X, int Y, double beta, double* result);    /* u1_utils.h:386-421 */
int qmg_u1_gauge_transform(qmg_cplx* gauge, const qmg_cplx* trans, int X, int Y);          /* apply_gauge_trans_u1, u1_utils.h:241-272 */
/* apply_ape_smear_u1, u1_utils.h:276-383.  textbook = 0: bit-for-bit what the reference computes (its y staples land on the
 * x links, :352,:372); textbook = 1: each link smeared with its own two staples */
int qmg_u1_ape_smear(qmg_cplx* smeared, const qmg_cplx* gauge, int X, int Y, double alpha, int n_iter, int textbook);
/* heatbath_noncompact_update (u1_utils.h:607-667) as four independent subsets per update (the reference sweeps serially
 * and notes "We would need subsets"); counter-based RNG keyed by (seed, global link, update0 + update) */
int qmg_u1_heatbath(double* phases, int X, int Y, double beta, int n_update, uint64_t seed, uint64_t update0);

/* ------------------------------------------------------------ transfer ---- */
/* Regular non-overlapping blocking of a fine (Xf,Yf,ncf) lattice onto a coarse
 * (Xc,Yc) lattice with ncc dof per coarse site (transfer/transfer.h:118-140).
 * null vectors: HOST array of nvec DEVICE pointers, each a fine cv vector.
 * nvec may be smaller than ncc (the 1-vector calls of block_orthonormalize,
 * transfer.h:540-602, always address coarse dof 0.. of an ncc-strided vector). */
typedef struct qmg_transfer_desc
{
  int Xf, Yf, ncf;
  int Xc, Yc, ncc;
} qmg_transfer_desc;
int qmg_prolong(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec,
                const qmg_cplx* coarse, qmg_cplx* fine);     /* fine += P coarse  (transfer.h:455-480) */
int qmg_restrict(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec,
                 const qmg_cplx* fine, qmg_cplx* coarse);    /* coarse += P^dag fine (transfer.h:487-511) */
/* coarse = P^dag fine, written outright (zero_vector + restrict_f2c, stateful_multigrid.h:876-878, in one pass); nvec == ncc */
int qmg_restrict_overwrite(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec,
                           const qmg_cplx* fine, qmg_cplx* coarse);
/* fine_out = base + P coarse (base NULL: P coarse): zero_vector + prolong_c2f + cxpyz (stateful_multigrid.h:1005-1019) in
 * one pass, the sum formed from zero first so the result is bit-identical; nvec <= 8; fine_out may alias base */
int qmg_prolong_add(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec,
                    const qmg_cplx* coarse, const qmg_cplx* base, qmg_cplx* fine_out);
/* Chirality-packed null vectors (B200 extension).  With QMG_DOUBLE_PROJECTION (every K-cycle of the reference) vector j holds the
 * upper-chirality, vector j + ncc/2 the lower-chirality components of one solve: at a fine element of chirality h = (c >= ncf/2)
 * only vectors [h ncc/2, (h+1) ncc/2) are non-zero.  pack_chiral writes packed[idx (ncc/2) + i] = nv[h ncc/2 + i][idx] (N_f ncc/2
 * complex) and returns the sum of |nv|^2 over the entries it DROPS: the packed restrict / prolong (80 / 96 instead of 144 / 160 bytes
 * per fine dof) may replace qmg_restrict / qmg_prolong(_add) only when that is exactly 0.  Same sums up to the order of the
 * cross-lane additions of the restriction. */
int qmg_transfer_packed_supported(const qmg_transfer_desc* t);
int qmg_transfer_pack_chiral(const qmg_transfer_desc* t, const qmg_cplx* const* nullvecs_host, int nvec, qmg_cplx* packed, double* dropped_norm2);
int qmg_restrict_packed(const qmg_transfer_desc* t, const qmg_cplx* packed, const qmg_cplx* fine, qmg_cplx* coarse, int overwrite);
int qmg_prolong_packed(const qmg_transfer_desc* t, const qmg_cplx* packed, const qmg_cplx* coarse, const qmg_cplx* base, qmg_cplx* fine_out, int use_base);
/* one pass of per-aggregate Gram-Schmidt, in place (transfer.h:514-607);
 * cholesky: optional V_c*ncc*ncc output of the triangular factor (:555-594) or NULL */
int qmg_block_orthonormalize(const qmg_transfer_desc* t, qmg_cplx* const* nullvecs_host, int nvec, qmg_cplx* cholesky);

/* ------------------------------------------------------- coarse operator -- */
/* Galerkin coarse stencil  R A P  of a nearest-neighbour fine stencil
 * (operators/coarse.h:90-471): clover_c (V_c*ncc^2) and hopping_c (4*V_c*ncc^2), overwritten. */
int qmg_coarse_build(const qmg_transfer_desc* t, const qmg_stencil_desc* fine,
                     const qmg_cplx* const* prolong_vecs_host, const qmg_cplx* const* restrict_vecs_host,
                     qmg_cplx* clover_c, qmg_cplx* hopping_c);

#ifdef __cplusplus
}
#endif
#endif /* QMG_B200_H */
