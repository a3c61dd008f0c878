/*
 * phb200 — B200-native (sm_100a) sparse linear-solve path for PhiML's torch backend.  C ABI.
 *
 * This header is the drop-in boundary.  Every entry point takes plain device pointers, sizes and a
 * cudaStream_t (as void*); no torch types.  All functions return 0 on success or a negative PHB200_ERR_* code
 * (message via phb200_last_error()); they never throw and never allocate user-visible memory: vectors and
 * factor storage are supplied by the caller after a *_bytes query.  Opaque handles own only small internal
 * bookkeeping (reduction scratch, per-system scalars, level tables).
 *
 * Reference interfaces replaced (paths relative to the PhiML 1.15.1 tree):
 *   phb200_spmv / phb200_spmv_coo      Backend.mul_matrix_batched_vector  phiml/backend/torch/_torch_backend.py:541-547
 *                                      Backend.mul_csr_dense              phiml/backend/torch/_torch_backend.py:912-922
 *                                      Backend.mul_coo_dense              phiml/backend/_backend.py:1303-1328
 *   phb200_matrix_gradient_*           implicit_gradient_solve (dm)       phiml/math/_optimize.py:785-817
 *   phb200_matrix_csr                  TorchBackend.csr_matrix/csc_matrix phiml/backend/torch/_torch_backend.py:869-896
 *   phb200_stencil_* / _assemble_*     tracer_to_coo + compress           phiml/math/_lin_trace.py:750-807, phiml/math/_sparse.py:347-390
 *   phb200_solver_* (PHB200_CG)        cg + stop_on_l2                    phiml/backend/_linalg.py:23-90
 *                 (PHB200_CG_ADAPTIVE) cg_adaptive                        phiml/backend/_linalg.py:93-128
 *                 (PHB200_CG_TORCH, PHB200_CG_ADAPTIVE_TORCH)
 *                                      torch_sparse_cg(_adaptive)         phiml/backend/torch/_torch_backend.py:932-963,1353-1413
 *                 (PHB200_BICGSTAB)    bicg / bicg_stab_first_order       phiml/backend/_linalg.py:131-274
 *                                      Backend.while_loop (device-side)   phiml/backend/_backend.py:697-735
 *   phb200_jacobi_setup                NEW ('jacobi' is absent from compute_preconditioner, phiml/math/_optimize.py:836-880)
 *   phb200_precond_ilu_sweeps          incomplete_lu_coo                  phiml/backend/_linalg.py:524-594
 *   phb200_precond_ilu0                its fixed point (exact ILU(0))
 *   phb200_precond_cluster             ExplicitClusterSolve.apply          phiml/backend/_linalg.py:734-766
 *   phb200_sptrsv                      Backend.solve_triangular_sparse    phiml/backend/_backend.py:1580-1586, _numpy_backend.py:579-580
 */
#ifndef PHB200_H
#define PHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHB200_VERSION 100

/* arithmetic type of values and vectors */
#define PHB200_F32 0
#define PHB200_F64 1

/* status codes */
#define PHB200_OK 0
#define PHB200_ERR_ARG (-1)
#define PHB200_ERR_CUDA (-2)
#define PHB200_ERR_UNSUPPORTED (-3)
#define PHB200_ERR_WORKSPACE (-4)
#define PHB200_ERR_NOT_STENCIL (-5)

/* solver methods (Backend.linear_solve method strings -> codes, phiml/backend/_backend.py:1499-1516) */
#define PHB200_CG 0                /* 'CG' generic loop (_linalg.cg), any preconditioner                     */
#define PHB200_CG_ADAPTIVE 1       /* 'CG-adaptive' / 'auto' generic loop (_linalg.cg_adaptive)              */
#define PHB200_CG_TORCH 2          /* 'CG' torch fast path: |y|^2 tolerance, refresh every 20, div>100      */
#define PHB200_CG_ADAPTIVE_TORCH 3 /* 'auto' torch fast path                                                  */
#define PHB200_BICGSTAB 4          /* 'biCG-stab(L)', poly_order = L >= 1                                     */

/* preconditioner kinds */
#define PHB200_PRE_NONE 0
#define PHB200_PRE_JACOBI 1
#define PHB200_PRE_ILU 2
#define PHB200_PRE_CLUSTER 3

/* matrix kinds */
#define PHB200_MAT_CSR 0
#define PHB200_MAT_STENCIL 1
#define PHB200_MAT_CODED 2

typedef struct phb200_matrix phb200_matrix;
typedef struct phb200_precond phb200_precond;
typedef struct phb200_solver phb200_solver;

int phb200_version(void);
const char* phb200_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches evidence) */
int64_t phb200_launch_count(void);

/* ---------------------------------------------------------------- matrices ------------------------------- */

/* Non-owning CSR view (int32 row pointers / column indices as produced by SparseCoordinateTensor.compress).
 * `transposed != 0` declares that the arrays describe A^T, i.e. a CSC native of A (backward pass,
 * phiml/math/_optimize.py:824-833); products then scatter instead of gather. */
int phb200_matrix_csr(phb200_matrix** out, const int32_t* rowptr, const int32_t* col, const void* val,
                      int64_t n_rows, int64_t n_cols, int64_t nnz, int dtype, int transposed);

/* Constant-coefficient stencil on a C-ordered grid (rank 1..3, slowest dim first) with ZERO boundaries
 * (out-of-domain neighbours dropped) — what matrix_from_function emits for laplace / central differences.
 * offsets: npts*rank ints (source = output + offset), coeffs: npts doubles (rounded to dtype). npts <= 16. */
int phb200_matrix_stencil(phb200_matrix** out, int rank, const int64_t* grid, int npts,
                          const int32_t* offsets, const double* coeffs, int dtype);

/* Device-side check whether a CSR matrix is exactly such a stencil on `grid` (columns AND values bit-equal for
 * every row).  On success *out is a new stencil handle; PHB200_ERR_NOT_STENCIL otherwise. Synchronises stream. */
int phb200_stencil_from_csr(phb200_matrix** out, const phb200_matrix* csr, int rank, const int64_t* grid, void* stream);

/* Stencil -> canonical CSR on the device (sorted columns, int32), byte-identical to the reference's compress().
 * phb200_stencil_nnz gives the array sizes. */
int phb200_stencil_nnz(const phb200_matrix* stencil, int64_t* nnz);
int phb200_stencil_to_csr(const phb200_matrix* stencil, int32_t* rowptr, int32_t* col, void* val, void* stream);

/* General device assembly of a traced operator (phiml/math/_lin_trace.py:750-807 + SparseCoordinateTensor.compress, _sparse.py:347-390): a
 * ShiftLinTracer is `nshifts` (<= 16) shifts (rank ints each, source = output + shift, wrapped periodically) with ONE value per grid cell and
 * shift — values: device array [nshifts][N] in C order, N = prod(grid).  Entries with value 0 are dropped (np.nonzero, :774), rows are ordered by
 * column: the result is byte-identical to the reference's compress_rows() for variable coefficients, PERIODIC wrap and ZERO boundaries alike.
 * phb200_shift_count fills rowptr (N + 1) and *nnz (synchronises stream; PHB200_ERR_ARG if nnz does not fit int32), phb200_shift_fill writes col / val
 * (nnz each; PHB200_ERR_UNSUPPORTED if two shifts wrap onto one column — the reference asserts there, _sparse.py:378). */
int phb200_shift_count(int rank, const int64_t* grid, int nshifts, const int32_t* shifts, const void* values, int dtype, int32_t* rowptr, int64_t* nnz, void* stream);
int phb200_shift_fill(int rank, const int64_t* grid, int nshifts, const int32_t* shifts, const void* values, int dtype, const int32_t* rowptr, int32_t* col, void* val, void* stream);

/* Pattern-coded form of a CSR matrix whose rows fall into few distinct (offset, value) patterns — what matrix_from_function emits for
 * stencils with ZERO_GRADIENT / PERIODIC boundaries, masks or piecewise-constant coefficients (phiml/math/_lin_trace.py:750-807: one
 * value array per shift, wrapped columns).  The device recognises <= 16 distinct flat offsets (col - row) and <= 256 distinct rows, verifies
 * every row bit for bit and stores ONE BYTE per row plus a table of the distinct rows; products (phb200_spmv, every Krylov loop) then move
 * N + 2 N w bytes instead of 8 nnz + 4 (N+1) + 2 N w and are bit-identical to the CSR product (ascending-column order, separately rounded
 * products).  *out views the CSR arrays as well (they must stay alive) and owns its code / table arrays.  PHB200_ERR_NOT_STENCIL when the
 * matrix does not qualify (unsorted columns, too many offsets or patterns, non-square).  Synchronises stream. */
int phb200_coded_from_csr(phb200_matrix** out, const phb200_matrix* csr, void* stream);
/* number of offsets / patterns of a coded matrix; offsets (may be NULL) receives up to 16 flat offsets in ascending order; *star_diag = 1 if the
 * matrix is a STAR on a grid (centre + nearest neighbours per axis, rows of 16-byte aligned length) whose rows differ only at the boundary: a
 * neighbour outside the grid is cut off (ZERO / ZERO_GRADIENT-type boundaries; the diagonal may differ per row) or, on a PERIODIC axis, sits at the
 * other end of the axis — verified for every row.  Such matrices run the TMA-staged stencil product (diagonal looked up per row, wrap boxes for
 * periodic axes) and, without periodic axes, the fused CG / CG-adaptive kernels. */
int phb200_coded_info(const phb200_matrix* coded, int* n_offsets, int* n_patterns, int64_t* offsets, int* star_diag);

/* Slab (row-partition) mode for multi-GPU solves of ONE system: the arrays span `grid[0]` planes (owned planes plus one
 * halo plane per neighbour), but only planes [x_begin, x_end) are produced / updated / reduced over.  Halo planes are
 * read like ordinary neighbours; the caller keeps them current (phiml_b200/multi_gpu.py exchanges them over NCCL). */
int phb200_matrix_set_slab(phb200_matrix* m, int64_t x_begin, int64_t x_end);

int phb200_matrix_info(const phb200_matrix* m, int* kind, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, int* dtype);
int phb200_matrix_destroy(phb200_matrix* m);

/* ---------------------------------------------------------------- products ------------------------------- */

/* Y[b,:] = A X[b,:] for b < batch; X (batch, ldx>=n_cols), Y (batch, ldy>=n_rows), row-major. */
int phb200_spmv(const phb200_matrix* A, const void* X, void* Y, int64_t batch, int64_t ldx, int64_t ldy, void* stream);

/* COO product with int64 indices as torch stores them (sparse_coo_tensor, _torch_backend.py:857-867);
 * Y must be zero-filled by the caller; accumulation by atomics (order not deterministic). */
int phb200_spmv_coo(const int64_t* row, const int64_t* col, const void* val, int64_t nnz, int64_t n_rows, int64_t n_cols,
                    const void* X, void* Y, int64_t batch, int64_t ldx, int64_t ldy, int dtype, void* stream);

/* Batched CSR x dense product outside the solver — replaces Backend.mul_csr_dense (phiml/backend/_backend.py:1366-1387; torch
 * _torch_backend.py:912-922 / numpy _numpy_backend.py:510-519: a Python loop over (batch, channel) pairs with one native matrix each;
 * called from dot_compressed_dense, phiml/math/_sparse.py:1590) with ONE launch.  col (batch, nnz) int32, rowptr (batch, n_rows+1) int32,
 * val (batch, nnz, channels), dense (batch, n_cols, channels, dense_cols) -> out (batch, n_rows, channels, dense_cols), all C-contiguous.
 * Every output is summed over its row's entries in entry order with separately rounded products (SciPy's csr_matvecs, bit for bit). */
int phb200_csr_dense_batched(const int32_t* col, const int32_t* rowptr, const void* val, const void* dense, void* out, int64_t batch, int64_t n_rows,
                             int64_t n_cols, int64_t nnz, int64_t channels, int64_t dense_cols, int dtype, void* stream);

/* CSR of A^T from the CSR of A (= reading the arrays of a CSC native, TorchBackend.csc_matrix _torch_backend.py:885-896, the layout the
 * backward solve receives, phiml/math/_optimize.py:823-833, as the plain CSR the solver kernels take).  Rows of the result are
 * sorted by column; values are moved, not changed.  t_rowptr: n_cols+1, t_col / t_val: nnz.  Temporaries (12 bytes per entry +
 * sort scratch) come from the stream-ordered pool and are released before returning. */
int phb200_csr_transpose(const int32_t* rowptr, const int32_t* col, const void* val, int64_t n_rows, int64_t n_cols, int64_t nnz, int dtype,
                         int32_t* t_rowptr, int32_t* t_col, void* t_val, void* stream);

/* Matrix gradient of the implicit-gradient (adjoint) step, phiml/math/_optimize.py:803-807 (`dm_values = dy[row] * x[col]`,
 * negated and summed over the batch): out[k] = -sum_b DY[b,row[k]] * X[b,col[k]] for every stored entry k, in entry order.
 * DY = solution of the adjoint solve A^T dy = dx (phb200_solver_* on the transposed matrix), X = forward solution.
 * COO form: int64 coordinates as torch stores them; CSR form: int32 pointers / columns (entry order = CSR order, which is
 * what CompressedSparseMatrix.decompress() yields, _optimize.py:800-801). */
int phb200_matrix_gradient_coo(const int64_t* row, const int64_t* col, int64_t nnz, int64_t n_rows, int64_t n_cols, const void* DY, const void* X,
                               int64_t batch, int64_t ld_dy, int64_t ld_x, void* out, int dtype, void* stream);
int phb200_matrix_gradient_csr(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols, int64_t nnz, const void* DY, const void* X,
                               int64_t batch, int64_t ld_dy, int64_t ld_x, void* out, int dtype, void* stream);

/* ------------------------------------------------------------- preconditioners --------------------------- */

/* inv_diag[i] = 1 / A[i,i]  (caller-owned, n_rows elements of A's dtype). */
int phb200_jacobi_setup(const phb200_matrix* A, void* inv_diag, void* stream);
/* `constant` != 0 promises that all entries of inv_diag are equal (constant-diagonal stencils): the fused stencil
 * kernels then use the scalar and never read the vector. */
int phb200_precond_jacobi(phb200_precond** out, const void* inv_diag, int64_t n, int dtype, int constant);

/* Exact ILU(0) on the pattern of a CSR matrix with sorted columns and a full diagonal.
 * lu_val (nnz elements, caller-owned) receives L (strict lower, unit diagonal implied) and U (upper incl. diagonal)
 * in A's own pattern.  The handle keeps the dependency-level schedule used by factorisation and both solves; both run
 * sync-free in ONE launch each (rows wait on per-row completion flags; PHB200_TRSV=levels selects one launch per level).
 * If A was recognised as a star stencil, pass the stencil handle: the levels are then computed analytically on the
 * device instead of from a host copy of the pattern. */
int phb200_precond_ilu0(phb200_precond** out, const phb200_matrix* A, void* lu_val, const phb200_matrix* stencil_twin /* may be NULL */, void* stream);
/* The reference's own ILU: `sweeps` Jacobi-style fixed-point sweeps on the stored pattern (incomplete_lu_coo,
 * phiml/backend/_linalg.py:524-594; compute_preconditioner picks sweeps = ceil(sqrt(d n^(1/d))), d = (nnz/n - 1)/2,
 * phiml/math/_optimize.py:855-869), i.e. the factors behind the reference's 'ilu (iter=k)'.  Every entry is an independent
 * gather + multiply-add over the previous iterate, so a sweep is one embarrassingly parallel kernel.  `safe` != 0 uses
 * divide_no_nan for the lower entries (rank-deficient matrices, _linalg.py:576).  lu_val must not alias A's values.
 * Same handle, same solves as phb200_precond_ilu0; on a recognised star stencil the structured wavefront solves stay
 * available (L is then described by the diagonal of the sweep before the last, U by the last one). */
int phb200_precond_ilu_sweeps(phb200_precond** out, const phb200_matrix* A, void* lu_val, int sweeps, int safe,
                              const phb200_matrix* stencil_twin /* may be NULL */, void* stream);
/* The reference's 'cluster' preconditioner (ExplicitClusterSolve, phiml/backend/_linalg.py:734-766; built by explicit_coarse /
 * coarse_explicit_preconditioner_coo, phiml/math/_optimize.py:942-971, _linalg.py:769-799): M^-1 v = P C^-1 R v + (v - P (R v / size)),
 * R = sum over the elements of a cluster (batched_bincount), C^-1 = dense inverse of the cluster_count^2 coarse matrix, P = gather by
 * cluster.  All arrays are caller-owned device memory: cluster_of (n, int32), members (n, int32: element indices grouped by cluster,
 * ascending inside a cluster), member_ptr (cluster_count + 1), inv_coarse (cluster_count^2, row-major) and cluster_size (cluster_count)
 * in the vector dtype.  Application = three kernels: segmented sums (deterministic order), dense coarse product, prolongation + correction. */
int phb200_precond_cluster(phb200_precond** out, const int32_t* cluster_of, const int32_t* members, const int32_t* member_ptr, int64_t n,
                           int cluster_count, const void* inv_coarse, const void* cluster_size, int dtype);
/* *star_wavefront = 1 if the triangular solves of an ILU preconditioner run as the structured wavefront of
 * csrc/star_trsv.cuh (factors of a recognised star stencil; 3 N w bytes per solve) instead of the CSR level solves.
 * The environment variable PHB200_TRSV=csr disables the wavefront form at set-up. */
int phb200_precond_info(const phb200_precond* p, int* kind, int* star_wavefront);
/* number of dependency levels of the lower / upper solve (diagnostics) */
int phb200_precond_levels(const phb200_precond* p, int64_t* lower_levels, int64_t* upper_levels);

/* Generic sparse triangular solve X = T^-1 B, T given as CSR with sorted columns (diagonal stored; ignored when
 * unit_diagonal).  B, X (batch, ld) row-major; X may alias B.  Builds the level schedule on every call unless a
 * `plan` from phb200_sptrsv_plan is passed. */
typedef struct phb200_trsv_plan phb200_trsv_plan;
int phb200_sptrsv_plan(phb200_trsv_plan** out, const phb200_matrix* T, int lower, void* stream);
int phb200_sptrsv(const phb200_matrix* T, const phb200_trsv_plan* plan, int lower, int unit_diagonal,
                  const void* B, void* X, int64_t batch, int64_t ld, void* stream);
int phb200_sptrsv_plan_destroy(phb200_trsv_plan* p);

/* apply M^-1 to (batch, ld) vectors: out = M^-1 in */
int phb200_precond_apply(const phb200_precond* p, const void* in, void* out, int64_t batch, int64_t ld, void* stream);
int phb200_precond_destroy(phb200_precond* p);

/* ---------------------------------------------------------------- solver --------------------------------- */

/* A solver object is one batched solve in flight: `batch` independent right-hand sides sharing one matrix.
 * Life cycle:  create -> workspace_bytes -> start -> (advance | run)* -> result -> destroy.
 * X and R given to start() ARE the live iterate and residual (SolveResult.x / .residual); they are valid after
 * any advance()/run() has completed on the stream. */
int phb200_solver_create(phb200_solver** out, const phb200_matrix* A, const phb200_precond* M /* may be NULL */,
                         int method, int poly_order, int64_t batch);
size_t phb200_solver_workspace_bytes(const phb200_solver* s);

/* Y, X0: (batch, n) row-major (not modified).  X, R: (batch, n) outputs.  rtol, atol: device arrays (batch,) of the
 * matrix dtype.  max_iter: device int32 (batch,).  Enqueues the initial residual, tolerance (minimum over the batch —
 * the (batch,batch) broadcast of stop_on_l2, phiml/backend/_linalg.py:26-29) and first progress check. */
int phb200_solver_start(phb200_solver* s, const void* Y, const void* X0, void* X, void* R,
                        const void* rtol, const void* atol, const int32_t* max_iter,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Batch sharding over several GPUs (or a batch split over several solver objects): the tolerance every system must reach
 * couples the WHOLE batch (phiml/backend/_linalg.py:26-29) — it is the minimum over all systems, and for the generic
 * cg_adaptive / bicg loops it is relative to the SUM of |y_b|^2 over all systems (stop_on_l2 sums the (batch,) vector it
 * receives once more).  `vals` is a device array of two doubles {tol_min, yy_sum}.  After start():
 *   set = 0  read the local pair;
 *   set = 2  install yy_sum summed over all shards: tolerances and the local minimum are recomputed (no-op for methods whose
 *            tolerance is per system: PHB200_CG and the torch fast paths); read again afterwards;
 *   set = 1  install tol_min reduced (min) over all shards — the first progress check is repeated with it. */
int phb200_solver_tol_min(phb200_solver* s, void* tol_min, int set, void* stream);

/* Distributed (slab) mode: `totals` is a caller-owned device buffer of batch*4 doubles.  After set_dist, start() and every
 * half iteration leave their LOCAL dot products in `totals` and stop; the caller all-reduces `totals` (sum) over the ranks
 * and calls dist_step, which applies the scalar step on every rank identically:
 *     start -> allreduce -> dist_step(10) [-> halo exchange of R and the direction buffer]
 *     loop:   dist_step(0) -> allreduce -> dist_step(1) -> allreduce -> dist_step(2) -> halo exchange of R
 *     end:    dist_step(20)   (applies the pending x update)
 * Supported: 'CG' on a star stencil, no or constant-diagonal Jacobi preconditioner. */
int phb200_solver_set_dist(phb200_solver* s, void* totals);
int phb200_solver_dist_step(phb200_solver* s, int step, void* stream);

/* Fused halo exchange for the step-driven mode (one process per GPU): after phb200_solver_set_dist, hand the solver the
 * addresses of the slab neighbours' halo planes of THEIR residual arrays (device memory of the peer GPUs, mapped with
 * phb200_ipc_export / phb200_ipc_open = cudaIpcGet/OpenMemHandle).  The residual kernel (dist_step 1) then writes its first / last
 * owned plane there itself — coalesced stores over NVLink inside the kernel that produces them — and no separate exchange of R
 * follows dist_step(2).  The all-reduce between dist_step(1) and dist_step(2) orders the writes before the neighbours' next read. */
int phb200_solver_set_peer_halos(phb200_solver* s, void* lower_neighbour_halo, int64_t ld_lower, void* upper_neighbour_halo, int64_t ld_upper);
/* In-kernel all-reduce for the step-driven mode: every rank allocates a zero-filled mailbox of phb200_peer_mailbox_bytes(world,
 * batch) bytes, maps the mailboxes of all ranks (CUDA IPC) and passes the `world` device addresses (its own at index `rank`)
 * BEFORE start().  The kernels then all-reduce their dot products themselves over NVLink peer memory (stores into every rank's
 * mailbox, release/acquire at system scope, sum in rank order) and apply the scalar step at once: a CG iteration is
 * dist_step(0), dist_step(1), dist_step(2) = two kernels, no NCCL call and no separate finalize launch (dist_step(2) and (10)
 * only keep the host's pass count).  Up to 8 ranks. */
int phb200_solver_set_peer_reduce(phb200_solver* s, void* const* mailboxes, int world, int rank, void* stream);
size_t phb200_peer_mailbox_bytes(int world, int64_t batch);
/* With BOTH peer halos and the in-kernel all-reduce installed nothing of an iteration needs the host: enqueue n_iter whole
 * iterations (dist_step 0, 1, 2 each) in one call — two kernels per iteration back to back on the stream. */
int phb200_solver_dist_advance(phb200_solver* s, int n_iter, void* stream);
int phb200_ipc_export(const void* dev_ptr, void* handle64 /* 64 bytes out */, int64_t* offset /* of dev_ptr inside its allocation */);
int phb200_ipc_open(const void* handle64, void** base_ptr /* mapping of the peer's allocation */);
int phb200_ipc_close(void* base_ptr);

/* Callback-driven distributed mode: EVERY method ('CG', 'CG-adaptive', torch rules, 'biCG-stab(L)'; no / Jacobi
 * preconditioner) on a star stencil in slab mode.  The loop stays inside the library (start / advance / run as usual);
 * wherever the algorithm has a global reduction the kernel leaves the rank-local sums in `totals` (batch*4 doubles) and
 * the library calls `allreduce(totals, count, stream, user)` — the caller sums over the ranks (ncclAllReduce on `stream`;
 * phiml_b200/multi_gpu.py uses torch.distributed) — then applies the scalar step identically on every rank; before every
 * product it calls `halo(vec, stream, user)` so that the caller refreshes the two halo planes of the (batch, ld) extended
 * array `vec` (NVLink P2P / NCCL send-recv).  Callbacks return 0 on success.  BiCGstab(2): 9 reductions + 4 halo exchanges
 * per iteration (BASELINE config 5). */
typedef int (*phb200_allreduce_fn)(void* totals, int64_t count, void* stream, void* user);
typedef int (*phb200_halo_fn)(void* vec, void* stream, void* user);
int phb200_solver_set_dist_callbacks(phb200_solver* s, void* totals, phb200_allreduce_fn allreduce, phb200_halo_fn halo, void* user);

/* Native distributed mode: the callback-driven mode above with NOTHING on the host inside the loop (BASELINE config 5 across GPUs).
 * Reductions all-reduce inside the kernels that produce them (mailboxes as for phb200_solver_set_peer_reduce); before every product the
 * library launches a small kernel that stores the first / last owned plane of the product's input vector into the slab neighbours' halo
 * planes of THEIR copy of that vector (peer memory over NVLink) and exchanges one sequence word with each neighbour.  For that every rank
 * maps its neighbours' X, R and workspace allocations (phb200_ipc_export / _open) and passes them here, before start():
 *   X, R, ws   the neighbour's arrays as given to ITS phb200_solver_start (same method, same batch on every rank);
 *   ld         its leading dimension (elements per system; ranks may own different numbers of planes);
 *   halo_off   element offset, inside one of its vectors, of the halo plane this rank feeds (its upper halo plane for the lower neighbour,
 *              its lower halo plane for the upper neighbour);
 *   flag       address, in the neighbour's memory, of the sequence word this rank writes: its my_flags[1] for the lower neighbour
 *              (this rank is ITS upper neighbour), its my_flags[0] for the upper one.
 * my_flags: two zero-filled 8-byte words in this rank's memory.  A missing neighbour is passed as NULL (or X == NULL). */
typedef struct {
    void* X;
    void* R;
    void* ws;
    int64_t ld;
    int64_t halo_off;
    void* flag;
} phb200_peer_slab;
int phb200_solver_set_peer_native(phb200_solver* s, void* totals, void* const* mailboxes, int world, int rank,
                                  const phb200_peer_slab* lower, const phb200_peer_slab* upper, void* my_flags, void* stream);

/* Enqueue up to n_iter further iterations (asynchronous; iterations after all systems finished are skipped on
 * the device). */
int phb200_solver_advance(phb200_solver* s, int n_iter, void* stream);

/* Run until no system continues or iteration_cap loop passes were made; polls a pinned flag every few
 * iterations, never once per iteration.  Synchronises the stream before returning. */
int phb200_solver_run(phb200_solver* s, int64_t iteration_cap, void* stream);

/* Small systems (N*3 vectors fit into the shared memory of one CTA or one cluster of <= 8 CTAs; 'CG' on a star stencil,
 * no or constant-diagonal Jacobi preconditioner): phb200_solver_run executes the whole Krylov loop of a system inside ONE
 * kernel, iterate/direction on chip, HBM touched once per solve.  PHB200_OPT_RESIDENT: -1 a cost model decides (default),
 * 0 never, 1 whenever the system fits.  PHB200_OPT_RESIDENT_USED (read-only via *previous): did the last run() use it. */
#define PHB200_OPT_RESIDENT 1
#define PHB200_OPT_RESIDENT_USED 2
int phb200_solver_option(phb200_solver* s, int option, int value, int* previous /* may be NULL */);

/* *any_continue = 1 while at least one system is still iterating.  Synchronises the stream. */
int phb200_solver_poll(phb200_solver* s, int* any_continue, void* stream);

/* Device outputs, (batch,) each: iterations, function_evaluations int32; converged, diverged uint8.
 * PHB200_CG / PHB200_CG_ADAPTIVE (generic loops, phiml/backend/_linalg.py:52-128): systems that stopped on a non-finite residual norm while other
 * systems of the batch went on iterating get X and R filled with NaN here — the reference's masked updates (step_size * continue_ = NaN * 0) do
 * that on the next pass; the device loop skips stopped systems instead. */
int phb200_solver_result(phb200_solver* s, int32_t* iterations, int32_t* function_evaluations,
                         uint8_t* converged, uint8_t* diverged, void* stream);
/* loop passes executed so far (host counter), kernels launched by this solver */
int phb200_solver_stats(const phb200_solver* s, int64_t* passes_enqueued, int64_t* launches);
/* Optional per-kernel timing: while enabled, the k-th kernel of every pass is bracketed by CUDA events on the launch
 * stream; profile_read() sums the elapsed milliseconds per position k (and clears the record).  Measurement aid for
 * bench.py's roofline line; leave disabled otherwise. */
int phb200_solver_profile(phb200_solver* s, int enable);
int phb200_solver_profile_read(phb200_solver* s, double* ms_per_slot, int64_t* count_per_slot, int max_slots, int* n_slots);
int phb200_solver_destroy(phb200_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* PHB200_H */
