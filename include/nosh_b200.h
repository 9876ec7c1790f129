/* =============================================================================
 * nosh_b200.h -- C ABI of the B200-native Newton-Krylov hot path of nschloe/nosh.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every entry point names the reference interface (file:line under the
 * reference tree) it replaces.  The C++ mirror of the reference's classes
 * (nosh::parameter_matrix::keo, nosh::jacobian_operator,
 * nosh::model_evaluator::nls, ...) lives in nosh_b200/hostcpp/nosh.hpp and is a thin
 * forwarder to these functions; INTEGRATION.md shows the binding a reference
 * maintainer would add.
 *
 * Conventions
 *  - One nosh_ctx per process per GPU.  Not thread-safe per ctx; safe across ctxs.
 *  - Every function returns a nosh_status; nosh_last_error(ctx) gives the message.
 *    Nothing throws across this boundary.
 *  - State vectors are interleaved (re,im) doubles, entry 2k/2k+1 for the k-th
 *    OWNED vertex (reference: src/mesh.cpp:595-606, src/jacobian_operator.cpp:95-100);
 *    multi-vectors are column-major with an explicit leading dimension.
 *  - Vector arguments may be HOST or DEVICE pointers (detected with
 *    cudaPointerGetAttributes).  Host vectors are staged through the ctx's
 *    device buffers (H2D / D2H inside the call); device vectors are used in place.
 *    Calls whose vectors are all on the device only ENQUEUE work on the ctx's stream and may return
 *    before it has run: order them against other streams by creating the ctx on your own stream
 *    (nosh_ctx_create) or with nosh_ctx_synchronize.  Calls with host vectors return when the
 *    result is in host memory.
 *  - Per-vertex field arrays given at set-up are HOST arrays in LOCAL numbering:
 *    owned vertices first, then ghosts (nosh_mesh_local_gids).  On one GPU
 *    local == global.
 *  - Model parameters are passed as (names[], values[], n), the C image of the
 *    reference's std::map<std::string,double>; a missing name yields NOSH_EKEY
 *    (the reference's params.at() throws std::out_of_range).
 *  - All arithmetic is IEEE fp64; indices are int32, as in
 *    Tpetra::CrsMatrix<double,int,int> (src/jacobian_operator.hpp:26).
 *  - There is NO CPU fallback: every compute entry point runs sm_100a kernels.
 * ============================================================================= */
#ifndef NOSH_B200_H
#define NOSH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NOSH_API __attribute__((visibility("default")))
#else
#define NOSH_API
#endif

typedef struct nosh_ctx nosh_ctx;

typedef enum {
  NOSH_OK = 0,
  NOSH_EINVAL = 1,       /* bad argument (incl. unsupported mode/alpha/beta of apply) */
  NOSH_ECUDA = 2,        /* CUDA runtime failure */
  NOSH_ESTATE = 3,       /* call sequence error (e.g. apply before fill) */
  NOSH_EMESH = 4,        /* illegal mesh (flat tetrahedron, degenerate cell) */
  NOSH_EKEY = 5,         /* parameter name missing (std::out_of_range in the reference) */
  NOSH_ECOMM = 6,        /* communication failure (peer-memory time-out, NCCL error, callback error) */
  NOSH_EUNSUPPORTED = 7  /* out of scope (e.g. .h5m / Exodus mesh files: MOAB is not available) */
} nosh_status;

/* Teuchos::ETransp of Tpetra::Operator::apply */
typedef enum { NOSH_NO_TRANS = 0, NOSH_TRANS = 1, NOSH_CONJ_TRANS = 2 } nosh_transp;

/* storage layout of the complex block matrix in HBM */
typedef enum { NOSH_LAYOUT_CSR = 0, NOSH_LAYOUT_SELL32 = 1 } nosh_layout;

/* which matrix nosh_get_block_csr exports */
typedef enum { NOSH_MAT_KEO = 0, NOSH_MAT_DKEO = 1 } nosh_matrix_id;

/* linear operator selector for the Krylov solvers */
typedef enum { NOSH_OP_JACOBIAN = 0, NOSH_OP_KEO = 1, NOSH_OP_KEOREG = 2 } nosh_operator_id;

/* preconditioner selector of the Krylov solvers: none (the live reference path,
 * src/model_evaluator_nls.cpp:291) or one AMG V-cycle on the regularised KEO -- what
 * create_W_prec hands to the solver (src/model_evaluator_nls.cpp:301-314) */
typedef enum { NOSH_PREC_NONE = 0, NOSH_PREC_KEOREG_AMG = 1 } nosh_precond;

/* hierarchy reuse across nosh_keoreg_rebuild calls: NONE = new hierarchy per rebuild; FULL = the
 * reference's MueLu setting "reuse: type" = "full" (src/keo_regularized.cpp:300): aggregates,
 * prolongators and coarse operators of the first build are kept, the finest level follows the
 * current matrix */
typedef enum { NOSH_AMG_REUSE_NONE = 0, NOSH_AMG_REUSE_FULL = 1 } nosh_amg_reuse;

/* Krylov solver of the Newton / continuation drivers: the "Solver Type" values of the reference's Belos
 * parameter lists (src/model_evaluator_nls.cpp:280-282, examples/conf.xml:103-105) */
typedef enum { NOSH_SOLVER_MINRES = 0, NOSH_SOLVER_CG = 1, NOSH_SOLVER_GMRES = 2 } nosh_linear_solver;

/* ---- lifecycle ------------------------------------------------------------- */
NOSH_API const char *nosh_version(void);
/* stream: a cudaStream_t (as void*) all work is enqueued on, or NULL for a
 * ctx-owned stream. */
NOSH_API nosh_status nosh_ctx_create(int device, void *stream, nosh_ctx **out);
NOSH_API void nosh_ctx_destroy(nosh_ctx *ctx);
NOSH_API const char *nosh_last_error(const nosh_ctx *ctx);
/* options must be set before the mesh: layout (default SELL32, env NOSH_B200_LAYOUT
 * =csr|sell32) and the reduction/partition granularity in vertices (multiple of 512,
 * default 65536, env NOSH_B200_GROUP). */
NOSH_API nosh_status nosh_ctx_set_layout(nosh_ctx *ctx, nosh_layout layout);
NOSH_API nosh_status nosh_ctx_set_group_vertices(nosh_ctx *ctx, int64_t group_vertices);
NOSH_API nosh_status nosh_ctx_synchronize(nosh_ctx *ctx);
/* Pipelined host I/O (optional).  nosh_prefetch starts the host-to-device copy of a HOST state vector (n_owned
 * complex entries; pinned memory for a truly asynchronous copy) on the ctx's copy stream and returns; the next
 * call that is given the same host pointer waits for that copy instead of making its own -- issue it before a
 * long call (a Krylov solve) and the next step's inputs travel while that solve runs.  With async output enabled,
 * calls whose result vector is in host memory enqueue the device-to-host copy on the copy stream and return
 * without waiting for it: the data is valid after nosh_ctx_synchronize (scalar results -- iteration counts,
 * norms -- are always returned synchronously).  Default: off, every call returns with its result in host memory. */
NOSH_API nosh_status nosh_prefetch(nosh_ctx *ctx, const double *host_vector);
NOSH_API nosh_status nosh_ctx_set_async_output(nosh_ctx *ctx, int enabled);

/* ---- multi-GPU (one process per GPU).  Replaces Teuchos::MpiComm / Tpetra
 * Import / reduceAll (src/mesh_reader.cpp:53-57; Tpetra, not in tree). ---------- */
NOSH_API nosh_status nosh_comm_unique_id(void *id128 /* 128 bytes out */);
/* host-only helper (no GPU needed): the contiguous, group-aligned global vertex range
 * [begin,end) owned by `rank` of `nranks`, and the group size actually used (the requested one
 * doubled until at most 1024 groups cover the mesh).  This is the partition every ctx uses. */
NOSH_API nosh_status nosh_partition_range(int64_t n_global, int nranks, int rank,
                                          int64_t group_vertices, int64_t *begin, int64_t *end,
                                          int64_t *group_used);
NOSH_API nosh_status nosh_ctx_comm_init(nosh_ctx *ctx, const void *id128, int rank, int nranks);
/* The same with the CALLER'S communicator instead of a library-owned NCCL one: `allgather` is the image of
 * MPI_Allgather on the Teuchos::Comm the reference's mesh carries (src/mesh_reader.cpp:53-57) -- it receives
 * one host record of bytes_per_rank bytes and must return the records of all ranks, in rank order, in recv
 * (0 = success).  It is only called during set-up (nosh_mesh_*: halo plan, CUDA IPC handles) and by GMRES'
 * batched sums; every other exchange -- halos, MINRES / CG / Newton reductions -- is done by kernels that
 * store into the peers' HBM over NVLink (CUDA IPC).  Ranks may also share one GPU (tests). */
typedef int (*nosh_allgather_fn)(void *user, const void *send, void *recv, int64_t bytes_per_rank);
NOSH_API nosh_status nosh_ctx_comm_init_host(nosh_ctx *ctx, int rank, int nranks, nosh_allgather_fn allgather,
                                             void *user);
/* set-up timings and counters by name ("setup.mesh_s", "setup.halo_s", "setup.p2p_s", "p2p", "sell_sigma", ...);
 * NOSH_EKEY if unknown */
NOSH_API nosh_status nosh_ctx_get_stat(nosh_ctx *ctx, const char *key, double *value);
/* all recorded stats as "key=value\n" lines ("amg.setup.<phase>" = seconds of every phase of the last hierarchy
 * build, "sell.stored_over_blocks", ...); NUL terminated, truncated to cap bytes */
NOSH_API nosh_status nosh_ctx_list_stats(nosh_ctx *ctx, char *buf, int64_t cap);

/* ---- mesh (a1-a3).  Replaces nosh::read + mesh_tetra/mesh_tri ctor:
 * src/mesh_reader.cpp:19-162, src/mesh.cpp:18-51,629-691, src/mesh_tetra.cpp:14-35,
 * 41-271, src/mesh_tri.cpp:44-210.  Builds, on the device: the unique edge list,
 * cell->edge relation, edge lengths and FVM edge coefficients ("covolume"),
 * circumcentric control volumes and the block-CSR graph (src/mesh.cpp:787-881).
 * With nranks > 1 every rank passes the same GLOBAL mesh and keeps the part it
 * owns (contiguous vertex ranges) plus a one-cell halo. */
NOSH_API nosh_status nosh_mesh_set(nosh_ctx *ctx, int dim, int64_t n_vertices,
                                   const double *coords /* host, n_vertices x 3 */,
                                   int64_t n_cells,
                                   const int32_t *cells /* host, n_cells x (dim+1) */);
/* Partitioned ingestion -- the analogue of MOAB's "PARALLEL=READ_PART;PARTITION=PARALLEL_PARTITION;
 * PARALLEL_RESOLVE_SHARED_ENTS" read of src/mesh_reader.cpp:32-35: every rank passes only ITS part of a mesh with
 * n_global vertices: the cells that touch a vertex of its owned range (nosh_partition_range; further cells are
 * allowed and dropped), the nv_local vertices those cells use -- global id and coordinates -- and the cells as
 * indices into that local vertex list.  Nothing global is uploaded; results are identical to nosh_mesh_set with
 * the whole mesh on every rank.  NOSH_EMESH for ids out of range, duplicate ids or degenerate cells. */
NOSH_API nosh_status nosh_mesh_set_local(nosh_ctx *ctx, int dim, int64_t n_global, int64_t nv_local,
                                         const int64_t *vertex_gids /* host, nv_local */,
                                         const double *coords /* host, nv_local x 3 */, int64_t nc_local,
                                         const int32_t *cells /* host, nc_local x (dim+1), local indices */);
/* Synthetic input of SURVEY.md 8(d): nx*ny*nz structured vertices on [lo,hi], x fastest,
 * 6 Kuhn tetrahedra per hex cell, interior vertices displaced by jitter*h*U(-1,1)
 * (splitmix64 keyed on seed and the global vertex id).  Generated on the device, each
 * rank its own part. */
NOSH_API nosh_status nosh_mesh_tetgrid(nosh_ctx *ctx, int nx, int ny, int nz, const double lo[3],
                                       const double hi[3], double jitter, uint64_t seed);

typedef struct {
  int32_t dim;
  int64_t n_global;    /* vertices of the whole mesh */
  int64_t owned_begin; /* first global vertex id owned by this ctx */
  int64_t n_owned;     /* rows / vector length 2*n_owned */
  int64_t n_ghost;
  int64_t n_cells;     /* local cells (touching an owned vertex) */
  int64_t n_edges;     /* local edges (>= 1 owned endpoint) */
  int64_t n_blocks;    /* complex blocks in the owned rows (CSR count, no padding) */
  int64_t n_stored;    /* stored complex blocks incl. layout padding */
} nosh_mesh_info_t;
NOSH_API nosh_status nosh_mesh_info(const nosh_ctx *ctx, nosh_mesh_info_t *info);
/* parity accessors (host outputs; any may be NULL) */
NOSH_API nosh_status nosh_mesh_local_gids(nosh_ctx *ctx, int64_t *gids /* n_owned+n_ghost */);
NOSH_API nosh_status nosh_mesh_get_coords(nosh_ctx *ctx, double *coords /* (n_owned+n_ghost) x 3 */);
NOSH_API nosh_status nosh_mesh_get_cells(nosh_ctx *ctx, int32_t *cells /* n_cells x (dim+1), local ids */);
NOSH_API nosh_status nosh_mesh_get_edges(nosh_ctx *ctx, int32_t *edges /* n_edges x 2 local ids */,
                                         double *length, double *covolume);
NOSH_API nosh_status nosh_mesh_get_control_volumes(nosh_ctx *ctx, double *cv /* n_owned */);

/* ---- fields (a5-a8) ---------------------------------------------------------- */
/* scalar_field::constant(mesh, c) as thickness (src/scalar_field_constant.cpp:43-58), or
 * explicit per-vertex values (values != NULL, local numbering). */
NOSH_API nosh_status nosh_set_thickness(nosh_ctx *ctx, const double *values, double c);
/* scalar_field::constant(mesh, c, param1_name, .) as the scalar potential:
 * V = c (+ params[param1_name] if present), dV/dp = 1 for its own name
 * (src/scalar_field_constant.cpp:43-75).  param1_name may be NULL/"". */
NOSH_API nosh_status nosh_set_potential_constant(nosh_ctx *ctx, double c, const char *param1_name);
/* scalar_field::explicit_values: V = params["beta"] * values, dV/dbeta = values
 * (src/scalar_field_explicit_values.cpp:33-62). */
NOSH_API nosh_status nosh_set_potential_values(nosh_ctx *ctx, const double *values);
/* vector_field::explicit_values(mesh, "A", mu): cache_e = 0.5(A_v0+A_v1).(x_v0-x_v1),
 * a_e = mu*cache_e (src/vector_field_explicit_values.cpp:13-90).  A: local n x 3. */
NOSH_API nosh_status nosh_set_mvp_explicit(nosh_ctx *ctx, const double *A);
/* same, with A = 0.5 B x X evaluated on the device from the vertex coordinates
 * (examples/state-equippers/plain-gl:22-39) -- no host array needed at scale. */
NOSH_API nosh_status nosh_set_mvp_explicit_curl(nosh_ctx *ctx, const double B[3]);
/* vector_field::constantCurl(mesh, b, u): a_e = mu * R_theta(b) . (0.5 x_v1 x x_v0)
 * (src/vector_field_constant_curl.cpp:74-227; edge cache restated).
 * u may be NULL.  b and u must be exactly normalised (:35-44) else NOSH_EINVAL. */
NOSH_API nosh_status nosh_set_mvp_constcurl(nosh_ctx *ctx, const double b[3], const double u[3]);
/* parity accessors */
NOSH_API nosh_status nosh_get_alpha_cache(nosh_ctx *ctx, double *alpha /* n_edges */);
NOSH_API nosh_status nosh_get_edge_projection(nosh_ctx *ctx, int np, const char *const *names,
                                              const double *values, const char *dname /* or NULL */,
                                              double *a /* n_edges */, double *da /* or NULL */);

/* ---- KEO (a9, a10).  parameter_matrix::keo::set_parameters -> refill_
 * (src/parameter_object.cpp:6-47, src/parameter_matrix_keo.cpp:74-184) and
 * DkeoDP::refill_ (src/parameter_matrix_dkeo_dp.cpp:60-155).  Needs "mu" (and "theta"
 * for constantCurl). */
NOSH_API nosh_status nosh_keo_fill(nosh_ctx *ctx, int np, const char *const *names,
                                   const double *values);
NOSH_API nosh_status nosh_dkeo_fill(nosh_ctx *ctx, int np, const char *const *names,
                                    const double *values, const char *dname);
/* Tpetra::CrsMatrix::apply of the KEO / dKEO (call sites src/model_evaluator_nls.cpp:537,641,
 * test/keo.cpp:79-80): Y = alpha*op(A)*X + beta*Y.  A is Hermitian as a complex matrix,
 * i.e. symmetric in the real layout, so TRANS == NO_TRANS; all alpha/beta supported. */
NOSH_API nosh_status nosh_matrix_apply(nosh_ctx *ctx, nosh_matrix_id which, const double *X,
                                       int64_t ldx, double *Y, int64_t ldy, int nvec,
                                       nosh_transp mode, double alpha, double beta);
/* block-CSR export for entry-wise parity: rowptr n_owned+1 (int64), cols n_blocks (local
 * ids, ascending global id within a row), vals n_blocks complex (re,im). */
NOSH_API nosh_status nosh_get_block_csr(nosh_ctx *ctx, nosh_matrix_id which, int64_t *rowptr,
                                        int32_t *cols, double *vals);

/* ---- Jacobian (a11, a12).  jacobian_operator::rebuild / apply
 * (src/jacobian_operator.cpp:38-199).  rebuild needs "g" + KEO params.  apply supports
 * only NO_TRANS, alpha == 1, beta == 0 -- anything else is NOSH_EINVAL (:48-59 throw). */
NOSH_API nosh_status nosh_jac_rebuild(nosh_ctx *ctx, int np, const char *const *names,
                                      const double *values, const double *psi);
NOSH_API nosh_status nosh_jac_apply(nosh_ctx *ctx, const double *X, int64_t ldx, double *Y,
                                    int64_t ldy, int nvec, nosh_transp mode, double alpha,
                                    double beta);
NOSH_API nosh_status nosh_jac_get_diags(nosh_ctx *ctx, double *d0 /* 2n */, double *d1b /* n */);

/* ---- model evaluator (a13, a14).  nls::compute_f_ / computeDFDP_
 * (src/model_evaluator_nls.cpp:527-695).  Both refill the (d)KEO first, as the
 * reference does; an unchanged parameter set is detected and the refill skipped
 * (the cache the reference meant to have, src/parameter_object.cpp:17-44). */
NOSH_API nosh_status nosh_compute_f(nosh_ctx *ctx, int np, const char *const *names,
                                    const double *values, const double *psi, double *f);
NOSH_API nosh_status nosh_compute_dfdp(nosh_ctx *ctx, int np, const char *const *names,
                                       const double *values, const char *pname, const double *psi,
                                       double *dfdp);

/* ---- preconditioner (a16, a17).  keo_regularized::rebuild
 * (src/keo_regularized.cpp:181-264): P = K + blockdiag([[al+ga, be],[be, al-ga]]), g > 0
 * only.  nosh_keoreg_matrix_apply applies P.  nosh_keoreg_apply is keo_regularized::apply
 * (:88-165): ONE V-cycle of a smoothed-aggregation AMG hierarchy for P with 2 equations per
 * node.  The reference gets that hierarchy from MueLu (third party, not in its tree:
 * parity unpinned); here it is built and applied on the device (nosh_b200/csrc/amg.cu,
 * restated on the CPU in oracle/amg.py).  Like the reference's apply it supports only
 * NO_TRANS, alpha == 1, beta == 0 (:98-100).  The hierarchy is built lazily by the first
 * apply / preconditioned solve after a rebuild and kept according to the reuse policy.
 * With several GPUs every rank preconditions its own diagonal block (no communication). */
NOSH_API nosh_status nosh_keoreg_rebuild(nosh_ctx *ctx, int np, const char *const *names,
                                         const double *values, const double *psi);
NOSH_API nosh_status nosh_keoreg_matrix_apply(nosh_ctx *ctx, const double *X, int64_t ldx,
                                              double *Y, int64_t ldy, int nvec);
NOSH_API nosh_status nosh_keoreg_get_diags(nosh_ctx *ctx, double *d0 /* 2n */, double *d1b /* n */);
NOSH_API nosh_status nosh_keoreg_apply(nosh_ctx *ctx, const double *X, int64_t ldx, double *Y,
                                       int64_t ldy, int nvec, nosh_transp mode, double alpha,
                                       double beta);

/* AMG options (before the hierarchy is built; changing them drops it): Chebyshev degree of the
 * pre-/post-smoother on the finest level (>= 1; 1 = damped l1-Jacobi) and on the coarse levels,
 * number of nodes at which coarsening stops and a dense inverse is used (<= 4096), maximum number
 * of levels, reuse policy.  Values <= 0 (reuse < 0) keep the current setting.
 * Defaults: 1, 2, 512, 10, FULL. */
NOSH_API nosh_status nosh_amg_set_options(nosh_ctx *ctx, int degree, int coarse_degree, int coarse_max,
                                          int max_levels, int reuse);
/* (re)build the hierarchy now for the current regularised KEO */
NOSH_API nosh_status nosh_amg_setup(nosh_ctx *ctx);
#define NOSH_AMG_MAX_LEVELS 16
typedef struct {
  int32_t levels;
  int32_t degree;        /* finest level */
  int32_t coarse_degree; /* levels >= 1 */
  int32_t reserved;
  int64_t nodes[NOSH_AMG_MAX_LEVELS];     /* block rows per level */
  int64_t blocks[NOSH_AMG_MAX_LEVELS];    /* 2x2 blocks per level (level 0: complex blocks of the owned columns) */
  int64_t p_blocks[NOSH_AMG_MAX_LEVELS];  /* blocks of the prolongator from level l+1 to l */
  double lambda_max[NOSH_AMG_MAX_LEVELS]; /* power-iteration estimate of lambda_max(D^-1 A) (prolongator damping) */
  double setup_seconds;
} nosh_amg_info_t;
NOSH_API nosh_status nosh_amg_info(nosh_ctx *ctx, nosh_amg_info_t *info);
/* parity accessors (host outputs, any may be NULL): aggregate of every node of `level`;
 * block CSR of the level matrix (level >= 1) / of the prolongator from level+1 to level:
 * rowptr (rows+1, int64), cols, vals (4 doubles per block, row-major 2x2) */
NOSH_API nosh_status nosh_amg_get_aggregates(nosh_ctx *ctx, int level, int32_t *agg);
NOSH_API nosh_status nosh_amg_get_matrix(nosh_ctx *ctx, int level, int64_t *rowptr, int32_t *cols,
                                         double *vals);
NOSH_API nosh_status nosh_amg_get_prolongator(nosh_ctx *ctx, int level, int64_t *rowptr, int32_t *cols,
                                              double *vals);

/* ---- vector reductions (Tpetra::MultiVector::dot / norm2; partition independent) */
NOSH_API nosh_status nosh_dot(nosh_ctx *ctx, const double *x, const double *y, double *result);
NOSH_API nosh_status nosh_norm2(nosh_ctx *ctx, const double *x, double *result);

/* ---- Krylov (a18).  Belos::MinresSolMgr / PseudoBlockCGSolMgr as selected at
 * src/model_evaluator_nls.cpp:280-291 and examples/conf.xml:104-123: unpreconditioned,
 * x0 = 0, implicit relative residual <= tol, at most maxit iterations.
 * hist (or NULL): maxit+1 doubles, relative residual estimate after each iteration. */
typedef struct {
  int32_t iterations;
  int32_t converged;
  double relres;     /* final implicit relative residual */
  int32_t breakdown; /* 1: the recurrence broke down (gamma == 0, or <r, M r> < 0 with an indefinite M) */
  int32_t reserved;
} nosh_krylov_result;
NOSH_API nosh_status nosh_minres(nosh_ctx *ctx, nosh_operator_id op, const double *b, double *x,
                                 double tol, int maxit, nosh_krylov_result *res, double *hist);
NOSH_API nosh_status nosh_cg(nosh_ctx *ctx, nosh_operator_id op, const double *b, double *x,
                             double tol, int maxit, nosh_krylov_result *res, double *hist);
/* The same solvers with a preconditioner M (Belos with a left preconditioner from
 * Thyra::nonconstUnspecifiedPrec, src/model_evaluator_nls.cpp:313).  MINRES: beta_k^2 = <r_k, M r_k>,
 * stop on the implicit M-norm residual phibar / beta_1 <= tol.  CG: stop on ||r||_2 / ||r_0||_2. */
NOSH_API nosh_status nosh_minres_prec(nosh_ctx *ctx, nosh_operator_id op, nosh_precond prec,
                                      const double *b, double *x, double tol, int maxit,
                                      nosh_krylov_result *res, double *hist);
NOSH_API nosh_status nosh_cg_prec(nosh_ctx *ctx, nosh_operator_id op, nosh_precond prec, const double *b,
                                  double *x, double tol, int maxit, nosh_krylov_result *res,
                                  double *hist);
/* Restarted GMRES(restart): the solver examples/conf.xml:104 selects ("Pseudo Block GMRES", tolerance
 * 1e-10, 1000 iterations; Belos' default restart length "Num Blocks" = 300).  Two-pass iterated
 * classical Gram-Schmidt (Belos' default "ICGS"), x0 = 0, stop on the implicit relative residual
 * |g_{j+1}| / ||b|| <= tol; prec is applied from the RIGHT (x = M y), so residuals are true residuals.
 * Needs (restart+1) extra vectors of device memory.  restart in [1, 500]. */
NOSH_API nosh_status nosh_gmres(nosh_ctx *ctx, nosh_operator_id op, nosh_precond prec, const double *b,
                                double *x, double tol, int maxit, int restart, nosh_krylov_result *res,
                                double *hist);
/* preconditioner the Newton / continuation drivers give their linear solves (default NONE);
 * with KEOREG_AMG every Newton step also does keo_regularized::rebuild at the current state,
 * the evalModel(W_prec) of src/model_evaluator_nls.cpp:507-522 */
NOSH_API nosh_status nosh_ctx_set_preconditioner(nosh_ctx *ctx, nosh_precond prec);
/* Krylov solver of the Newton / continuation drivers (default MINRES; the reference's live default is
 * "Pseudo Block CG", its conf.xml selects "Pseudo Block GMRES"); gmres_restart <= 0 keeps the setting */
NOSH_API nosh_status nosh_ctx_set_linear_solver(nosh_ctx *ctx, nosh_linear_solver solver, int gmres_restart);

/* ---- Newton.  NOX "Line Search Based"/"Full Step" with a NormF test as configured in
 * examples/conf.xml:76-191, driving evalModel(f), evalModel(W_op) and the MINRES solve
 * (call stack SURVEY.md 3.4).  psi: in = initial guess, out = solution.
 * lin_iters: nl_maxit ints; fnorms: nl_maxit+1 doubles (either may be NULL). */
typedef struct {
  int32_t steps;
  int32_t converged;
  int32_t total_linear_iterations;
  int32_t linear_solve_status; /* 0: every linear solve converged; 1: one hit lin_maxit (its step was still
                                  taken, as NOX does); 2: one broke down -- no step taken, Newton stopped */
  double fnorm;
} nosh_newton_result;
NOSH_API nosh_status nosh_newton(nosh_ctx *ctx, int np, const char *const *names,
                                 const double *values, double *psi, double nl_tol, int nl_maxit,
                                 double lin_tol, int lin_maxit, nosh_newton_result *res,
                                 int32_t *lin_iters, double *fnorms);

/* ---- model-evaluator scalars.  nls::inner_product / norm / gibbs_energy
 * (src/model_evaluator_nls.cpp:699-770, src/model_evaluator_base.hpp:50-60).  The reference
 * computes the local sums, leaves the global reduction as a TODO and returns 0.0; these return
 * what its commented code states: sum_k c_k Re(conj(phi_k) psi_k) / sum_k c_k, its square root,
 * and -sum_k c_k |psi_k|^4 / sum_k c_k. */
NOSH_API nosh_status nosh_inner_product(nosh_ctx *ctx, const double *phi, const double *psi,
                                        double *result);
NOSH_API nosh_status nosh_gibbs_energy(nosh_ctx *ctx, const double *psi, double *result);

/* ---- parameter continuation ("next" row f2; nosh-cont, executables/nosh-cont/nosh-cont.cpp:206-344
 * with the LOCA settings of examples/conf.xml:35-75): natural continuation in `pname` with a
 * tangent predictor (J t = -dF/dp by MINRES, psi += dp t) and the Newton corrector, nsteps steps
 * of size dp starting at the value of `pname` in the parameter list.  steps: nsteps+1 records --
 * the columns src/observer.cpp:134-159 writes (step, parameter, Gibbs energy, ||x||) plus solver
 * counts.  Stops at the first step whose Newton corrector fails (later records have step = -1). */
typedef struct {
  int32_t step;
  int32_t converged;
  int32_t newton_steps;
  int32_t linear_iterations;           /* MINRES iterations of the corrector */
  int32_t predictor_linear_iterations; /* MINRES iterations of the tangent solve */
  int32_t reserved;
  double param;
  double gibbs_energy;
  double norm;  /* sqrt(inner_product(psi, psi)) */
  double fnorm;
} nosh_continuation_step;
/* LOCA::Thyra::SaveDataStrategy::saveSolution(x, p) (src/continuation_data_saver.hpp:24-50: the outNNNN dumps) and
 * observer::observeSolution (src/observer.cpp:134-159: the CSV row): called by both continuation drivers after every
 * accepted step with the step index, the parameter value, the Gibbs energy and scaled norm of the CSV, and the
 * solution -- this rank's owned entries, interleaved (re,im), in HOST memory valid during the call.  A non-zero
 * return stops the continuation (the steps so far are returned; with several ranks every rank's observer is
 * called and all must return the same decision).  NULL removes the observer. */
typedef int (*nosh_step_observer_fn)(void *user, int step, double param, double gibbs_energy, double norm,
                                     const double *psi_host, int64_t n_doubles);
NOSH_API nosh_status nosh_ctx_set_step_observer(nosh_ctx *ctx, nosh_step_observer_fn fn, void *user);
NOSH_API nosh_status nosh_continuation(nosh_ctx *ctx, int np, const char *const *names,
                                       const double *values, const char *pname, double dp, int nsteps,
                                       double *psi, double nl_tol, int nl_maxit, double lin_tol,
                                       int lin_maxit, nosh_continuation_step *steps);

/* Pseudo-arclength continuation: the LOCA configuration nosh-cont actually uses
 * (examples/conf.xml:35-75: "Continuation Method" = "Arc Length", "Predictor" = "Tangent", adaptive
 * step size with aggressiveness 2, failed steps halved).  LOCA is third party (not in the reference
 * tree, unpinned); the algorithm is the bordering form restated in oracle/continuation.py:
 * constraint <xdot, x - x0>/len + pdot (p - p0) = ds with LOCA's scaled dot product (Euclidean / vector
 * length, parameter scaling 1), Newton on the bordered system by two MINRES solves with the same
 * Jacobian (J a = -F, J b = -dF/dp), tangent J t = -dF/dp after every accepted step.
 * steps: max_steps+1 records (step 0 = the solution at the initial parameter value; +1 with NOSH_ARC_HIT_BOUND);
 * *n_records = number written.  Follows the branch through turning points, where the natural
 * continuation of nosh_continuation fails. */
typedef struct {
  double initial_step_size; /* signed: the direction of the first step in the parameter */
  double min_step_size, max_step_size;
  double aggressiveness;
  int32_t max_steps;
  int32_t nl_maxit;
  double nl_tol;   /* on sqrt(||F||^2 + g^2) */
  double lin_tol;
  int32_t lin_maxit;
  int32_t flags;               /* NOSH_ARC_SCALING | NOSH_ARC_HIT_BOUND; 0: neither */
  double min_value, max_value; /* stop once the parameter leaves [min_value, max_value] */
  /* NOSH_ARC_SCALING (LOCA "Enable Arc Length Scaling", its default): 0 selects LOCA's defaults 0.5, 0.8, 1e-3, 1 */
  double goal_contribution, max_contribution, min_scale, initial_scale;
} nosh_arclength_options;
/* flags: the two LOCA stepper defaults nosh-cont inherits (examples/conf.xml sets neither).
 * NOSH_ARC_SCALING: the parameter enters the arc length as s*p; s is reset whenever the parameter's share of the
 * tangent s*|dp/ds| exceeds max_contribution so that it becomes goal_contribution; the step sizes of the options
 * are then parameter increments (converted with the first tangent).
 * NOSH_ARC_HIT_BOUND (LOCA "Hit Continuation Bound"): the step that would cross min_value / max_value is shortened
 * to land on the bound, and a final natural-continuation step ends the run ON the bound (one more record:
 * steps[] needs max_steps + 2 entries). */
#define NOSH_ARC_SCALING 1
#define NOSH_ARC_HIT_BOUND 2
typedef struct {
  int32_t step;
  int32_t converged;
  int32_t newton_steps;
  int32_t linear_iterations;           /* MINRES iterations of the corrector (two solves per Newton step) */
  int32_t predictor_linear_iterations; /* MINRES iterations of the tangent solve this step started from */
  int32_t reserved;
  double param;
  double gibbs_energy;
  double norm;
  double fnorm;
  double step_size;  /* arc length of this step (the final step onto a bound: the parameter increment) */
  double dparam_ds;  /* parameter component of the unit tangent at the new point */
  double scale;      /* parameter scale factor s in force after this step (1 without NOSH_ARC_SCALING) */
} nosh_arclength_step;
NOSH_API nosh_status nosh_continuation_arclength(nosh_ctx *ctx, int np, const char *const *names,
                                                 const double *values, const char *pname,
                                                 const nosh_arclength_options *opt, double *psi,
                                                 nosh_arclength_step *steps, int32_t *n_records);

/* ---- mesh files and vertex ordering ("next" row f3; host-only, no GPU needed) -------------------
 * nosh::read (src/mesh_reader.cpp:19-162) + the vertex tags the drivers take from the file --
 * mesh::get_complex_vector("psi"), get_vector("V"), get_multi_vector("A") (src/mesh.cpp:249-446) -- and
 * mesh::write for the outNNNN state dumps (src/mesh.cpp:249-263, src/continuation_data_saver.hpp:24-50).
 * The reference goes through MOAB (.h5m / Exodus); MOAB, HDF5 and netCDF are not available here.  Read: the legacy
 * VTK unstructured grid, ASCII or BINARY (what `meshio-convert in.e out.vtk` writes; MOAB reads it too), and
 * Exodus II (.e .exo .ex2 .g .gen) in the netCDF CLASSIC container (CDF-1 / CDF-2 / CDF-5; own reader of the
 * container) -- nodal variables of the last time step, X_R / X_Z joined into the complex tag X, X_X / X_Y / X_Z
 * into the vector tag X -- and gmsh MSH (.msh, ASCII, formats 2.x and 4.1: nodes, 3-node triangles, 4-node
 * tetrahedra, $NodeData views).  Written: legacy VTK.  Triangles / tetrahedra only, the highest-dimensional kind
 * present; vertices no kept cell uses are dropped (MSH).
 * .h5m and Exodus files in the netCDF-4 (HDF5) container return NOSH_EUNSUPPORTED.
 * Errors: nosh_meshfile_last_error(). */
typedef struct nosh_meshfile nosh_meshfile;
NOSH_API const char *nosh_meshfile_last_error(void);
NOSH_API nosh_status nosh_meshfile_read(const char *path, nosh_meshfile **out);
NOSH_API void nosh_meshfile_free(nosh_meshfile *m);
NOSH_API nosh_status nosh_meshfile_info(const nosh_meshfile *m, int32_t *dim, int64_t *n_vertices,
                                        int64_t *n_cells, int32_t *n_fields);
NOSH_API nosh_status nosh_meshfile_get(const nosh_meshfile *m, double *coords /* n x 3 */,
                                       int32_t *cells /* n_cells x (dim+1) */);
NOSH_API nosh_status nosh_meshfile_field_name(const nosh_meshfile *m, int32_t index, const char **name,
                                              int32_t *ncomp);
/* vertex tag by name (NOSH_EKEY if absent); values may be NULL to query ncomp */
NOSH_API nosh_status nosh_meshfile_get_field(const nosh_meshfile *m, const char *name, int32_t *ncomp,
                                             double *values /* n x ncomp */);
NOSH_API nosh_status nosh_meshfile_write(const char *path, int32_t dim, int64_t n_vertices,
                                         const double *coords, int64_t n_cells, const int32_t *cells,
                                         int32_t n_fields, const char *const *names, const int32_t *ncomps,
                                         const double *const *values, int32_t binary);
/* Spatially local vertex numbering for the contiguous-range partition (the stand-in for the `mbpart`
 * step of test/data/CMakeLists.txt:36-52): perm[i] = old id of the vertex that gets new id i along a
 * Morton curve through the bounding box. */
NOSH_API nosh_status nosh_morton_order(int64_t n_vertices, const double *coords /* n x 3 */, int64_t *perm);

/* ---- generic finite-volume matrix / operator ("next" row f4): nosh::fvm_matrix::fill (src/fvm_matrix.hpp:44-70)
 * and nosh::fvm_operator::apply (src/fvm_operator.hpp:45-94) for REAL scalar problems on the mesh's vertex graph
 * (src/mesh.cpp build_graph), enough for examples/poisson and examples/bratu.  The reference assembles from
 * user-defined virtual cores generated by nfc; a device cannot call host virtuals per edge, so:
 *   - edge_lhs == NULL: the built-in edge core of  integrate(-n_dot_grad(u), dS):
 *       covolume/length * edge_coeff[e] * [[1,-1],[-1,1]]   (edge_coeff == NULL: 1)
 *   - edge_lhs != NULL: arbitrary cores, evaluated once on the host: E x 4 doubles (lhs00, lhs01, lhs10, lhs11 of
 *       matrix_core_edge::eval, vertex 0 = the edge's first vertex in nosh_mesh_get_edges) and edge_rhs E x 2
 *   - vertex_lhs / vertex_rhs: n_owned doubles each, what matrix_core_vertex::eval returns (already times the
 *       control volume) -- boundary cores are added into the same arrays by the caller
 *   - dirichlet_mask (n_owned int32, non-zero = Dirichlet vertex) + dirichlet_values: rows replaced by unit rows,
 *       right-hand side = value; columns are NOT eliminated (as the reference, :208-250)
 * All arrays are HOST arrays (or NULL).  rhs (or NULL): n_owned doubles out, host or device.
 * Vectors of the apply / solve calls: n_owned doubles, host or device.  One rank only. */
typedef enum { NOSH_FVM_VERTEX_NONE = 0,
               NOSH_FVM_VERTEX_EXP = 1,            /* y_k -= alpha c_k exp(x_k)          (bratu.py: F)        */
               NOSH_FVM_VERTEX_EXP_LINEARIZED = 2  /* y_k -= alpha c_k exp(u0_k) x_k     (bratu.py: Jacobian) */
} nosh_fvm_vertex_core;
typedef enum { NOSH_FVM_DIRICHLET_NONE = 0,
               NOSH_FVM_DIRICHLET_IDENTITY = 1, /* y_k = x_k            (lambda u, x: u(x))      */
               NOSH_FVM_DIRICHLET_ZERO = 2,     /* y_k = 0              (lambda x, u: 0.0)       */
               NOSH_FVM_DIRICHLET_VALUE = 3     /* y_k = x_k - value_k  (residual of u = g)      */
} nosh_fvm_dirichlet_kind;
/* the mesh's "boundary" subdomain (src/mesh.cpp:76-131): flags[k] = 1 for owned vertices on the skin */
NOSH_API nosh_status nosh_mesh_boundary_vertices(nosh_ctx *ctx, int32_t *flags /* host, n_owned */);
NOSH_API nosh_status nosh_fvm_matrix_fill(nosh_ctx *ctx, const double *edge_coeff, const double *edge_lhs,
                                          const double *edge_rhs, const double *vertex_lhs, const double *vertex_rhs,
                                          const int32_t *dirichlet_mask, const double *dirichlet_values, double *rhs);
NOSH_API nosh_status nosh_fvm_matrix_apply(nosh_ctx *ctx, const double *x, double *y);
/* entry-wise access for parity: CSR of the n_owned x n_owned real matrix (rowptr n_owned+1, cols/vals n_blocks) */
NOSH_API nosh_status nosh_fvm_get_csr(nosh_ctx *ctx, int64_t *rowptr, int32_t *cols, double *vals);
/* fvm_operator::apply: y = [A x] + vertex core, Dirichlet rows last.  with_matrix = 0 drops the matrix term
 * (bratu.py: dFdp = -integrate(exp(u), dV) is NOSH_FVM_VERTEX_EXP with alpha = 1 and no matrix).  The matrix is the
 * one of the last nosh_fvm_matrix_fill (fill it WITHOUT Dirichlet rows for an operator: they are applied here). */
NOSH_API nosh_status nosh_fvm_operator_apply(nosh_ctx *ctx, int with_matrix, nosh_fvm_vertex_core vertex_core, double alpha,
                                             const double *u0, const int32_t *dirichlet_mask,
                                             nosh_fvm_dirichlet_kind dirichlet_kind, const double *dirichlet_values,
                                             const double *x, double *y);
/* Belos "Pseudo Block CG" on the filled matrix (examples/poisson/poisson.cpp:58-66), ||r|| / ||r0|| <= tol.
 * x0 = the Dirichlet lift of the fill (g on the Dirichlet vertices, 0 elsewhere) instead of poisson.cpp:28's 0:
 * the reference eliminates Dirichlet ROWS only, so its matrix is not symmetric; from the lift CG runs on the
 * symmetric interior block (from 0 plain CG does not converge -- the reference leans on its MueLu preconditioner). */
NOSH_API nosh_status nosh_fvm_cg(nosh_ctx *ctx, const double *b, double *x, double tol, int maxit,
                                 nosh_krylov_result *res);

/* ---- measurement helpers: device-resident scratch vectors so that benchmarks can
 * time kernels with inputs already in HBM.  slot in [0,8). Returns a device pointer to
 * 2*(n_owned+n_ghost) doubles owned by the ctx. */
NOSH_API nosh_status nosh_scratch_vector(nosh_ctx *ctx, int slot, double **dev_ptr);
/* tuning knobs for measurements (profiles/apply_variants.py): key "apply_variant" = which compiled variant
 * of the SELL-32 apply kernel the MINRES loop uses (0 = default).  Results do not depend on it.
 * Other keys (all leave results unchanged unless stated): "persistent_minres", "persistent_mgpu", "mgpu_lean",
 * "mgpu_fence" (loop schedules), "sell_sigma" (row-length sorting window; before the mesh is set), "amg_graph"
 * (V-cycle replayed as a CUDA graph), "amg_panel_products" (set-up panel size), and
 * "amg_mixed" = 1: the V-cycle's finest-level smoother and transfer operators read fp32 copies of K and P -- a
 * DIFFERENT (fp32-close, still symmetric positive definite) preconditioner, ~15 % faster per iteration; the
 * solvers' own operator, all vectors and the convergence test stay fp64.  Off by default. */
NOSH_API nosh_status nosh_ctx_set_tuning(nosh_ctx *ctx, const char *key, int value);
/* number of kernels this library has launched on ctx so far */
NOSH_API int64_t nosh_launch_count(const nosh_ctx *ctx);
/* CUDA-event timing on the ctx stream */
NOSH_API nosh_status nosh_timer_start(nosh_ctx *ctx);
NOSH_API nosh_status nosh_timer_stop(nosh_ctx *ctx, float *milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* NOSH_B200_H */
