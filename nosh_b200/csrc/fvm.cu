// fvm.cu -- generic real-valued finite-volume matrix / operator on the device ("next" row f4 of SURVEY.md 8):
// the fixed cores that carry the reference's examples/poisson and examples/bratu.
//
//   fvm_matrix::fill   src/fvm_matrix.hpp:44-70    edge + vertex (+ boundary) contributions, then Dirichlet rows
//   fvm_operator::apply src/fvm_operator.hpp:45-94 operators + edge + vertex contributions, Dirichlet last
//   boundary vertices  src/mesh.cpp:76-131         (MOAB Skinner there; here: solid-angle deficit)
//
// The reference assembles these from user-defined virtual eval(edge) / eval(vertex) cores that nfc generates
// from sympy (nfc/templates/*.tpl).  A device cannot call host virtuals per edge, so the boundary here is:
//   * built-in cores evaluated on the device: the edge core of  integrate(-n_dot_grad(u), dS)
//     = covolume/length * [[1,-1],[-1,1]]  (nfc/nfc/discretize_edge_integral.py:108-114) scaled by an optional
//     per-edge coefficient, and vertex cores  control_volume * a_k  /  control_volume * f_k  from per-vertex arrays;
//   * arbitrary cores evaluated ONCE by the host (the C++ mirror loops over edges and calls eval()) into
//     per-edge 2x2 blocks / per-vertex pairs that are scattered here -- same slots, no atomics.
// The matrix shares the complex KEO's graph (one entry per vertex pair + diagonal, src/mesh.cpp:build_graph) and
// its SELL-32(-sigma) storage positions; values are one double per slot.  One rank only (NOSH_EUNSUPPORTED on
// several): the examples are single-field real problems and the halo machinery moves complex entries.
#include "fvm.h"

#include <cmath>

#include "krylov.h"

namespace nosh {

namespace {

inline dim3 grid_for(int64_t n, int tpb = 256) { return dim3((unsigned)cdiv(n > 0 ? n : 1, tpb)); }
#define FLAUNCH(ctx, kernel, n, ...)                             \
  do {                                                           \
    kernel<<<grid_for(n), 256, 0, (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                           \
    CUDA_CHECK(cudaGetLastError());                              \
  } while (0)

// ---- boundary vertices: the angles (2D) / solid angles (3D) of the cells around an interior vertex add up to
// 2 pi / 4 pi; on the skin they do not.  One thread per cell, atomic adds (a flag, not a parity-relevant sum).
__device__ __forceinline__ double3 sub3(const double *c, int a, int b) {
  return make_double3(c[3 * a] - c[3 * b], c[3 * a + 1] - c[3 * b + 1], c[3 * a + 2] - c[3 * b + 2]);
}
__device__ __forceinline__ double dot3(double3 a, double3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double3 cross3(double3 a, double3 b) {
  return make_double3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double norm3(double3 a) { return sqrt(dot3(a, a)); }

template <int NVC>
__global__ void k_angle_sums(const double *coords, const int32_t *cells, int64_t nc, int64_t No, double *sum) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int v[NVC];
#pragma unroll
  for (int i = 0; i < NVC; i++) v[i] = cells[c * NVC + i];
#pragma unroll
  for (int i = 0; i < NVC; i++) {
    if (v[i] >= No) continue;
    double w;
    if (NVC == 3) {
      const double3 a = sub3(coords, v[(i + 1) % 3], v[i]), b = sub3(coords, v[(i + 2) % 3], v[i]);
      w = atan2(norm3(cross3(a, b)), dot3(a, b));
    } else {
      // Van Oosterom & Strackee: tan(Omega/2) = |a.(b x c)| / (|a||b||c| + (a.b)|c| + (a.c)|b| + (b.c)|a|)
      const double3 a = sub3(coords, v[(i + 1) % 4], v[i]), b = sub3(coords, v[(i + 2) % 4], v[i]),
                    d = sub3(coords, v[(i + 3) % 4], v[i]);
      const double la = norm3(a), lb = norm3(b), ld = norm3(d);
      const double num = fabs(dot3(a, cross3(b, d)));
      const double den = la * lb * ld + dot3(a, b) * ld + dot3(a, d) * lb + dot3(b, d) * la;
      w = 2.0 * atan2(num, den);
    }
    atomicAdd(sum + v[i], w);
  }
}
__global__ void k_boundary_flag(const double *sum, int64_t No, double full, int32_t *flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < No) flag[i] = sum[i] < full * (1.0 - 1e-6) ? 1 : 0;
}

// ---- fill -------------------------------------------------------------------------------------------------------
// off-diagonal entries: one thread per edge, both slots (storage positions of the KEO layout)
__global__ void k_fvm_edges(const double *elen, const double *ecov, const double *coeff /* E or NULL */,
                            const double *lhs /* E x 4 (00, 01, 10, 11) or NULL */, const int32_t *slot_ij,
                            const int32_t *slot_ji, int64_t E, double *val, double *ealpha /* E x 2: ii, jj parts */) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  double l00, l01, l10, l11;
  if (lhs) {
    l00 = lhs[4 * e];
    l01 = lhs[4 * e + 1];
    l10 = lhs[4 * e + 2];
    l11 = lhs[4 * e + 3];
  } else {
    const double a = (ecov[e] / elen[e]) * (coeff ? coeff[e] : 1.0);
    l00 = a;
    l01 = -a;
    l10 = -a;
    l11 = a;
  }
  const int sij = slot_ij[e], sji = slot_ji[e];
  if (sij >= 0) val[sij] = l01;
  if (sji >= 0) val[sji] = l10;
  ealpha[2 * e] = l00;
  ealpha[2 * e + 1] = l11;
}
// diagonal + right-hand side: one thread per owned row, contributions gathered in ascending CSR order
__global__ void k_fvm_rows(const int32_t *rowptr, const int32_t *edge_of, const int32_t *csr_col, const int32_t *edges,
                           const int32_t *diag_slot, const double *ealpha, const double *erhs /* E x 2 or NULL */,
                           const double *vlhs, const double *vrhs, int64_t No, double *val, double *rhs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= No) return;
  double d = 0.0, r = 0.0;
  for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const int e = edge_of[p];
    if (e < 0) continue;
    const int side = edges[2 * e] == (int)i ? 0 : 1;  // am I vertex 0 or vertex 1 of this edge?
    d += ealpha[2 * e + side];
    if (erhs) r += erhs[2 * e + side];
  }
  if (vlhs) d += vlhs[i];
  if (vrhs) r += vrhs[i];
  val[diag_slot[i]] = d;
  rhs[i] = r;
}
// Dirichlet rows: the row becomes the unit row, the right-hand side the boundary value; the COLUMN stays
// (the reference eliminates rows only, src/fvm_matrix.hpp:208-250)
__global__ void k_fvm_dirichlet(const int32_t *rowptr, const int32_t *csr_pos, const int32_t *diag_slot,
                                const int32_t *mask, const double *dval, int64_t No, double *val, double *rhs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= No || !mask[i]) return;
  for (int p = rowptr[i]; p < rowptr[i + 1]; p++) val[csr_pos[p]] = 0.0;
  val[diag_slot[i]] = 1.0;
  rhs[i] = dval[i];
}

// ---- apply: y = A x (+ vertex core of the operator) with the Dirichlet override, one thread per row ----------------
struct FvmApply {
  int64_t No, nslices;
  const int32_t *rowptr, *slice_off, *sell_row, *col;
  const double *val, *x;
  double *y;
  int sell, vertex_kind, dirichlet_kind;
  double alpha;
  const double *cv, *u0;
  const int32_t *mask;
  const double *dval;
};
__global__ void __launch_bounds__(256) k_fvm_apply(const FvmApply A) {
  const int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t row = pos;
  double acc = 0.0;
  if (A.val) {
    if (A.sell) {
      const int64_t slice = pos >> 5;
      if (slice >= A.nslices) return;
      row = A.sell_row ? (int64_t)A.sell_row[pos] : pos;
      const int pend = A.slice_off[slice + 1];
      for (int p = A.slice_off[slice] + (int)(pos & 31); p < pend; p += 32) acc = fma(A.val[p], __ldg(A.x + A.col[p]), acc);
    } else if (row < A.No) {
      for (int p = A.rowptr[row]; p < A.rowptr[row + 1]; p++) acc = fma(A.val[p], __ldg(A.x + A.col[p]), acc);
    }
  }
  if (row >= A.No) return;
  const double xi = A.x[row];
  switch (A.vertex_kind) {  // operator_core_vertex::eval (nfc/templates/operator_core_vertex.tpl)
    case NOSH_FVM_VERTEX_EXP: acc -= A.alpha * A.cv[row] * exp(xi); break;                      // - int alpha e^u dV
    case NOSH_FVM_VERTEX_EXP_LINEARIZED: acc -= A.alpha * A.cv[row] * exp(A.u0[row]) * xi; break;  // its derivative
    default: break;
  }
  if (A.mask && A.mask[row]) {  // Dirichlet comes at the end, overriding everything (fvm_operator.hpp:91-92)
    if (A.dirichlet_kind == NOSH_FVM_DIRICHLET_IDENTITY) acc = xi;
    else if (A.dirichlet_kind == NOSH_FVM_DIRICHLET_ZERO) acc = 0.0;
    else if (A.dirichlet_kind == NOSH_FVM_DIRICHLET_VALUE) acc = xi - A.dval[row];
  }
  A.y[row] = acc;
}

// ---- real CG (Belos "Pseudo Block CG", the solver of examples/poisson/poisson.cpp:60) ------------------------------
// deterministic two-stage reductions: fixed chunk partials, then one CTA in chunk order
__global__ void __launch_bounds__(256) k_rdot(const double *x, const double *y, int64_t n, double *partials) {
  __shared__ double red[8];
  const int64_t i0 = (int64_t)blockIdx.x * 512 + threadIdx.x, i1 = i0 + 256;
  double c = 0.0;
  if (i0 < n) c = x[i0] * y[i0];
  if (i1 < n) c += x[i1] * y[i1];
  const double s = block_sum<8>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_rsum(const double *partials, int64_t np, double *out) {
  __shared__ double red[32];
  double c = 0.0;
  for (int64_t i = threadIdx.x; i < np; i += 1024) c += partials[i];
  const double s = block_sum<32>(c, red);
  if (threadIdx.x == 0) out[0] = s;
}
__global__ void k_cg_xr(double al, const double *p, const double *ap, int64_t n, double *x, double *r) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    x[i] += al * p[i];
    r[i] -= al * ap[i];
  }
}
__global__ void k_lift(const int32_t *mask, const double *dval, int64_t n, double *x) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && mask[i]) x[i] = dval[i];
}
__global__ void k_residual(const double *b, const double *ax, int64_t n, double *r, double *p) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = b[i] - ax[i];
    r[i] = v;
    p[i] = v;
  }
}
__global__ void k_cg_p(double be, const double *r, int64_t n, double *p) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = r[i] + be * p[i];
}

void require_single(Ctx *ctx) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  if (ctx->nranks > 1) NOSH_THROW(NOSH_EUNSUPPORTED, "the generic FVM cores run on one rank");
}

double rdot(Ctx *ctx, const double *x, const double *y, int64_t n, DBuf<double> &partials, DBuf<double> &out) {
  const int64_t np = cdiv(n, 512);
  partials.ensure(np > 0 ? np : 1);
  out.ensure(1);
  if (np) {
    k_rdot<<<(unsigned)np, 256, 0, ctx->stream>>>(x, y, n, partials.p);
    ctx->launches++;
  }
  k_rsum<<<1, 1024, 0, ctx->stream>>>(partials.p, np, out.p);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
  double h;
  CUDA_CHECK(cudaMemcpyAsync(&h, out.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return h;
}

}  // namespace

void fvm_boundary_vertices(Ctx *ctx, int32_t *flags_dev) {
  require_single(ctx);
  DBuf<double> sum;
  sum.alloc(ctx->No);
  CUDA_CHECK(cudaMemsetAsync(sum.p, 0, sizeof(double) * ctx->No, ctx->stream));
  const double pi = 3.14159265358979323846;
  if (ctx->dim == 3)
    FLAUNCH(ctx, (k_angle_sums<4>), ctx->nc, ctx->coords.p, ctx->cells.p, ctx->nc, ctx->No, sum.p);
  else
    FLAUNCH(ctx, (k_angle_sums<3>), ctx->nc, ctx->coords.p, ctx->cells.p, ctx->nc, ctx->No, sum.p);
  FLAUNCH(ctx, k_boundary_flag, ctx->No, sum.p, ctx->No, ctx->dim == 3 ? 4.0 * pi : 2.0 * pi, flags_dev);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void fvm_matrix_fill(Ctx *ctx, const double *edge_coeff, const double *edge_lhs, const double *edge_rhs,
                     const double *vertex_lhs, const double *vertex_rhs, const int32_t *dmask, const double *dval) {
  require_single(ctx);
  const int64_t E = ctx->E, No = ctx->No;
  if (dmask && !dval) NOSH_THROW(NOSH_EINVAL, "Dirichlet mask without values");
  auto up = [&](DBuf<double> &b, const double *h, int64_t n) -> const double * {
    if (!h) return nullptr;
    b.alloc(n);
    CUDA_CHECK(cudaMemcpyAsync(b.p, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    return b.p;
  };
  DBuf<double> dcoef, dlhs, derhs, dvl, dvr, ealpha;
  const double *pc = up(dcoef, edge_coeff, E), *pl = up(dlhs, edge_lhs, 4 * E), *per = up(derhs, edge_rhs, 2 * E);
  const double *pvl = up(dvl, vertex_lhs, No), *pvr = up(dvr, vertex_rhs, No);
  ealpha.alloc(2 * E);
  ctx->fvm_val.ensure(ctx->nstored > 0 ? ctx->nstored : 1);
  ctx->fvm_rhs.ensure(No > 0 ? No : 1);
  CUDA_CHECK(cudaMemsetAsync(ctx->fvm_val.p, 0, sizeof(double) * ctx->nstored, ctx->stream));
  FLAUNCH(ctx, k_fvm_edges, E, ctx->elen.p, ctx->ecov.p, pc, pl, ctx->slot_ij.p, ctx->slot_ji.p, E, ctx->fvm_val.p,
          ealpha.p);
  FLAUNCH(ctx, k_fvm_rows, No, ctx->rowptr.p, ctx->edge_of.p, ctx->csr_col.p, ctx->edges.p, ctx->diag_slot.p, ealpha.p,
          per, pvl, pvr, No, ctx->fvm_val.p, ctx->fvm_rhs.p);
  ctx->fvm_mask.release();
  ctx->fvm_dval.release();
  if (dmask) {
    ctx->fvm_mask.alloc(No);
    ctx->fvm_dval.alloc(No);
    CUDA_CHECK(cudaMemcpyAsync(ctx->fvm_mask.p, dmask, sizeof(int32_t) * No, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->fvm_dval.p, dval, sizeof(double) * No, cudaMemcpyHostToDevice, ctx->stream));
    FLAUNCH(ctx, k_fvm_dirichlet, No, ctx->rowptr.p, ctx->csr_pos.p, ctx->diag_slot.p, ctx->fvm_mask.p, ctx->fvm_dval.p,
            No, ctx->fvm_val.p, ctx->fvm_rhs.p);
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the staging buffers go out of scope
  ctx->fvm_filled = true;
}

void fvm_apply_dev(Ctx *ctx, bool with_matrix, int vertex_kind, double alpha, const double *u0, const int32_t *mask,
                   int dirichlet_kind, const double *dval, const double *x, double *y) {
  require_single(ctx);
  if (with_matrix && !ctx->fvm_filled) NOSH_THROW(NOSH_ESTATE, "FVM matrix not filled (nosh_fvm_matrix_fill)");
  FvmApply A;
  memset(&A, 0, sizeof(A));
  A.No = ctx->No;
  A.nslices = ctx->nslices;
  A.rowptr = ctx->rowptr.p;
  A.slice_off = ctx->slice_off.p;
  A.sell_row = ctx->sell_permuted ? ctx->sell_row.p : nullptr;
  A.col = ctx->col.p;
  A.val = with_matrix ? ctx->fvm_val.p : nullptr;
  A.x = x;
  A.y = y;
  A.sell = ctx->layout == NOSH_LAYOUT_SELL32;
  A.vertex_kind = vertex_kind;
  A.dirichlet_kind = dirichlet_kind;
  A.alpha = alpha;
  A.cv = ctx->cv.p;
  A.u0 = u0;
  A.mask = mask;
  A.dval = dval;
  const int64_t n = (A.sell && with_matrix) ? ctx->nslices * 32 : ctx->No;
  if (n > 0) FLAUNCH(ctx, k_fvm_apply, n, A);
}

// Belos PseudoBlockCG organisation (oracle/fvm.py:cg).  x0 = the Dirichlet lift of the last fill (g on the
// Dirichlet vertices, 0 elsewhere; 0 everywhere without Dirichlet rows): the reference eliminates rows only, so
// the matrix is not symmetric, but from the lift the residual and all search directions vanish on those rows and
// CG runs on the symmetric interior block.  Stop on ||r|| / ||r0|| <= tol, r0 = b - A x0.
void fvm_cg_dev(Ctx *ctx, const double *b, double *x, double tol, int maxit, nosh_krylov_result *res) {
  require_single(ctx);
  if (!ctx->fvm_filled) NOSH_THROW(NOSH_ESTATE, "FVM matrix not filled (nosh_fvm_matrix_fill)");
  const int64_t n = ctx->No;
  DBuf<double> R, P, AP, partials, out;
  R.alloc(n);
  P.alloc(n);
  AP.alloc(n);
  CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * n, ctx->stream));
  if (ctx->fvm_mask.p) FLAUNCH(ctx, k_lift, n, ctx->fvm_mask.p, ctx->fvm_dval.p, n, x);
  fvm_apply_dev(ctx, true, NOSH_FVM_VERTEX_NONE, 0.0, nullptr, nullptr, NOSH_FVM_DIRICHLET_NONE, nullptr, x, AP.p);
  FLAUNCH(ctx, k_residual, n, b, AP.p, n, R.p, P.p);
  double rho = rdot(ctx, R.p, R.p, n, partials, out);
  const double r0 = sqrt(rho);
  int iter = 0;
  if (r0 > 0.0) {
    while (iter < maxit && !(sqrt(rho) / r0 <= tol)) {
      iter++;
      fvm_apply_dev(ctx, true, NOSH_FVM_VERTEX_NONE, 0.0, nullptr, nullptr, NOSH_FVM_DIRICHLET_NONE, nullptr, P.p, AP.p);
      const double pAp = rdot(ctx, P.p, AP.p, n, partials, out);
      const double al = rho / pAp;
      FLAUNCH(ctx, k_cg_xr, n, al, P.p, AP.p, n, x, R.p);
      const double rho_new = rdot(ctx, R.p, R.p, n, partials, out);
      const double be = rho_new / rho;
      rho = rho_new;
      FLAUNCH(ctx, k_cg_p, n, be, R.p, n, P.p);
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (res) {
    res->iterations = iter;
    res->relres = r0 > 0.0 ? sqrt(rho) / r0 : 0.0;
    res->converged = res->relres <= tol;
    res->breakdown = 0;
    res->reserved = 0;
  }
}

}  // namespace nosh
