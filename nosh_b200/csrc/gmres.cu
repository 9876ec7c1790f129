// gmres.cu -- restarted GMRES ("next" row f1 of SURVEY.md section 8): the solver examples/conf.xml:104
// actually selects ("Pseudo Block GMRES", tolerance 1e-10, 1000 iterations, one right-hand side).
// Belos is not in the reference tree (unpinned); restated (oracle/gmres.py) is GMRES(m) with the iterated
// classical Gram-Schmidt orthogonalisation Belos uses by default ("ICGS", two passes), Givens rotations
// and an optional RIGHT preconditioner (how the Stratimikos/Belos adapter installs an unspecified-side
// preconditioner).
//
// Device side, per Arnoldi step j (basis V_0..V_j of Nl-entry vectors, all resident in HBM):
//   w = A (M v_j)                                   fused apply (apply.cu)
//   2 x { h = V^T w   : k_multi_dot   -- one pass over V, one warp per basis vector and 512-vertex chunk,
//                                        fixed lane/warp order, then the fixed three-level tree per entry
//         w -= V h    : k_multi_axpy  -- one pass over V; the second pass also emits ||w||^2 partials }
//   v_{j+1} = w / ||w||
//   Hessenberg column, Givens rotations, residual estimate, convergence test: k_gmres_givens (one thread);
//   it sets KrylovState::done, which turns every later launch of the cycle into a no-op -- the same
//   device-side control flow as CG.  The host polls the state every 4, 8, ... 32 steps and once per restart
//   cycle; the least-squares solve and x += M (V y) at the end of a cycle run on the device too.
// Reductions use the tree of common.cuh, so the iteration is bit-identical for any number of GPUs (group sums
// cross ranks in one all-reduce).
#include <cmath>
#include <vector>

#include "amg.h"
#include "apply.cuh"
#include "comm.h"
#include "krylov.h"
#include "reduce.cuh"

namespace nosh {

namespace {

#define GLAUNCH(ctx, kernel, grid, block, ...)                  \
  do {                                                          \
    kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                          \
    CUDA_CHECK(cudaGetLastError());                             \
  } while (0)

// partials[k * n_chunks + chunk] = <V_k, w> over the chunk's 512 vertices; k = 0..nv-1.
// CTA = one chunk, 8 warps; warp q handles k = q, q+8, ...; lane l sums vertices l, l+32, ... in order.
__global__ void __launch_bounds__(256) k_multi_dot(const double2 *V, int64_t ldv, int nv, const double2 *w, int64_t No,
                                                   int64_t n_chunks, double *partials, const KrylovState *gate) {
  __shared__ double2 ws[CHUNK];
  if (gate && gate->done) return;
  const int64_t base = (int64_t)blockIdx.x * CHUNK;
  for (int t = threadIdx.x; t < CHUNK; t += 256) ws[t] = base + t < No ? w[base + t] : make_double2(0.0, 0.0);
  __syncthreads();
  const int q = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int k = q; k < nv; k += 8) {
    const double2 *v = V + (int64_t)k * ldv + base;
    double s = 0.0;
#pragma unroll 4
    for (int t = l; t < CHUNK; t += 32) {
      if (base + t < No) {
        const double2 a = ld_stream2(v + t), b = ws[t];
        s += a.x * b.x + a.y * b.y;
      }
    }
    s = warp_sum(s);
    if (l == 0) partials[(int64_t)k * n_chunks + blockIdx.x] = s;
  }
}

// one CTA per entry k: levels 2 and 3 of the fixed tree over that entry's chunk partials
struct MultiFin {
  const double *partials;  // nv x n_chunks
  int64_t n_chunks;
  int cpg;
  int64_t group_begin, n_groups_local;
  int n_groups_global;
  double *gsums;        // nv x MAX_GROUPS (global group index; zero outside the local groups)
  const double *grecv;  // all-reduced copy (== gsums on one GPU)
  double *out;          // nv
  int stage;            // 0: one GPU; 1: level 2 only; 2: level 3 only
  const KrylovState *gate;
};
__global__ void __launch_bounds__(1024) k_multi_finalize(const MultiFin M) {
  __shared__ double sm[32];
  if (M.gate && M.gate->done) return;
  const int k = blockIdx.x;
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const double *part = M.partials + (int64_t)k * M.n_chunks;
  double *gs = M.gsums + (int64_t)k * MAX_GROUPS;
  if (M.stage != 2) {
    const int per = (M.cpg + 31) / 32;
    for (int64_t g = w; g < M.n_groups_local; g += nw) {
      const int64_t base = g * M.cpg;
      double s = 0.0;
      for (int t = 0; t < per; t++) {
        const int c = l * per + t;
        if (c < M.cpg && base + c < M.n_chunks) s += part[base + c];
      }
      s = warp_sum(s);
      if (l == 0) gs[M.group_begin + g] = s;
    }
    if (M.stage == 1) return;
    __syncthreads();
  }
  const double *src = M.stage == 2 ? M.grecv + (int64_t)k * MAX_GROUPS : gs;
  for (int seg = w; seg < 32; seg += nw) {
    const int i = 32 * seg + l;
    double v = i < M.n_groups_global ? src[i] : 0.0;
    v = warp_sum(v);
    if (l == 0) sm[seg] = v;
  }
  __syncthreads();
  if (w == 0) {
    double t = sm[l];
    t = warp_sum(t);
    if (l == 0) M.out[k] = t;
  }
}

// w -= sum_k h_k V_k ; optionally partials[chunk] = sum |w_new|^2 over the chunk (fixed order)
template <bool NORM>
__global__ void __launch_bounds__(TPB) k_multi_axpy(const double2 *V, int64_t ldv, int nv, const double *h, double2 *w,
                                                    int64_t No, double *partials, const KrylovState *gate) {
  __shared__ double hs[512];
  __shared__ double red[32];
  if (gate && gate->done) return;
  for (int t = threadIdx.x; t < nv; t += TPB) hs[t] = h[t];
  __syncthreads();
  double c = 0.0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + half * TPB;
    if (i < No) {
      double2 a = w[i];
      for (int k = 0; k < nv; k++) {
        const double2 v = ld_stream2(V + (int64_t)k * ldv + i);
        a.x -= hs[k] * v.x;
        a.y -= hs[k] * v.y;
      }
      w[i] = a;
      if (NORM) c += a.x * a.x + a.y * a.y;
    }
  }
  if (NORM) {
    const double s = block_sum<TPB / 32>(c, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
  }
}
// out = a * x (+ y); a_dev (optional): the scale factor is read from device memory
__global__ void k_scale_add(double a, const double *a_dev, const double2 *x, const double2 *y, double2 *out, int64_t n,
                            const KrylovState *gate) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || (gate && gate->done)) return;
  if (a_dev) a = *a_dev;
  double2 r = make_double2(a * x[i].x, a * x[i].y);
  if (y) {
    r.x += y[i].x;
    r.y += y[i].y;
  }
  out[i] = r;
}
// out = sum_{k < *nv_dev} y_k V_k
__global__ void __launch_bounds__(TPB) k_combine(const double2 *V, int64_t ldv, const int *nv_dev, const double *y,
                                                 double2 *out, int64_t No) {
  __shared__ double ys[512];
  const int nv = *nv_dev;
  for (int t = threadIdx.x; t < nv; t += TPB) ys[t] = y[t];
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)TPB + threadIdx.x;
  if (i >= No) return;
  double2 a = make_double2(0.0, 0.0);
  for (int k = 0; k < nv; k++) {
    const double2 v = V[(int64_t)k * ldv + i];
    a.x += ys[k] * v.x;
    a.y += ys[k] * v.y;
  }
  out[i] = a;
}

// The small dense part of GMRES(m), resident on the device: Hessenberg matrix (row-major, (m+1) x m, holds R after
// the rotations), rotations, right-hand side g, solution y.
struct GmresDense {
  double *H, *cs, *sn, *g, *y;
  double *h1, *h2, *nn;  // this step's Gram-Schmidt coefficients (two passes) and ||w||^2
  double *scal;          // [0]: 1 / (norm of the vector to normalise next); [1]: ||r_0||
  int *ncol;             // Arnoldi steps taken in the current cycle
  double *hist;
  int m;
};
// start of a restart cycle: beta = ||r|| from *G.nn; g = beta e_0
__global__ void k_gmres_begin_cycle(KrylovState *st, const GmresDense G, int first, double tol, int maxit) {
  if (threadIdx.x || blockIdx.x) return;
  const double beta = sqrt(*G.nn);
  if (first) {
    KrylovState z = {};
    *st = z;
    st->tol = tol;
    st->maxit = maxit;
    st->r0norm = beta;
    st->relres = beta == 0.0 ? 0.0 : 1.0;
    G.scal[1] = beta;
    G.hist[0] = 1.0;
    if (beta == 0.0 || maxit == 0) {
      st->converged = beta == 0.0;
      st->done = 1;
    }
  }
  if (beta == 0.0) st->done = 1;  // exact solution at a restart
  for (int i = 0; i <= G.m; i++) G.g[i] = 0.0;
  G.g[0] = beta;
  G.scal[0] = beta > 0.0 ? 1.0 / beta : 0.0;
  *G.ncol = 0;
}
// step j: Hessenberg column from the two Gram-Schmidt passes, previous rotations, new rotation, residual estimate
__global__ void k_gmres_givens(KrylovState *st, const GmresDense G, int j) {
  if (threadIdx.x || blockIdx.x || st->done) return;
  const int m = G.m;
  const double hn = sqrt(*G.nn);
  for (int i = 0; i <= j; i++) G.H[(size_t)i * m + j] = G.h1[i] + G.h2[i];
  G.H[(size_t)(j + 1) * m + j] = hn;
  for (int i = 0; i < j; i++) {
    const double a = G.H[(size_t)i * m + j], c = G.H[(size_t)(i + 1) * m + j];
    G.H[(size_t)i * m + j] = G.cs[i] * a + G.sn[i] * c;
    G.H[(size_t)(i + 1) * m + j] = -G.sn[i] * a + G.cs[i] * c;
  }
  const double a = G.H[(size_t)j * m + j], d = hypot(a, hn);
  G.cs[j] = d == 0.0 ? 1.0 : a / d;
  G.sn[j] = d == 0.0 ? 0.0 : hn / d;
  G.H[(size_t)j * m + j] = d;
  G.H[(size_t)(j + 1) * m + j] = 0.0;
  G.g[j + 1] = -G.sn[j] * G.g[j];
  G.g[j] = G.cs[j] * G.g[j];
  st->iter++;
  st->relres = fabs(G.g[j + 1]) / G.scal[1];
  G.hist[st->iter] = st->relres;
  *G.ncol = j + 1;
  if (st->relres <= st->tol || hn == 0.0) {  // hn == 0: happy breakdown, the Krylov space is invariant
    st->converged = 1;
    st->done = 1;
  } else if (st->iter >= st->maxit) {
    st->done = 1;
  }
  G.scal[0] = hn > 0.0 ? 1.0 / hn : 0.0;
}
// end of a cycle: y = R^-1 g over the *G.ncol columns taken (once per cycle; the oracle's summation order)
__global__ void k_gmres_solve_y(const GmresDense G) {
  if (threadIdx.x || blockIdx.x) return;
  const int k = *G.ncol, m = G.m;
  for (int i = k - 1; i >= 0; i--) {
    double s = G.g[i];
    for (int c = i + 1; c < k; c++) s -= G.H[(size_t)i * m + c] * G.y[c];
    G.y[i] = s / G.H[(size_t)i * m + i];
  }
}

int poll_state(Ctx *ctx, KrylovState *host) {
  CUDA_CHECK(cudaMemcpyAsync(host, ctx->kstate.p, sizeof(KrylovState), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return host->done;
}

struct Reducer {
  Ctx *ctx;
  DBuf<double> partials, gsums, grecv;
  void init(Ctx *c, int nvmax) {
    ctx = c;
    const int64_t nch = c->n_chunks > 0 ? c->n_chunks : 1;
    partials.alloc((size_t)nvmax * nch);
    gsums.alloc((size_t)nvmax * MAX_GROUPS);
    grecv.alloc((size_t)nvmax * MAX_GROUPS);
    CUDA_CHECK(cudaMemsetAsync(gsums.p, 0, sizeof(double) * nvmax * MAX_GROUPS, c->stream));
  }
  // finishes the nv reductions whose chunk partials are in `partials`; results go to out[0..nv) (device)
  void finalize(int nv, double *out, const KrylovState *gate) {
    MultiFin M;
    M.partials = partials.p;
    M.n_chunks = ctx->n_chunks;
    M.cpg = ctx->chunks_per_group;
    M.group_begin = ctx->group_begin;
    M.n_groups_local = ctx->n_groups_local;
    M.n_groups_global = (int)ctx->n_groups_global;
    M.gsums = gsums.p;
    M.grecv = grecv.p;
    M.out = out;
    M.gate = gate;
    if (ctx->nranks == 1) {
      M.stage = 0;
      GLAUNCH(ctx, k_multi_finalize, nv, 1024, M);
    } else {
      // the all-reduce is a collective: every rank makes it, converged or not (the values are then unused)
      M.stage = 1;
      GLAUNCH(ctx, k_multi_finalize, nv, 1024, M);
      comm_allreduce_sum(ctx, gsums.p, grecv.p, (int64_t)nv * MAX_GROUPS);
      M.stage = 2;
      GLAUNCH(ctx, k_multi_finalize, nv, 1024, M);
    }
  }
};

}  // namespace

void gmres_dev(Ctx *ctx, int op, int prec, const double2 *b, double bscale, double2 *x_out, double tol, int maxit,
               int restart, nosh_krylov_result *res, double *hist_host) {
  ensure_work(ctx);
  if (maxit < 0) NOSH_THROW(NOSH_EINVAL, "maxit < 0");
  if (restart < 1 || restart > 500) NOSH_THROW(NOSH_EINVAL, "restart length must be in [1, 500]");
  if (prec != NOSH_PREC_NONE && prec != NOSH_PREC_KEOREG_AMG) NOSH_THROW(NOSH_EINVAL, "unknown preconditioner %d", prec);
  const bool pc = prec != NOSH_PREC_NONE;
  if (pc) amg_ensure(ctx);
  const int m = restart < maxit ? restart : (maxit > 0 ? maxit : 1);
  const int64_t No = ctx->No, ld = ctx->Nl > 0 ? ctx->Nl : 1;
  const unsigned gch = (unsigned)ctx->n_chunks, g1 = (unsigned)cdiv(No > 0 ? No : 1, 256);
  ctx->gmres_basis.ensure((size_t)(m + 1) * ld);
  double2 *V = ctx->gmres_basis.p;
  CUDA_CHECK(cudaMemsetAsync(V, 0, sizeof(double2) * (size_t)(m + 1) * ld, ctx->stream));
  double2 *W = ctx->work[2].p, *Z = ctx->work[0].p, *T = ctx->work[1].p, *X = x_out;
  Reducer R;
  R.init(ctx, m + 2);
  ctx->hist.ensure((size_t)maxit + 2);
  // the dense part: one allocation
  DBuf<double> dense;
  const size_t nH = (size_t)(m + 1) * m;
  dense.alloc(nH + 7 * (size_t)(m + 2) + 8);
  CUDA_CHECK(cudaMemsetAsync(dense.p, 0, sizeof(double) * dense.n, ctx->stream));
  GmresDense G;
  G.m = m;
  G.H = dense.p;
  G.cs = G.H + nH;
  G.sn = G.cs + (m + 2);
  G.g = G.sn + (m + 2);
  G.y = G.g + (m + 2);
  G.h1 = G.y + (m + 2);
  G.h2 = G.h1 + (m + 2);
  G.nn = G.h2 + (m + 2);
  G.scal = G.nn + 2;
  G.ncol = (int *)(G.scal + 4);
  G.hist = ctx->hist.p;
  KrylovState *st = ctx->kstate.p;
  const double2 *nil = nullptr;
  if (No) CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(double2) * No, ctx->stream));
  // r0 = bscale * b   (x0 = 0)
  if (No) GLAUNCH(ctx, k_scale_add, g1, 256, bscale, (const double *)nullptr, b, nil, W, No, (const KrylovState *)nullptr);
  auto norm2_to_nn = [&](const double2 *v) {
    if (gch) GLAUNCH(ctx, k_multi_dot, gch, 256, v, ld, 1, v, No, ctx->n_chunks, R.partials.p, (const KrylovState *)nullptr);
    R.finalize(1, G.nn, nullptr);
  };
  norm2_to_nn(W);
  GLAUNCH(ctx, k_gmres_begin_cycle, 1, 32, st, G, 1, tol, maxit);
  KrylovState hs;
  poll_state(ctx, &hs);
  int check = 4, since_poll = 0;
  while (!hs.done) {
    // v_0 = r / beta
    if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 0.0, (const double *)G.scal, W, nil, V, No, (const KrylovState *)nullptr);
    const int left = maxit - hs.iter;
    for (int j = 0; j < m && j < left; j++) {
      double2 *vj = V + (int64_t)j * ld;
      // w = A (M v_j)
      if (pc) {
        {  // every basis vector is a different source pointer: not worth one captured graph each
          struct NoGraph {
            Ctx *c;
            int keep;
            ~NoGraph() { c->amg_graph = keep; }
          } ng{ctx, ctx->amg_graph};
          ctx->amg_graph = 0;
          amg_vcycle(ctx, vj, Z, st);
        }
        apply_op_gated_dev(ctx, op, Z, W, st);
      } else {
        apply_op_gated_dev(ctx, op, vj, W, st);
      }
      // ICGS, two passes
      const int nv = j + 1;
      if (gch) GLAUNCH(ctx, k_multi_dot, gch, 256, V, ld, nv, W, No, ctx->n_chunks, R.partials.p, st);
      R.finalize(nv, G.h1, st);
      if (gch) GLAUNCH(ctx, k_multi_axpy<false>, gch, TPB, V, ld, nv, G.h1, W, No, (double *)nullptr, st);
      if (gch) GLAUNCH(ctx, k_multi_dot, gch, 256, V, ld, nv, W, No, ctx->n_chunks, R.partials.p, st);
      R.finalize(nv, G.h2, st);
      if (gch) GLAUNCH(ctx, k_multi_axpy<true>, gch, TPB, V, ld, nv, G.h2, W, No, R.partials.p, st);
      R.finalize(1, G.nn, st);
      GLAUNCH(ctx, k_gmres_givens, 1, 32, st, G, j);
      // v_{j+1} = w / ||w||  (a no-op once the step above has set done)
      if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 0.0, (const double *)G.scal, W, nil, V + (int64_t)(j + 1) * ld, No, st);
      if (++since_poll >= check) {
        since_poll = 0;
        if (check < 32) check *= 2;
        if (poll_state(ctx, &hs)) break;
      }
    }
    // y = R^-1 g over the columns taken, x += M (V y): always, converged or not
    GLAUNCH(ctx, k_gmres_solve_y, 1, 32, G);
    if (No) GLAUNCH(ctx, k_combine, g1, TPB, V, ld, (const int *)G.ncol, (const double *)G.y, T, No);
    if (pc) {
      amg_vcycle(ctx, T, Z, nullptr);
      if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 1.0, (const double *)nullptr, Z, (const double2 *)X, X, No, (const KrylovState *)nullptr);
    } else {
      if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 1.0, (const double *)nullptr, T, (const double2 *)X, X, No, (const KrylovState *)nullptr);
    }
    if (poll_state(ctx, &hs)) break;
    since_poll = 0;
    // restart: explicit residual r = b - A x
    if (No) CUDA_CHECK(cudaMemcpyAsync(T, X, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
    apply_op_dev(ctx, op, T, W);
    if (No) GLAUNCH(ctx, k_scale_add, g1, 256, -1.0, (const double *)nullptr, W, nil, W, No, (const KrylovState *)nullptr);
    if (No) GLAUNCH(ctx, k_scale_add, g1, 256, bscale, (const double *)nullptr, b, (const double2 *)W, W, No, (const KrylovState *)nullptr);
    norm2_to_nn(W);
    GLAUNCH(ctx, k_gmres_begin_cycle, 1, 32, st, G, 0, tol, maxit);
    poll_state(ctx, &hs);
  }
  if (p2p_check_error(ctx)) NOSH_THROW(NOSH_ECOMM, "peer-memory exchange timed out (a rank is not responding)");
  if (res) {
    res->iterations = hs.iter;
    res->converged = hs.converged;
    res->relres = hs.relres;
    res->breakdown = 0;
  }
  if (hist_host) {
    CUDA_CHECK(cudaMemcpyAsync(hist_host, ctx->hist.p, sizeof(double) * (hs.iter + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  // the Arnoldi basis is (restart+1) full vectors (38 GB for GMRES(300) on 8.0M vertices): keep a small one
  // for the next solve, give a large one back
  if (ctx->gmres_basis.n * sizeof(double2) > ((size_t)4 << 30)) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->gmres_basis.release();
  }
}

}  // namespace nosh
