// gmres.cu -- restarted GMRES ("next" row f1 of SURVEY.md section 8): the solver examples/conf.xml:104
// actually selects ("Pseudo Block GMRES", tolerance 1e-10, 1000 iterations, one right-hand side).
// Belos is not in the reference tree (unpinned); restated (oracle/gmres.py) is GMRES(m) with the iterated
// classical Gram-Schmidt orthogonalisation Belos uses by default ("ICGS", two passes), Givens rotations
// on the host and an optional RIGHT preconditioner (how the Stratimikos/Belos adapter installs an
// unspecified-side preconditioner).
//
// Device side, per Arnoldi step j (basis V_0..V_j of Nl-entry vectors, all resident in HBM):
//   w = A (M v_j)                                   fused apply (apply.cu)
//   2 x { h = V^T w   : k_multi_dot   -- one pass over V, one warp per basis vector and 512-vertex chunk,
//                                        fixed lane/warp order, then the fixed three-level tree per entry
//         w -= V h    : k_multi_axpy  -- one pass over V; the second pass also emits ||w||^2 partials }
//   v_{j+1} = w / ||w||
// One host synchronisation per step (the Hessenberg column).  Reductions use the tree of common.cuh, so
// the iteration is bit-identical for any number of GPUs (group sums cross ranks in one ncclAllReduce).
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
                                                   int64_t n_chunks, double *partials) {
  __shared__ double2 ws[CHUNK];
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
};
__global__ void __launch_bounds__(1024) k_multi_finalize(const MultiFin M) {
  __shared__ double sm[32];
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
                                                    int64_t No, double *partials) {
  __shared__ double hs[512];
  __shared__ double red[32];
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
// out = a * x (+ y)
__global__ void k_scale_add(double a, const double2 *x, const double2 *y, double2 *out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 r = make_double2(a * x[i].x, a * x[i].y);
  if (y) {
    r.x += y[i].x;
    r.y += y[i].y;
  }
  out[i] = r;
}
// out = sum_k y_k V_k
__global__ void __launch_bounds__(TPB) k_combine(const double2 *V, int64_t ldv, int nv, const double *y, double2 *out,
                                                 int64_t No) {
  __shared__ double ys[512];
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

struct Reducer {
  Ctx *ctx;
  DBuf<double> partials, gsums, grecv, out;
  std::vector<double> host;
  void init(Ctx *c, int nvmax) {
    ctx = c;
    const int64_t nch = c->n_chunks > 0 ? c->n_chunks : 1;
    partials.alloc((size_t)nvmax * nch);
    gsums.alloc((size_t)nvmax * MAX_GROUPS);
    grecv.alloc((size_t)nvmax * MAX_GROUPS);
    out.alloc(nvmax);
    CUDA_CHECK(cudaMemsetAsync(gsums.p, 0, sizeof(double) * nvmax * MAX_GROUPS, c->stream));
    host.resize(nvmax);
  }
  // finishes the nv reductions whose chunk partials are in `partials`; results stay in out.p
  void finalize(int nv) {
    MultiFin M;
    M.partials = partials.p;
    M.n_chunks = ctx->n_chunks;
    M.cpg = ctx->chunks_per_group;
    M.group_begin = ctx->group_begin;
    M.n_groups_local = ctx->n_groups_local;
    M.n_groups_global = (int)ctx->n_groups_global;
    M.gsums = gsums.p;
    M.grecv = grecv.p;
    M.out = out.p;
    if (ctx->nranks == 1) {
      M.stage = 0;
      GLAUNCH(ctx, k_multi_finalize, nv, 1024, M);
    } else {
      M.stage = 1;
      GLAUNCH(ctx, k_multi_finalize, nv, 1024, M);
      comm_allreduce_sum(ctx, gsums.p, grecv.p, (int64_t)nv * MAX_GROUPS);
      M.stage = 2;
      GLAUNCH(ctx, k_multi_finalize, nv, 1024, M);
    }
  }
  void fetch(int nv, double *dst) {
    CUDA_CHECK(cudaMemcpyAsync(dst, out.p, sizeof(double) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
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
  DBuf<double> hdev;
  hdev.alloc(m + 2);
  if (No) CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(double2) * No, ctx->stream));
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m, 0.0), sn(m, 0.0), g(m + 1, 0.0), h1(m + 2), h2(m + 2), y(m);
  std::vector<double> hist;
  // r0 = bscale * b   (x0 = 0)
  if (No) GLAUNCH(ctx, k_scale_add, g1, 256, bscale, b, (const double2 *)nullptr, W, No);
  auto norm2_of = [&](const double2 *v) {
    if (gch) GLAUNCH(ctx, k_multi_dot, gch, 256, v, ld, 1, v, No, ctx->n_chunks, R.partials.p);
    R.finalize(1);
    double s;
    R.fetch(1, &s);
    return sqrt(s);
  };
  const double r0 = norm2_of(W);
  hist.push_back(1.0);
  int iters = 0, converged = 0;
  double relres = 1.0;
  if (r0 == 0.0) {
    converged = 1;
    relres = 0.0;
  }
  double beta = r0;
  while (!converged && iters < maxit && beta > 0.0) {
    // v_0 = r / beta
    if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 1.0 / beta, W, (const double2 *)nullptr, V, No);
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    for (; j < m && iters < maxit; j++) {
      double2 *vj = V + (int64_t)j * ld;
      // w = A (M v_j)
      if (pc) {
        amg_vcycle(ctx, vj, Z, nullptr);
        apply_op_dev(ctx, op, Z, W);
      } else {
        apply_op_dev(ctx, op, vj, W);
      }
      // ICGS, two passes
      const int nv = j + 1;
      if (gch) GLAUNCH(ctx, k_multi_dot, gch, 256, V, ld, nv, W, No, ctx->n_chunks, R.partials.p);
      R.finalize(nv);
      CUDA_CHECK(cudaMemcpyAsync(hdev.p, R.out.p, sizeof(double) * nv, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(h1.data(), R.out.p, sizeof(double) * nv, cudaMemcpyDeviceToHost, ctx->stream));
      if (gch) GLAUNCH(ctx, k_multi_axpy<false>, gch, TPB, V, ld, nv, hdev.p, W, No, (double *)nullptr);
      if (gch) GLAUNCH(ctx, k_multi_dot, gch, 256, V, ld, nv, W, No, ctx->n_chunks, R.partials.p);
      R.finalize(nv);
      CUDA_CHECK(cudaMemcpyAsync(hdev.p, R.out.p, sizeof(double) * nv, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(h2.data(), R.out.p, sizeof(double) * nv, cudaMemcpyDeviceToHost, ctx->stream));
      if (gch) GLAUNCH(ctx, k_multi_axpy<true>, gch, TPB, V, ld, nv, hdev.p, W, No, R.partials.p);
      R.finalize(1);
      double nn;
      R.fetch(1, &nn);  // synchronises: h1, h2 are on the host too
      const double hn = sqrt(nn);
      // Hessenberg column j, previous rotations, new rotation
      for (int i = 0; i <= j; i++) H[(size_t)i * m + j] = h1[i] + h2[i];
      H[(size_t)(j + 1) * m + j] = hn;
      for (int i = 0; i < j; i++) {
        const double a = H[(size_t)i * m + j], c = H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = cs[i] * a + sn[i] * c;
        H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * c;
      }
      {
        const double a = H[(size_t)j * m + j], c = hn, d = hypot(a, c);
        cs[j] = d == 0.0 ? 1.0 : a / d;
        sn[j] = d == 0.0 ? 0.0 : c / d;
        H[(size_t)j * m + j] = d;
        H[(size_t)(j + 1) * m + j] = 0.0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
      }
      iters++;
      relres = fabs(g[j + 1]) / r0;
      hist.push_back(relres);
      if (relres <= tol) {
        converged = 1;
        j++;
        break;
      }
      if (hn == 0.0) {  // happy breakdown: the Krylov space is invariant
        converged = 1;
        j++;
        break;
      }
      if (j + 1 < m + 1 && No) GLAUNCH(ctx, k_scale_add, g1, 256, 1.0 / hn, W, (const double2 *)nullptr, V + (int64_t)(j + 1) * ld, No);
    }
    // y = R^-1 g (j columns), x += M (V y)
    const int k = j;
    for (int i = k - 1; i >= 0; i--) {
      double s = g[i];
      for (int c = i + 1; c < k; c++) s -= H[(size_t)i * m + c] * y[c];
      y[i] = s / H[(size_t)i * m + i];
    }
    if (k > 0) {
      CUDA_CHECK(cudaMemcpyAsync(hdev.p, y.data(), sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream));
      if (No) GLAUNCH(ctx, k_combine, g1, TPB, V, ld, k, hdev.p, T, No);
      if (pc) {
        amg_vcycle(ctx, T, Z, nullptr);
        if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 1.0, Z, X, X, No);
      } else {
        if (No) GLAUNCH(ctx, k_scale_add, g1, 256, 1.0, T, X, X, No);
      }
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // y (host memory) was the source of an async copy
    }
    if (converged || iters >= maxit) break;
    // restart: explicit residual r = b - A x
    if (No) CUDA_CHECK(cudaMemcpyAsync(T, X, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
    apply_op_dev(ctx, op, T, W);
    if (No) GLAUNCH(ctx, k_scale_add, g1, 256, -1.0, W, (const double2 *)nullptr, W, No);
    if (No) GLAUNCH(ctx, k_scale_add, g1, 256, bscale, b, W, W, No);
    beta = norm2_of(W);
  }
  if (p2p_check_error(ctx)) NOSH_THROW(NOSH_ECOMM, "peer-memory exchange timed out (a rank is not responding)");
  if (res) {
    res->iterations = iters;
    res->converged = converged;
    res->relres = relres;
    res->breakdown = 0;
  }
  if (hist_host)
    for (size_t i = 0; i < hist.size() && (int)i <= maxit; i++) hist_host[i] = hist[i];
  // the Arnoldi basis is (restart+1) full vectors (38 GB for GMRES(300) on 8.0M vertices): keep a small one
  // for the next solve, give a large one back
  if (ctx->gmres_basis.n * sizeof(double2) > ((size_t)4 << 30)) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->gmres_basis.release();
  }
}

}  // namespace nosh
