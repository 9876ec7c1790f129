// krylov.cu -- Krylov solvers (a18) and the Newton driver with device-resident control flow.
//
// MINRES follows Belos::MinresIter's organisation of Paige-Saunders MINRES (restated in
// oracle/nosh_oracle.cpp:minres; Belos itself is not in the reference tree).  One
// iteration is three vector kernels and two one-CTA "finalize" kernels:
//
//   A  y   = J (r_k/beta_k) - (beta_k/beta_{k-1}) r_{k-1},  partials of <v_k, y>   [apply.cu, fused]
//   fA alpha_k
//   B  r_{k+1} = y - (alpha_k/beta_k) r_k,                  partials of <r_{k+1}, r_{k+1}>
//   fB beta_{k+1}, Givens rotation, phi, convergence test, iteration counter
//   C  w_k = (v_k - eps w_{k-2} - delta w_{k-1})/gamma,  x += phi w_k
//
// All scalars live in a KrylovState in HBM; the host never reads one inside the loop, it
// only polls `done` every few iterations.  Once `done` is set every later launch returns
// immediately, so x is exactly the iterate of the passing iteration.
// Reductions are the fixed three-level tree of common.cuh => bit-identical results for
// any number of GPUs.
#include "krylov.h"

#include <cooperative_groups.h>

#include <cstdlib>

#include "amg.h"
#include "apply.cuh"
#include "comm.h"
#include "reduce.cuh"
#include "keo.h"
#include "kmath.cuh"

namespace nosh {

namespace {

#define KLAUNCH(ctx, kernel, grid, block, ...)                    \
  do {                                                            \
    kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);   \
    (ctx)->launches++;                                            \
    CUDA_CHECK(cudaGetLastError());                               \
  } while (0)

// ---- chunked vector kernels: one CTA (256 threads) per CHUNK = 512 vertices ------------------

__global__ void __launch_bounds__(TPB) k_dot(const double2 *x, const double2 *y, int64_t No, double *partials,
        const FinArgs fin) {
  __shared__ double red[32];
  const int64_t i0 = (int64_t)blockIdx.x * CHUNK + threadIdx.x, i1 = i0 + TPB;
  double c = 0.0;
  if (i0 < No) c = cdot(x[i0], y[i0]);
  if (i1 < No) c += cdot(x[i1], y[i1]);
  const double s = block_sum<TPB / 32>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  last_cta_finalize(fin, red);
}

// control-volume weighted sums: mode 0: sum c_k; 1: sum c_k Re(conj(a_k) b_k)  (nls::inner_product,
// src/model_evaluator_nls.cpp:699-739); 2: -sum c_k |a_k|^4  (nls::gibbs_energy, :742-770)
__global__ void __launch_bounds__(TPB) k_weighted(int mode, const double *cv, const double2 *a, const double2 *b,
                                                  int64_t No, double *partials,
        const FinArgs fin) {
  __shared__ double red[32];
  double c = 0.0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + h * TPB;
    if (i < No) {
      if (mode == 0) {
        c += cv[i];
      } else if (mode == 1) {
        c += cv[i] * cdot(a[i], b[i]);
      } else {
        const double al = cdot(a[i], a[i]);
        c -= cv[i] * al * al;
      }
    }
  }
  const double s = block_sum<TPB / 32>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  last_cta_finalize(fin, red);
}

__global__ void __launch_bounds__(TPB) k_minres_init(const double2 *b, double bscale, int64_t No, double2 *r0,
                                                     double2 *w0, double2 *w1, double2 *w2, double2 *x,
                                                     double *partials,
        const FinArgs fin) {
  __shared__ double red[32];
  double c = 0.0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + h * TPB;
    if (i < No) {
      double2 r = b[i];
      r.x *= bscale;
      r.y *= bscale;
      r0[i] = r;
      const double2 z = make_double2(0.0, 0.0);
      w0[i] = z;
      w1[i] = z;
      w2[i] = z;
      x[i] = z;
      c += cdot(r, r);
    }
  }
  const double s = block_sum<TPB / 32>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  last_cta_finalize(fin, red);
}

__global__ void __launch_bounds__(TPB) k_minres_B(const KrylovState *st, int host_iter, const double2 *p,
                                                  const double2 *r2, double2 *rnew, int64_t No,
                                                  double *partials,
        const FinArgs fin) {
  if (st->done || st->iter != host_iter - 1) return;
  __shared__ double red[32];
  const double f = st->f_r2;
  const int64_t i0 = (int64_t)blockIdx.x * CHUNK + threadIdx.x, i1 = i0 + TPB;
  double2 p0, p1, q0, q1;
  if (i0 < No) { p0 = ld_stream2(p + i0); q0 = ld_stream2(r2 + i0); }
  if (i1 < No) { p1 = ld_stream2(p + i1); q1 = ld_stream2(r2 + i1); }
  double c = 0.0;
  if (i0 < No) {
    const double2 r = sub_scaled(p0, f, q0);
    rnew[i0] = r;
    c = cdot(r, r);
  }
  if (i1 < No) {
    const double2 r = sub_scaled(p1, f, q1);
    rnew[i1] = r;
    c += cdot(r, r);
  }
  const double s = block_sum<TPB / 32>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  last_cta_finalize(fin, red);
}

__global__ void __launch_bounds__(TPB) k_minres_C(const KrylovState *st, int host_iter, const double2 *rcur,
                                                  const double2 *w1, const double2 *w2, double2 *wnew,
                                                  double2 *x, int64_t No) {
  if (st->iter != host_iter) return;  // this iteration did not execute (already converged)
  const double ib = st->inv_beta_prev, oe = st->oldeps, de = st->delta, ig = st->inv_gamma, ph = st->phi;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + h * TPB;
    if (i < No) {
      const double2 r = ld_stream2(rcur + i), a = ld_stream2(w1 + i), b = ld_stream2(w2 + i);
      const double2 w = minres_w(r, a, b, ib, oe, de, ig);
      wnew[i] = w;
      x[i] = axpy2(ph, w, x[i]);
    }
  }
}

// ---- CG vector kernels --------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_cg_init(const double2 *b, double bscale, int64_t No, double2 *r,
                                                 double2 *p, double2 *x, double *partials,
        const FinArgs fin) {
  __shared__ double red[32];
  double c = 0.0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + h * TPB;
    if (i < No) {
      double2 v = b[i];
      v.x *= bscale;
      v.y *= bscale;
      r[i] = v;
      p[i] = v;
      x[i] = make_double2(0.0, 0.0);
      c += cdot(v, v);
    }
  }
  const double s = block_sum<TPB / 32>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  last_cta_finalize(fin, red);
}
__global__ void __launch_bounds__(TPB) k_cg_update(const KrylovState *st, int host_iter, const double2 *p,
                                                   const double2 *ap, double2 *r, double2 *x, int64_t No,
                                                   double *partials,
        const FinArgs fin) {
  if (st->done || st->iter != host_iter - 1) return;
  __shared__ double red[32];
  const double al = st->cg_alpha;
  double c = 0.0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + h * TPB;
    if (i < No) {
      const double2 pp = p[i], aa = ld_stream2(ap + i);
      double2 xx = x[i], rr = r[i];
      xx.x += al * pp.x;
      xx.y += al * pp.y;
      rr.x -= al * aa.x;
      rr.y -= al * aa.y;
      x[i] = xx;
      r[i] = rr;
      c += cdot(rr, rr);
    }
  }
  const double s = block_sum<TPB / 32>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
  last_cta_finalize(fin, red);
}
__global__ void __launch_bounds__(TPB) k_cg_direction(const KrylovState *st, int host_iter, const double2 *r,
                                                      double2 *p, int64_t No) {
  if (st->iter != host_iter) return;
  const double be = st->cg_beta;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x + h * TPB;
    if (i < No) {
      const double2 rr = r[i];
      double2 pp = p[i];
      pp.x = rr.x + be * pp.x;
      pp.y = rr.y + be * pp.y;
      p[i] = pp;
    }
  }
}

// ---- misc vector kernels ---------------------------------------------------------------------------
__global__ void k_axpy(double a, const double2 *x, double2 *y, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const double2 xx = x[i];
    double2 yy = y[i];
    yy.x += a * xx.x;
    yy.y += a * xx.y;
    y[i] = yy;
  }
}
// z = a x + b y   (z may alias x or y)
__global__ void k_lincomb(double a, const double2 *x, double b, const double2 *y, double2 *z, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const double2 xx = x[i], yy = y[i];
    z[i] = make_double2(a * xx.x + b * yy.x, a * xx.y + b * yy.y);
  }
}
// jacobian_operator::rebuild_diags_ (src/jacobian_operator.cpp:184-197)
__global__ void k_jac_diags(double g, const double *cv, const double *thick, const double *V,
                            const double2 *psi, int64_t No, double2 *d0, double *d1) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= No) return;
  const double2 x = psi[k];
  const double ct = cv[k] * thick[k];
  const double alpha = ct * (V[k] + g * 2.0 * (x.x * x.x + x.y * x.y));
  const double realX2 = g * ct * (x.x * x.x - x.y * x.y);
  d0[k] = make_double2(alpha + realX2, alpha - realX2);
  d1[k] = g * ct * (2.0 * x.x * x.y);
}
// keo_regularized::rebuild diagonal blocks (src/keo_regularized.cpp:233-259); zero if g <= 0 (:200)
__global__ void k_keoreg_diags(double g, const double *cv, const double *thick, const double2 *psi,
                               int64_t No, double2 *d0, double *d1) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= No) return;
  if (!(g > 0.0)) {
    d0[k] = make_double2(0.0, 0.0);
    d1[k] = 0.0;
    return;
  }
  const double2 x = psi[k];
  const double gct = g * cv[k] * thick[k];
  const double al = gct * 2.0 * (x.x * x.x + x.y * x.y);
  const double be = gct * (2.0 * x.x * x.y);
  const double ga = gct * (x.x * x.x - x.y * x.y);
  d0[k] = make_double2(al + ga, al - ga);
  d1[k] = be;
}

__global__ void __launch_bounds__(1024) k_finalize(const FinArgs F) {
  if (fin_is_iterative(F.what) && (F.st->done || F.st->iter != F.host_iter - 1)) return;
  __shared__ double sm[32];
  finalize_levels<true>(F, sm);
}

// One SELL-32 row of the persistent loops: acc += sum_p val[p] * (scale * x[col[p]]) over p = p0, p0 + 32, ... <
// pend, in ascending p (the order of every other apply kernel => identical bits), PU (column, value) pairs in
// flight per thread.  The remainder is ONE masked batch, not a loop of single dependent col -> x loads: with the
// 15 blocks per row of a Kuhn tet mesh and PU = 4 that loop was three serialised round trips per row; PU = 5
// (15 = 3 x 5) and the masked batch took 4 % off the MINRES iteration and 12 % off the stand-alone apply
// (profiles/r2_summary.md section 11).  x is read with plain loads: it is written inside the same launch.
constexpr int PERSIST_U = 5;
template <int PU>
__device__ __forceinline__ void sell_row_scaled(double2 &acc, const int32_t *col, const double2 *val, const double2 *x,
                                                int p, const int pend, const double scale) {
  for (; p + 32 * (PU - 1) < pend; p += 32 * PU) {
    int c[PU];
    double2 v[PU], xv[PU];
#pragma unroll
    for (int u = 0; u < PU; u++) c[u] = ld_stream_i32(col + p + 32 * u);
#pragma unroll
    for (int u = 0; u < PU; u++) v[u] = ld_stream2(val + p + 32 * u);
#pragma unroll
    for (int u = 0; u < PU; u++) xv[u] = x[c[u]];
#pragma unroll
    for (int u = 0; u < PU; u++) cfma(acc, v[u], scaled(xv[u], scale));
  }
  if (p < pend) {
    int c[PU];
    double2 v[PU], xv[PU];
#pragma unroll
    for (int u = 0; u < PU; u++) c[u] = p + 32 * u < pend ? ld_stream_i32(col + p + 32 * u) : 0;
#pragma unroll
    for (int u = 0; u < PU; u++) v[u] = p + 32 * u < pend ? ld_stream2(val + p + 32 * u) : make_double2(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < PU; u++) xv[u] = x[c[u]];  // masked entries read x[0]: a valid address, never used
#pragma unroll
    for (int u = 0; u < PU; u++)
      if (p + 32 * u < pend) cfma(acc, v[u], scaled(xv[u], scale));
  }
}

// -------------------------------------------------------------------------------------------------------
// Persistent MINRES (one GPU, no preconditioner, SELL-32): the whole iteration loop in ONE cooperative
// launch.  CTA c owns the chunks c, c + grid, ... in every phase, so the element-wise phases (B, C) need no
// grid-wide ordering; grid.sync() separates the SpMV gather from the writes of r and the two reductions:
//     A | sync | group sums | sync | alpha (every CTA, redundantly) . B | sync | group sums | sync | beta . C
// Same chunk partials, same tree, same scalar recurrences as the multi-launch path => bit-identical
// iterates; what is saved are 5 launches + 2 one-CTA finalize kernels per iteration.
// -------------------------------------------------------------------------------------------------------
struct PersistArgs {
  ApplyArgs A;  // matrix and diagonal pointers only
  double2 *R0, *R1, *Pv, *W0, *W1, *W2, *X;
  KrylovState *st;
  double *partials, *gsums, *hist;
  int64_t n_chunks, n_groups;
  int cpg, maxit;
};

__device__ __forceinline__ void persist_level2(const PersistArgs &P) {
  const int l = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (CHUNK / 32) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (CHUNK / 32);
  const int per = (P.cpg + 31) / 32;
  for (int64_t g = gw; g < P.n_groups; g += nw) {
    const int64_t base = g * P.cpg;
    double s = 0.0;
    for (int t = 0; t < per; t++) {
      const int k = l * per + t;
      if (k < P.cpg && base + k < P.n_chunks) s += __ldcg(P.partials + base + k);
    }
    s = warp_sum(s);
    if (l == 0) P.gsums[g] = s;
  }
}
// every CTA: total of the group sums (the level-3 tree of finalize_levels); valid in thread 0
__device__ __forceinline__ double persist_level3(const PersistArgs &P, double *sm) {
  const int nw = CHUNK / 32, w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int seg = w; seg < 32; seg += nw) {
    const int i = 32 * seg + l;
    double v = i < P.n_groups ? __ldcg(P.gsums + i) : 0.0;
    v = warp_sum(v);
    if (l == 0) sm[seg] = v;
  }
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = sm[l];
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

// PU: (column, value) pairs in flight per thread in phase A (same ascending summation order for every PU)
template <int EPI, int PU>
__global__ void __launch_bounds__(CHUNK, 2) k_minres_persistent(const PersistArgs P) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ KrylovState s_st;
  __shared__ double red[32], sm[32];
  __shared__ double pair[CHUNK];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_st = *P.st;
  __syncthreads();
  FinArgs F;
  F.st = &s_st;
  F.hist = blockIdx.x == 0 ? P.hist : nullptr;
  F.out = nullptr;
  F.tol = s_st.tol;
  F.maxit = s_st.maxit;
  for (int h = 1; h <= P.maxit; h++) {
    if (s_st.done) break;  // identical in every CTA
    double2 *rcur = (h & 1) ? P.R0 : P.R1, *rprev = (h & 1) ? P.R1 : P.R0;
    double2 *w1 = h % 3 == 0 ? P.W1 : (h % 3 == 1 ? P.W2 : P.W0);   // W[(h + 1) % 3]
    double2 *w2 = h % 3 == 0 ? P.W2 : (h % 3 == 1 ? P.W0 : P.W1);   // W[(h + 2) % 3]
    double2 *wn = h % 3 == 0 ? P.W0 : (h % 3 == 1 ? P.W1 : P.W2);   // W[h % 3]
    // ---- A: y = J (r_h / beta_h) - (beta_h / beta_{h-1}) r_{h-1}, chunk partials of <v, y> ----
    {
      const double scale = s_st.inv_beta, f = s_st.f_r1;
      for (int64_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int64_t pos = chunk * CHUNK + tid;
        const int64_t row = P.A.sell_row ? (int64_t)__ldg(P.A.sell_row + pos) : pos;
        const int64_t slice = pos >> 5;
        double2 acc = make_double2(0.0, 0.0);
        if (slice < P.A.nslices) {
          int p = __ldg(P.A.slice_off + slice) + lane;
          const int pend = __ldg(P.A.slice_off + slice + 1);
          sell_row_scaled<PU>(acc, P.A.col, P.A.val, rcur, p, pend, scale);
        }
        double contrib = 0.0;
        if (row < P.A.No) {
          const double2 xi = scaled(rcur[row], scale);
          double2 yi = acc;
          if (EPI == EPI_DIAG) yi = diag_epilogue(acc, ld_stream2(P.A.d0 + row), __ldg(P.A.d1 + row), xi);
          if (f != 0.0) yi = sub_scaled(yi, f, rprev[row]);
          contrib = cdot(xi, yi);
          P.Pv[row] = yi;
        }
        const double s = block_sum<CHUNK / 32>(contrib, red);
        if (tid == 0) P.partials[chunk] = s;
      }
    }
    grid.sync();
    persist_level2(P);
    grid.sync();
    {
      const double total = persist_level3(P, sm);
      if (tid == 0) {
        F.what = FIN_MINRES_ALPHA;
        fin_scalars(F, total);
      }
      __syncthreads();
    }
    // ---- B: r_{h+1} = y - (alpha_h / beta_h) r_h (over r_{h-1}), chunk partials of <r_{h+1}, r_{h+1}> ----
    {
      const double f = s_st.f_r2;
      for (int64_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int64_t i = chunk * CHUNK + tid;
        double c = 0.0;
        if (i < P.A.No) {
          const double2 r = sub_scaled(P.Pv[i], f, rcur[i]);
          rprev[i] = r;
          c = cdot(r, r);
        }
        // the multi-launch kernel sums vertex t and t + 256 in one thread, then 8 warps: same tree here
        pair[tid] = c;
        __syncthreads();
        const double c2 = tid < TPB ? pair[tid] + pair[tid + TPB] : 0.0;
        double v = warp_sum(c2);
        if (lane == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
          double s = 0.0;
#pragma unroll
          for (int q8 = 0; q8 < TPB / 32; q8++) s += red[q8];
          P.partials[chunk] = s;
        }
        __syncthreads();
      }
    }
    grid.sync();
    persist_level2(P);
    grid.sync();
    {
      const double total = persist_level3(P, sm);
      if (tid == 0) {
        F.what = FIN_MINRES_BETA;
        fin_scalars(F, total);
      }
      __syncthreads();
    }
    // ---- C: w_h = (v_h - eps w_{h-2} - delta w_{h-1}) / gamma, x += phi w_h (only if the iteration completed) ----
    if (s_st.iter == h) {
      const double ib = s_st.inv_beta_prev, oe = s_st.oldeps, de = s_st.delta, ig = s_st.inv_gamma, ph = s_st.phi;
      for (int64_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int64_t i = chunk * CHUNK + tid;
        if (i < P.A.No) {
          const double2 w = minres_w(rcur[i], w1[i], w2[i], ib, oe, de, ig);
          wn[i] = w;
          P.X[i] = axpy2(ph, w, P.X[i]);
        }
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) *P.st = s_st;
}

// -------------------------------------------------------------------------------------------------------
// The persistent loop for several GPUs: the same kernel with the peer-memory reductions of reduce.cuh
// (stage 3) and the halo push moved inside -- ONE launch per MINRES solve and rank, no NCCL call, the
// compute step and its NVLink exchange in one kernel.  Per reduction:
//     level 2 | sync | CTA 0: store my group sums into every rank's slot, raise the epoch flag, spin on the
//     peers' flags | sync | level 3 from my slot (every CTA, redundantly).
// The halo push of r_{h+1} is a grid-stride loop in the phase after B; every pushing thread fences
// system-wide before the grid sync that precedes the beta exchange, whose flags therefore also order the
// pushed ghosts (same argument as the multi-launch path).  All gathers are plain loads through L1, as in the
// one-GPU kernel: entries written by other SMs and ghost entries written by the peers are both ordered before
// them by the chain  store -> fence -> flag -> CTA 0's spin + fence -> grid sync  (the acquire side invalidates L1).
// Same chunk partials, same tree, same scalar recurrences => bit-identical to the one-GPU kernel.
// -------------------------------------------------------------------------------------------------------
struct PushView {
  double2 *dst[MAX_RANKS];
  int64_t off[MAX_RANKS + 1];
};
struct PersistMgpuArgs {
  PersistArgs S;
  P2PView q;
  int64_t group_begin, n_groups_local;
  const int32_t *send_idx;
  int64_t n_send;
  PushView push0, push1;  // targets when r_{h+1} lives in R0 / R1
  int fence_mode;
  // LEAN schedule: the send list sorted by source chunk, so that the CTA that has just written a chunk of
  // r_{h+1} pushes that chunk's boundary entries itself (no grid sync between phase B and the push)
  const int32_t *csend_ptr;   // n_chunks + 1
  const int32_t *csend_src;   // owned vertex
  const int32_t *csend_rank;  // destination rank
  const int32_t *csend_off;   // position in that rank's ghost block
};

__device__ __forceinline__ void persist_exchange(const PersistMgpuArgs &M, unsigned long long epoch) {
  // CTA 0 only: all-gather of the group sums over NVLink (finalize_levels stage 3)
  const P2PView &q = M.q;
  const int slot = (int)(epoch & 1ull);
  const int nloc = (int)M.n_groups_local;
  for (int i = threadIdx.x; i < nloc * q.P; i += blockDim.x) {
    const int r = i / nloc, g = (int)M.group_begin + i % nloc;
    q.red[r][slot * MAX_GROUPS + g] = __ldcg(M.S.gsums + g);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < q.P) {
    *((volatile unsigned long long *)&q.flags[threadIdx.x][q.me]) = epoch;
    const volatile unsigned long long *mine = (const volatile unsigned long long *)&q.flags[q.me][threadIdx.x];
    const long long t0 = clock64();
    while (*mine < epoch) {
      if (clock64() - t0 > q.timeout) {  // a peer is gone; fail instead of hanging
        *q.err = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

// LEAN (tuning key "mgpu_lean" = 1; NOT the default -- measured slower on 2 B200: 1377 vs 1441 MINRES it/s, twice):
// two grid syncs per iteration instead of six.  After a phase only the chunk partials have to be
// complete before CTA 0 may publish (one grid sync); CTA 0 then sums its rank's groups itself and stores them into
// every rank's slot, and EVERY CTA waits for the ranks' flags on its own (local memory) instead of meeting at a
// second grid-wide barrier; the halo entries of a chunk are pushed by the CTA that has just computed them.  Same
// partials, same group sums (same lane runs, same xor tree), same level 3 => same bits as the other schedules
// (mgpu_worker.py runs it).  Why it loses is not resolved: hiding the send-list latency and batching CTA 0's loads
// changed nothing; the suspects left are 296 CTAs polling the flags next to the one CTA that works, and the
// system-scope acquire fences in every CTA (the six-sync schedule fences system-wide in CTA 0 only).
template <int EPI, bool LEAN>
__global__ void __launch_bounds__(CHUNK, 2) k_minres_persistent_mgpu(const PersistMgpuArgs M) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const PersistArgs &P = M.S;
  __shared__ KrylovState s_st;
  __shared__ double red[32], sm[32];
  __shared__ double pair[CHUNK];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_st = *P.st;
  __syncthreads();
  FinArgs F;
  F.st = &s_st;
  F.hist = blockIdx.x == 0 ? P.hist : nullptr;
  F.out = nullptr;
  F.tol = s_st.tol;
  F.maxit = s_st.maxit;
  unsigned long long epoch = *M.q.epoch_ctr;  // reductions executed before this launch
  // level 2 over my groups, level 3 over the gathered sums of all ranks
  auto level2 = [&]() {
    const int64_t gw = (int64_t)blockIdx.x * (CHUNK / 32) + (tid >> 5), nw = (int64_t)gridDim.x * (CHUNK / 32);
    const int per = (P.cpg + 31) / 32;
    for (int64_t g = gw; g < M.n_groups_local; g += nw) {
      const int64_t base = g * P.cpg;
      double s = 0.0;
      for (int t = 0; t < per; t++) {
        const int k = lane * per + t;
        if (k < P.cpg && base + k < P.n_chunks) s += __ldcg(P.partials + base + k);
      }
      s = warp_sum(s);
      if (lane == 0) P.gsums[M.group_begin + g] = s;
    }
  };
  auto level3 = [&](unsigned long long ep) -> double {
    const double *gs = M.q.red[M.q.me] + (int)(ep & 1ull) * MAX_GROUPS;
    const int nw = CHUNK / 32, w = tid >> 5;
    for (int seg = w; seg < 32; seg += nw) {
      const int i = 32 * seg + lane;
      double v = i < P.n_groups ? __ldcg(gs + i) : 0.0;
      v = warp_sum(v);
      if (lane == 0) sm[seg] = v;
    }
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
      t = sm[lane];
      t = warp_sum(t);
    }
    __syncthreads();
    return t;
  };
  // LEAN, CTA 0 only: level 2 of my groups straight into every rank's slot (own included), then the flags
  auto publish = [&](unsigned long long ep) {
    const P2PView &q = M.q;
    const int slot = (int)(ep & 1ull);
    const int per = (P.cpg + 31) / 32;
    // four groups per warp and pass: their partials are loaded before any of them is reduced (the loads of one
    // group alone are a chain of L2 latencies; CTA 0 is the only one working here)
    for (int64_t g0 = (tid >> 5) * 4; g0 < M.n_groups_local; g0 += (CHUNK / 32) * 4) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      if (per <= 8) {
        double v[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int t = 0; t < 8; t++) {
            const int k = lane * per + t;
            const int64_t at = (g0 + u) * P.cpg + k;
            v[u][t] = (t < per && g0 + u < M.n_groups_local && k < P.cpg && at < P.n_chunks) ? __ldcg(P.partials + at) : 0.0;
          }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int t = 0; t < 8; t++)
            if (t < per) s[u] += v[u][t];
      } else {
        for (int u = 0; u < 4; u++)
          for (int t = 0; t < per; t++) {
            const int k = lane * per + t;
            const int64_t at = (g0 + u) * P.cpg + k;
            if (g0 + u < M.n_groups_local && k < P.cpg && at < P.n_chunks) s[u] += __ldcg(P.partials + at);
          }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double t = warp_sum(s[u]);  // xor butterfly: every lane holds the sum
        if (g0 + u < M.n_groups_local && lane < q.P) q.red[lane][slot * MAX_GROUPS + (int)(M.group_begin + g0 + u)] = t;
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid < q.P) *((volatile unsigned long long *)&q.flags[tid][q.me]) = ep;
  };
  // LEAN, every CTA: wait until all ranks' sums (and, behind them, their halo pushes) have landed here.  Only
  // CTA 0 watches the clock: on a time-out it raises the error flag and releases the other CTAs itself.
  auto wait_all = [&](unsigned long long ep) {
    const P2PView &q = M.q;
    if (tid < q.P) {
      volatile unsigned long long *mine = (volatile unsigned long long *)&q.flags[q.me][tid];
      if (blockIdx.x == 0) {
        const long long t0 = clock64();
        while (*mine < ep) {
          if (clock64() - t0 > q.timeout) {
            *q.err = 1;
            __threadfence();
            *mine = ep;
            break;
          }
        }
      } else {
        while (*mine < ep) {
        }
      }
      __threadfence_system();  // acquire: the observers fence, the CTA barrier extends it to the other threads
    }
    __syncthreads();
  };
  for (int h = 1; h <= P.maxit; h++) {
    if (s_st.done) break;
    double2 *rcur = (h & 1) ? P.R0 : P.R1, *rprev = (h & 1) ? P.R1 : P.R0;
    const PushView &push = (h & 1) ? M.push1 : M.push0;  // r_{h+1} is written into rprev
    double2 *w1 = h % 3 == 0 ? P.W1 : (h % 3 == 1 ? P.W2 : P.W0);
    double2 *w2 = h % 3 == 0 ? P.W2 : (h % 3 == 1 ? P.W0 : P.W1);
    double2 *wn = h % 3 == 0 ? P.W0 : (h % 3 == 1 ? P.W1 : P.W2);
    // ---- A ----
    {
      const double scale = s_st.inv_beta, f = s_st.f_r1;
      for (int64_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int64_t pos = chunk * CHUNK + tid;
        const int64_t row = P.A.sell_row ? (int64_t)__ldg(P.A.sell_row + pos) : pos;
        const int64_t slice = pos >> 5;
        double2 acc = make_double2(0.0, 0.0);
        if (slice < P.A.nslices) {
          int p = __ldg(P.A.slice_off + slice) + lane;
          const int pend = __ldg(P.A.slice_off + slice + 1);
          // owned entries: other SMs wrote them; ghosts: the peers did -- both ordered by the exchange + grid sync
          sell_row_scaled<PERSIST_U>(acc, P.A.col, P.A.val, rcur, p, pend, scale);
        }
        double contrib = 0.0;
        if (row < P.A.No) {
          const double2 xi = scaled(rcur[row], scale);
          double2 yi = acc;
          if (EPI == EPI_DIAG) yi = diag_epilogue(acc, ld_stream2(P.A.d0 + row), __ldg(P.A.d1 + row), xi);
          if (f != 0.0) yi = sub_scaled(yi, f, rprev[row]);
          contrib = cdot(xi, yi);
          P.Pv[row] = yi;
        }
        const double s = block_sum<CHUNK / 32>(contrib, red);
        if (tid == 0) P.partials[chunk] = s;
      }
    }
    grid.sync();
    epoch++;
    if (LEAN) {
      if (blockIdx.x == 0) publish(epoch);
      wait_all(epoch);
    } else {
      level2();
      grid.sync();
      if (blockIdx.x == 0) persist_exchange(M, epoch);
      grid.sync();
    }
    if (__ldcg(M.q.err)) break;  // a peer timed out: every CTA sees the flag after the wait and leaves
    {
      const double total = level3(epoch);
      if (tid == 0) {
        F.what = FIN_MINRES_ALPHA;
        fin_scalars(F, total);
      }
      __syncthreads();
    }
    // ---- B ----
    {
      const double f = s_st.f_r2;
      for (int64_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int64_t i = chunk * CHUNK + tid;
        // LEAN: this chunk's range of the send list, loaded now so that its latency hides behind the row work
        int e0 = 0, e1 = 0;
        if (LEAN) {
          e0 = __ldg(M.csend_ptr + chunk);
          e1 = __ldg(M.csend_ptr + chunk + 1);
        }
        double c = 0.0;
        if (i < P.A.No) {
          const double2 r = sub_scaled(P.Pv[i], f, rcur[i]);
          rprev[i] = r;
          c = cdot(r, r);
        }
        pair[tid] = c;
        __syncthreads();
        const double c2 = tid < TPB ? pair[tid] + pair[tid + TPB] : 0.0;
        double v = warp_sum(c2);
        if (lane == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
          double s = 0.0;
#pragma unroll
          for (int q8 = 0; q8 < TPB / 32; q8++) s += red[q8];
          P.partials[chunk] = s;
        }
        __syncthreads();
        if (LEAN) {  // this chunk's boundary entries of r_{h+1} (written above, visible after the barrier) go out now
          for (int e = e0 + tid; e < e1; e += CHUNK)
            push.dst[__ldg(M.csend_rank + e)][__ldg(M.csend_off + e)] = rprev[__ldg(M.csend_src + e)];
        }
      }
    }
    grid.sync();
    if (!LEAN) {
    level2();
    // halo push of r_{h+1}: my boundary entries into the neighbours' ghost segments over NVLink
    {
      bool pushed = false;
      for (int64_t i = (int64_t)blockIdx.x * CHUNK + tid; i < M.n_send; i += (int64_t)gridDim.x * CHUNK) {
        int r = 0;
        while (r + 1 < M.q.P && i >= push.off[r + 1]) r++;
        push.dst[r][i - push.off[r]] = rprev[M.send_idx[i]];
        pushed = true;
      }
      // fence_mode 2 (default): no fence here -- the grid sync orders the pushed stores before CTA 0 at GPU
      // scope, and CTA 0 fences system-wide before it raises the flags (fence cumulativity; the barrier-then-
      // one-thread-fences pattern NCCL's primitives use for peer memory).  1: the pushing threads fence
      // themselves; 0: every thread does.  Measured on 2 B200: 1436 / 1419 / 1304 MINRES it/s.
      if (M.fence_mode == 0 || (M.fence_mode == 1 && pushed)) __threadfence_system();
    }
    grid.sync();
    }
    epoch++;
    if (LEAN) {
      if (blockIdx.x == 0) publish(epoch);  // CTA 0's system fence in there also covers the pushes (cumulativity)
      wait_all(epoch);
    } else {
      if (blockIdx.x == 0) persist_exchange(M, epoch);
      grid.sync();
    }
    if (__ldcg(M.q.err)) break;  // a peer timed out: every CTA sees the flag after the wait and leaves
    {
      const double total = level3(epoch);
      if (tid == 0) {
        F.what = FIN_MINRES_BETA;
        fin_scalars(F, total);
      }
      __syncthreads();
    }
    // ---- C ----
    if (s_st.iter == h) {
      const double ib = s_st.inv_beta_prev, oe = s_st.oldeps, de = s_st.delta, ig = s_st.inv_gamma, ph = s_st.phi;
      for (int64_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const int64_t i = chunk * CHUNK + tid;
        if (i < P.A.No) {
          const double2 w = minres_w(rcur[i], w1[i], w2[i], ib, oe, de, ig);
          wn[i] = w;
          P.X[i] = axpy2(ph, w, P.X[i]);
        }
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    *P.st = s_st;
    *M.q.epoch_ctr = epoch;
  }
}

// Reduction descriptor.  Default: finalize_launch() runs the one-CTA k_finalize kernel after the
// producer.  With NOSH_B200_INKERNEL_FIN=1 (one GPU) the producer's last CTA finishes the reduction
// itself (fin.counter set).  MEASURED SLOWER on B200 (profiles/r1_summary.md, section 4): the
// __threadfence() every CTA needs before taking its ticket invalidates the SM's L1 (CCTL.IVALL), which
// the x gathers of the co-resident SpMV CTAs live on: 151.4 vs 140.4 ms per 200-iteration step.
FinArgs fin_args(Ctx *ctx, int what, int host_iter, double tol, int maxit, double *out, bool in_kernel = true) {
  static const bool enabled = [] {
    const char *e = getenv("NOSH_B200_INKERNEL_FIN");
    return e && atoi(e) != 0;
  }();
  in_kernel = in_kernel && enabled;
  FinArgs F;
  memset(&F, 0, sizeof(F));
  F.partials = ctx->partials.p;
  F.n_chunks = ctx->n_chunks;
  F.cpg = ctx->chunks_per_group;
  F.group_begin = ctx->group_begin;
  F.n_groups_local = ctx->n_groups_local;
  F.n_groups_global = (int)ctx->n_groups_global;
  F.gsend = ctx->group_sums.p;
  F.grecv = ctx->group_sums.p + MAX_GROUPS;
  F.st = ctx->kstate.p;
  F.out = out;
  F.hist = ctx->hist.p;
  F.what = what;
  F.host_iter = host_iter;
  F.tol = tol;
  F.maxit = maxit;
  F.stage = 0;
  F.counter = (in_kernel && ctx->nranks == 1 && ctx->n_chunks > 0) ? ctx->ticket.p : nullptr;
  return F;
}

void finalize_launch(Ctx *ctx, FinArgs &F) {
  if (F.counter) return;  // already done by the producer's last CTA
  if (ctx->nranks == 1) {
    F.stage = 0;
    KLAUNCH(ctx, k_finalize, 1, 1024, F);
  } else if (ctx->p2p.ok) {
    F.stage = 3;
    F.p2p = ctx->p2p.view;
    KLAUNCH(ctx, k_finalize, 1, 1024, F);
  } else {
    F.stage = 1;
    KLAUNCH(ctx, k_finalize, 1, 1024, F);
    comm_allreduce_sum(ctx, F.gsend, ctx->group_sums.p + MAX_GROUPS, ctx->n_groups_global);
    F.stage = 2;
    KLAUNCH(ctx, k_finalize, 1, 1024, F);
  }
}

ApplyArgs base_args(Ctx *ctx, const double2 *val, const double2 *x, double2 *y) {
  ApplyArgs A;
  memset(&A, 0, sizeof(A));
  A.No = ctx->No;
  A.nslices = ctx->nslices;
  A.rowptr = ctx->rowptr.p;
  A.slice_off = ctx->slice_off.p;
  A.sell_row = ctx->sell_permuted ? ctx->sell_row.p : nullptr;
  A.col = ctx->col.p;
  A.val = val;
  A.x = x;
  A.y = y;
  A.a = 1.0;
  A.b = 0.0;
  A.st = ctx->kstate.p;
  A.partials = ctx->partials.p;
  return A;
}

void op_args(Ctx *ctx, int op, ApplyArgs &A, int *epi) {
  if (!ctx->keo_filled) NOSH_THROW(NOSH_ESTATE, "KEO not filled (call nosh_keo_fill / nosh_jac_rebuild first)");
  A.val = ctx->Kval.p;
  switch (op) {
    case NOSH_OP_JACOBIAN:
      if (!ctx->jac_ok) NOSH_THROW(NOSH_ESTATE, "Jacobian not built (call nosh_jac_rebuild first)");
      A.d0 = ctx->jd0.p;
      A.d1 = ctx->jd1.p;
      *epi = EPI_DIAG;
      break;
    case NOSH_OP_KEO:
      *epi = EPI_NONE;
      break;
    case NOSH_OP_KEOREG:
      if (!ctx->keoreg_ok) NOSH_THROW(NOSH_ESTATE, "regularised KEO not built (call nosh_keoreg_rebuild first)");
      A.d0 = ctx->pd0.p;
      A.d1 = ctx->pd1.p;
      *epi = EPI_DIAG;
      break;
    default:
      NOSH_THROW(NOSH_EINVAL, "unknown operator id %d", op);
  }
}

int poll_done(Ctx *ctx, KrylovState *host) {
  CUDA_CHECK(cudaMemcpyAsync(host, ctx->kstate.p, sizeof(KrylovState), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return host->done;
}

}  // namespace

void ensure_work(Ctx *ctx) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  const int64_t n = ctx->Nl > 0 ? ctx->Nl : 1;
  for (auto &w : ctx->work)
    if (w.n < (size_t)n) {
      w.alloc(n);
      CUDA_CHECK(cudaMemsetAsync(w.p, 0, sizeof(double2) * n, ctx->stream));
    }
  ctx->n_chunks = cdiv(ctx->No, CHUNK);
  ctx->partials.ensure(2 * (ctx->n_chunks > 0 ? ctx->n_chunks : 1));
  if (!ctx->group_sums.p) {
    ctx->group_sums.alloc(2 * MAX_GROUPS);
    CUDA_CHECK(cudaMemsetAsync(ctx->group_sums.p, 0, sizeof(double) * 2 * MAX_GROUPS, ctx->stream));
  }
  if (!ctx->kstate.p) ctx->kstate.alloc(1);
  if (!ctx->ticket.p) {
    ctx->ticket.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(ctx->ticket.p, 0, sizeof(unsigned int), ctx->stream));
  }
  if (!ctx->scalar_out.p) {
    ctx->scalar_out.alloc(8);
    CUDA_CHECK(cudaMemsetAsync(ctx->scalar_out.p, 0, sizeof(double) * 8, ctx->stream));
  }
}

// scalar_out[0] = the reduced value, scalar_out[1] != 0: the peer-memory exchange timed out (stale sums)
static double fetch_scalar(Ctx *ctx) {
  double h[2];
  CUDA_CHECK(cudaMemcpyAsync(h, ctx->scalar_out.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (h[1] != 0.0) NOSH_THROW(NOSH_ECOMM, "peer-memory reduction timed out (a rank is not responding)");
  return h[0];
}

double dot_dev(Ctx *ctx, const double2 *x, const double2 *y) {
  ensure_work(ctx);
  FinArgs F = fin_args(ctx, FIN_DOT, 0, 0.0, 0, ctx->scalar_out.p);
  if (ctx->n_chunks) KLAUNCH(ctx, k_dot, (unsigned)ctx->n_chunks, TPB, x, y, ctx->No, ctx->partials.p, F);
  finalize_launch(ctx, F);
  return fetch_scalar(ctx);
}

double weighted_sum_dev(Ctx *ctx, int mode, const double2 *a, const double2 *b) {
  ensure_work(ctx);
  FinArgs F = fin_args(ctx, FIN_DOT, 0, 0.0, 0, ctx->scalar_out.p);
  if (ctx->n_chunks)
    KLAUNCH(ctx, k_weighted, (unsigned)ctx->n_chunks, TPB, mode, ctx->cv.p, a, b, ctx->No, ctx->partials.p, F);
  finalize_launch(ctx, F);
  return fetch_scalar(ctx);
}

void jac_diags_dev(Ctx *ctx, double g, const double2 *psi) {
  ctx->jd0.ensure(ctx->No);
  ctx->jd1.ensure(ctx->No);
  if (ctx->No)
    KLAUNCH(ctx, k_jac_diags, (unsigned)cdiv(ctx->No, 256), 256, g, ctx->cv.p, ctx->thick.p, ctx->Vcur.p, psi,
            ctx->No, ctx->jd0.p, ctx->jd1.p);
  ctx->jac_ok = true;
}

void keoreg_diags_dev(Ctx *ctx, double g, const double2 *psi) {
  ctx->pd0.ensure(ctx->No);
  ctx->pd1.ensure(ctx->No);
  if (ctx->No)
    KLAUNCH(ctx, k_keoreg_diags, (unsigned)cdiv(ctx->No, 256), 256, g, ctx->cv.p, ctx->thick.p, psi, ctx->No,
            ctx->pd0.p, ctx->pd1.p);
  ctx->keoreg_ok = true;
  ctx->keoreg_version++;
  if (ctx->amg_reuse == 0) ctx->amg_valid = false;  // "reuse: type" = "none": new hierarchy per rebuild
}

void axpy_dev(Ctx *ctx, double a, const double2 *x, double2 *y) {
  if (ctx->No) KLAUNCH(ctx, k_axpy, (unsigned)cdiv(ctx->No, 256), 256, a, x, y, ctx->No);
}

// One operator apply with its halo exchange (the Tpetra Import + local SpMV of CrsMatrix::apply,
// src/jacobian_operator.cpp:65).  Several GPUs with peer memory: my boundary entries start travelling
// into the neighbours' landing buffers, pushed by the first CTAs of the ONE apply launch, whose chunk list puts
// the chunks that reference no ghost first; the CTAs of the boundary chunks wait for the neighbours' flags and
// read the ghosts straight from the landing slot (A.xg) -- x itself needs no ghost room and is never copied.  NCCL fallback: exchange into x[No..), one launch.
void apply_halo_dev(Ctx *ctx, int epi, int fuse, ApplyArgs &A, double2 *x) {
  if (ctx->nranks == 1) {
    launch_apply(ctx, epi, fuse, A);
    return;
  }
  if (!ctx->p2p.ok) {
    halo_exchange(ctx, x);
    launch_apply(ctx, epi, fuse, A);
    return;
  }
  // one launch over all chunks, interior first: the CTAs of the boundary chunks (the last ones to be scheduled)
  // wait for the neighbours' flags themselves, so the exchange hides behind the interior rows
  A.chunk_list = ctx->chunks_all.p;
  A.n_list = (int)(ctx->n_chunks_int + ctx->n_chunks_bnd);
  if (ctx->layout == NOSH_LAYOUT_SELL32 && A.n_list > 0) {
    // the push rides on the first CTAs of the same launch: ONE kernel per apply, exchange included
    A.halo_epoch = halo_next_epoch(ctx);
    A.send_idx = ctx->send_idx.p;
    A.n_send = ctx->n_send;
    A.push_blocks = (int)std::min<int64_t>(A.n_list, cdiv(ctx->n_send, CHUNK));
  } else {
    halo_begin(ctx, x);
    A.halo_epoch = ctx->p2p.hepoch;
    A.push_blocks = 0;
  }
  A.xg = halo_slot(ctx);
  A.halo = ctx->p2p.halo_dev.p;
  A.halo_first_block = ctx->Ng > 0 ? (int)ctx->n_chunks_int : A.n_list;
  launch_apply(ctx, epi, fuse, A);
  A.chunk_list = nullptr;
  A.xg = nullptr;
  A.halo = nullptr;
  A.push_blocks = 0;
}

// y = op(x) with the plain/diag epilogues.  nranks > 1 without peer memory: x must have Nl entries.
void apply_op_dev(Ctx *ctx, int op, double2 *x, double2 *y) { apply_op_gated_dev(ctx, op, x, y, nullptr); }
// ... as a no-op once gate->done is set (solvers whose control flow lives on the device; every rank sees the same
// flag, so a skipped launch skips its halo push on all of them)
void apply_op_gated_dev(Ctx *ctx, int op, double2 *x, double2 *y, const KrylovState *gate) {
  ensure_work(ctx);
  ApplyArgs A = base_args(ctx, nullptr, x, y);
  int epi;
  op_args(ctx, op, A, &epi);
  A.gate = gate;
  apply_halo_dev(ctx, epi, FUSE_NONE, A, x);
}

// F(psi) = K psi + c t (V + g |psi|^2) psi   (nls::compute_f_)
void compute_f_dev(Ctx *ctx, double g, double2 *psi, double2 *f) {
  ensure_work(ctx);
  ApplyArgs A = base_args(ctx, ctx->Kval.p, psi, f);
  A.cv = ctx->cv.p;
  A.thick = ctx->thick.p;
  A.V = ctx->Vcur.p;
  A.g = g;
  apply_halo_dev(ctx, EPI_F, FUSE_NONE, A, psi);
}

// z = M r with the selected preconditioner (M = I is handled by the callers: z aliases r)
void precond_apply(Ctx *ctx, int prec, const double2 *r, double2 *z) {
  if (prec == NOSH_PREC_KEOREG_AMG) amg_vcycle(ctx, r, z, ctx->kstate.p);
}

void minres_dev(Ctx *ctx, int op, int prec, const double2 *b, double bscale, double2 *x_out, double tol, int maxit,
                nosh_krylov_result *res, double *hist_host) {
  ensure_work(ctx);
  if (maxit < 0) NOSH_THROW(NOSH_EINVAL, "maxit < 0");
  if (prec != NOSH_PREC_NONE && prec != NOSH_PREC_KEOREG_AMG) NOSH_THROW(NOSH_EINVAL, "unknown preconditioner %d", prec);
  const bool pc = prec != NOSH_PREC_NONE;
  if (pc) amg_ensure(ctx);
  ApplyArgs A = base_args(ctx, nullptr, nullptr, nullptr);
  int epi;
  op_args(ctx, op, A, &epi);
  ctx->hist.ensure((size_t)maxit + 2);
  // Z: what the operator is applied to (z_k = M r_k; the peer-memory halo push exports these two
  // buffers).  Without a preconditioner z_k is r_k itself.
  double2 *Z[2] = {ctx->work[0].p, ctx->work[1].p};
  double2 *R[2] = {pc ? ctx->work[12].p : Z[0], pc ? ctx->work[13].p : Z[1]};
  double2 *Pv = ctx->work[2].p;
  double2 *W[3] = {ctx->work[3].p, ctx->work[4].p, ctx->work[5].p};
  double2 *X = x_out;
  const unsigned grid = (unsigned)ctx->n_chunks;
  const int64_t No = ctx->No;
  FinArgs F = fin_args(ctx, FIN_MINRES_INIT, 0, tol, maxit, nullptr, !pc);
  FinArgs Fnone = fin_args(ctx, FIN_DOT, 0, tol, maxit, nullptr, false);
  if (grid) KLAUNCH(ctx, k_minres_init, grid, TPB, b, bscale, No, R[0], W[0], W[1], W[2], X, ctx->partials.p, pc ? Fnone : F);
  if (pc) {
    // beta_1^2 = <r_1, M r_1>
    amg_vcycle(ctx, R[0], Z[0], nullptr);
    if (grid) KLAUNCH(ctx, k_dot, grid, TPB, R[0], Z[0], No, ctx->partials.p, Fnone);
  }
  if (ctx->nranks > 1 && ctx->p2p.ok) p2p_halo_push(ctx, 0, Z[0]);  // z_1 ghosts; ordered by the reduction below
  finalize_launch(ctx, F);
  KrylovState hs;
  int check = 4;
  if (ctx->nranks == 1 && !pc && ctx->layout == NOSH_LAYOUT_SELL32 && ctx->persistent_minres && grid > 0 &&
      (epi == EPI_DIAG || epi == EPI_NONE) && !F.counter && ctx->persist_grid >= 0) {
    // the whole loop in one cooperative launch (k_minres_persistent)
    PersistArgs PA;
    memset(&PA, 0, sizeof(PA));
    PA.A = A;
    PA.R0 = Z[0];
    PA.R1 = Z[1];
    PA.Pv = Pv;
    PA.W0 = W[0];
    PA.W1 = W[1];
    PA.W2 = W[2];
    PA.X = X;
    PA.st = ctx->kstate.p;
    PA.partials = ctx->partials.p;
    PA.gsums = ctx->group_sums.p;
    PA.hist = ctx->hist.p;
    PA.n_chunks = ctx->n_chunks;
    PA.n_groups = ctx->n_groups_global;
    PA.cpg = ctx->chunks_per_group;
    PA.maxit = maxit;
    const bool u4 = ctx->apply_variant == 6;  // measurement variant (profiles/apply_variants.py): the round-1 batch size
    const void *fn = epi == EPI_DIAG ? (u4 ? (const void *)k_minres_persistent<EPI_DIAG, 4> : (const void *)k_minres_persistent<EPI_DIAG, PERSIST_U>)
                                     : (u4 ? (const void *)k_minres_persistent<EPI_NONE, 4> : (const void *)k_minres_persistent<EPI_NONE, PERSIST_U>);
    if (ctx->persist_grid == 0) {
      int per_sm = 0, sms = 0, coop = 0;
      CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_minres_persistent<EPI_DIAG, PERSIST_U>, CHUNK, 0));
      CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
      ctx->persist_grid = coop ? per_sm * sms : -1;
    }
    if (ctx->persist_grid <= 0) goto multi_launch;  // no cooperative launch on this device / configuration
    {
    const unsigned pgrid = (unsigned)std::min<int64_t>(ctx->persist_grid, ctx->n_chunks);
    void *kargs[] = {&PA};
    CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(pgrid), dim3(CHUNK), kargs, 0, ctx->stream));
    ctx->launches++;
    poll_done(ctx, &hs);
    if (res) {
      res->iterations = hs.iter;
      res->converged = hs.converged;
      res->relres = hs.relres;
      res->breakdown = hs.breakdown;
    }
    if (hist_host) {
      CUDA_CHECK(cudaMemcpyAsync(hist_host, ctx->hist.p, sizeof(double) * (hs.iter + 1), cudaMemcpyDeviceToHost,
                                 ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return;
    }
  }
multi_launch:
  if (ctx->nranks > 1 && ctx->p2p.ok && !pc && ctx->layout == NOSH_LAYOUT_SELL32 && ctx->persistent_mgpu && grid > 0 &&
      (epi == EPI_DIAG || epi == EPI_NONE) && ctx->persist_grid_mgpu >= 0) {
    // one cooperative launch per solve and rank (k_minres_persistent_mgpu).  Every rank must take the same
    // decision: all of them run the same binary on the same kind of device, and n_chunks > 0 everywhere is
    // checked at set-up (a rank that owns nothing makes every rank use the multi-launch loop).
    if (ctx->persist_grid_mgpu == 0) {
      int coop = 0, per_sm = 0, sms = 0;
      CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_minres_persistent_mgpu<EPI_DIAG, true>, CHUNK, 0));
      CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
      ctx->persist_grid_mgpu = (coop && per_sm > 0) ? per_sm * sms : -1;
    }
    bool all_own = true;  // part_begin is the same table on every rank
    for (int r = 0; r < ctx->nranks; r++) all_own = all_own && ctx->part_begin[r + 1] > ctx->part_begin[r];
    if (ctx->persist_grid_mgpu > 0 && all_own) {
      PersistMgpuArgs MA;
      memset(&MA, 0, sizeof(MA));
      MA.S.A = A;
      MA.S.R0 = Z[0];
      MA.S.R1 = Z[1];
      MA.S.Pv = Pv;
      MA.S.W0 = W[0];
      MA.S.W1 = W[1];
      MA.S.W2 = W[2];
      MA.S.X = X;
      MA.S.st = ctx->kstate.p;
      MA.S.partials = ctx->partials.p;
      MA.S.gsums = ctx->group_sums.p;
      MA.S.hist = ctx->hist.p;
      MA.S.n_chunks = ctx->n_chunks;
      MA.S.n_groups = ctx->n_groups_global;
      MA.S.cpg = ctx->chunks_per_group;
      MA.S.maxit = maxit;
      MA.q = ctx->p2p.view;
      MA.group_begin = ctx->group_begin;
      MA.n_groups_local = ctx->n_groups_local;
      MA.send_idx = ctx->send_idx.p;
      MA.n_send = ctx->n_send;
      MA.fence_mode = ctx->mgpu_fence;
      MA.csend_ptr = ctx->csend_ptr.p;
      MA.csend_src = ctx->csend_src.p;
      MA.csend_rank = ctx->csend_rank.p;
      MA.csend_off = ctx->csend_off.p;
      for (int r = 0; r < ctx->nranks; r++) {
        MA.push0.dst[r] = ctx->p2p.R[0][r] + ctx->p2p.ghost_base[r];
        MA.push1.dst[r] = ctx->p2p.R[1][r] + ctx->p2p.ghost_base[r];
        MA.push0.off[r] = MA.push1.off[r] = ctx->send_off[r];
      }
      MA.push0.off[ctx->nranks] = MA.push1.off[ctx->nranks] = ctx->send_off[ctx->nranks];
      const bool lean = ctx->mgpu_lean && ctx->csend_ptr.p;
      const void *fn = epi == EPI_DIAG ? (lean ? (const void *)k_minres_persistent_mgpu<EPI_DIAG, true>
                                               : (const void *)k_minres_persistent_mgpu<EPI_DIAG, false>)
                                       : (lean ? (const void *)k_minres_persistent_mgpu<EPI_NONE, true>
                                               : (const void *)k_minres_persistent_mgpu<EPI_NONE, false>);
      const unsigned pgrid = (unsigned)std::min<int64_t>(ctx->persist_grid_mgpu, ctx->n_chunks);
      void *kargs[] = {&MA};
      CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(pgrid), dim3(CHUNK), kargs, 0, ctx->stream));
      ctx->launches++;
      poll_done(ctx, &hs);
      if (p2p_check_error(ctx)) NOSH_THROW(NOSH_ECOMM, "peer-memory reduction timed out (a rank is not responding)");
      if (res) {
        res->iterations = hs.iter;
        res->converged = hs.converged;
        res->relres = hs.relres;
        res->breakdown = hs.breakdown;
      }
      if (hist_host) {
        CUDA_CHECK(cudaMemcpyAsync(hist_host, ctx->hist.p, sizeof(double) * (hs.iter + 1), cudaMemcpyDeviceToHost,
                                   ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      }
      return;
    }
  }
  // Multi-GPU schedule (same arithmetic, two streams): the halo of r_h travels on stream2
  // while the interior chunks of A(h) run; C(h-1) also runs on stream2, next to A(h) and the
  // alpha all-reduce, and must only be finished before B(h) overwrites the buffer it reads.
  const bool p2p = ctx->nranks > 1 && ctx->p2p.ok;
  const bool overlap = ctx->nranks > 1 && !p2p && !pc;
  cudaStream_t S0 = ctx->stream, S1 = ctx->stream2;
  if (overlap) CUDA_CHECK(cudaEventRecord(ctx->e_b, S0));  // r_1 is ready
  int h_last = 0;
  for (int h = 1; h <= maxit; h++) {
    double2 *rcur = R[(h - 1) & 1], *rprev = R[h & 1];
    double2 *zcur = Z[(h - 1) & 1], *zprev = Z[h & 1];
    A.x = zcur;
    A.y = Pv;
    A.r1 = rprev;
    A.host_iter = h;
    A.partials = ctx->partials.p;
    if (overlap) {
      CUDA_CHECK(cudaStreamWaitEvent(S1, ctx->e_b, 0));
      halo_exchange(ctx, rcur, S1);
      CUDA_CHECK(cudaEventRecord(ctx->e_halo, S1));
      if (h > 1) {
        CUDA_CHECK(cudaStreamWaitEvent(S1, ctx->e_finb, 0));
        if (grid) {
          k_minres_C<<<grid, TPB, 0, S1>>>(ctx->kstate.p, h - 1, rprev, W[h % 3], W[(h + 1) % 3], W[(h + 2) % 3], X,
                                           No);
          ctx->launches++;
          CUDA_CHECK(cudaGetLastError());
        }
        CUDA_CHECK(cudaEventRecord(ctx->e_c, S1));
      }
      // A on the chunks that need no ghost value, then (halo landed) on the rest
      FinArgs Fa = fin_args(ctx, FIN_MINRES_ALPHA, h, tol, maxit, nullptr, false);
      FinArgs Fb = fin_args(ctx, FIN_MINRES_BETA, h, tol, maxit, nullptr, false);
      A.fin = Fa;
      A.chunk_list = ctx->chunks_int.p;
      A.n_list = (int)ctx->n_chunks_int;
      launch_apply(ctx, epi, FUSE_MINRES, A);
      CUDA_CHECK(cudaStreamWaitEvent(S0, ctx->e_halo, 0));
      A.chunk_list = ctx->chunks_bnd.p;
      A.n_list = (int)ctx->n_chunks_bnd;
      launch_apply(ctx, epi, FUSE_MINRES, A);
      finalize_launch(ctx, Fa);
      if (h > 1) CUDA_CHECK(cudaStreamWaitEvent(S0, ctx->e_c, 0));
      if (grid) KLAUNCH(ctx, k_minres_B, grid, TPB, ctx->kstate.p, h, Pv, rcur, rprev, No, ctx->partials.p, Fb);
      CUDA_CHECK(cudaEventRecord(ctx->e_b, S0));
      finalize_launch(ctx, Fb);
      CUDA_CHECK(cudaEventRecord(ctx->e_finb, S0));
    } else {
      // A: y = J v - (beta/oldBeta) r_prev, partials <v,y>
      // (one GPU: the last CTA of A / B finishes the reduction and the scalar recurrences itself)
      FinArgs Fa = fin_args(ctx, FIN_MINRES_ALPHA, h, tol, maxit, nullptr);
      FinArgs Fb = fin_args(ctx, FIN_MINRES_BETA, h, tol, maxit, nullptr, !pc);
      if (ctx->nranks > 1 && !p2p) halo_exchange(ctx, zcur);  // NCCL fallback of the preconditioned loop
      A.fin = Fa;
      launch_apply(ctx, epi, FUSE_MINRES, A);
      finalize_launch(ctx, Fa);
      // B: r_next = y - (alpha/beta) r_cur  (written over r_prev)
      if (grid)
        KLAUNCH(ctx, k_minres_B, grid, TPB, ctx->kstate.p, h, Pv, rcur, rprev, No, ctx->partials.p, pc ? Fnone : Fb);
      if (pc) {
        // z_next = M r_next, beta_next^2 = <r_next, z_next>
        precond_apply(ctx, prec, rprev, zprev);
        if (grid) KLAUNCH(ctx, k_dot, grid, TPB, rprev, zprev, No, ctx->partials.p, Fnone);
      }
      // multi-GPU, peer-memory path: push the boundary entries of z_next into the neighbours'
      // ghost segments; the beta reduction that follows is also the barrier that orders them
      if (p2p) p2p_halo_push(ctx, h & 1, zprev);
      finalize_launch(ctx, Fb);
      // C: w_h, x
      if (grid)
        KLAUNCH(ctx, k_minres_C, grid, TPB, ctx->kstate.p, h, zcur, W[(h + 1) % 3], W[(h + 2) % 3], W[h % 3], X, No);
    }
    h_last = h;
    if (h % check == 0 || h == maxit) {
      if (poll_done(ctx, &hs)) break;
      if (check < 32) check *= 2;
    }
  }
  if (overlap && h_last >= 1) {
    // the update of the last launched iteration (a no-op if the solver stopped earlier)
    const int h = h_last;
    if (h > 1) CUDA_CHECK(cudaStreamWaitEvent(S0, ctx->e_c, 0));
    if (grid)
      KLAUNCH(ctx, k_minres_C, grid, TPB, ctx->kstate.p, h, R[(h - 1) & 1], W[(h + 1) % 3], W[(h + 2) % 3], W[h % 3],
              X, No);
  }
  poll_done(ctx, &hs);
  if (p2p_check_error(ctx)) NOSH_THROW(NOSH_ECOMM, "peer-memory reduction timed out (a rank is not responding)");
  if (res) {
    res->iterations = hs.iter;
    res->converged = hs.converged;
    res->relres = hs.relres;
    res->breakdown = hs.breakdown;
  }
  if (hist_host) {
    CUDA_CHECK(cudaMemcpyAsync(hist_host, ctx->hist.p, sizeof(double) * (hs.iter + 1), cudaMemcpyDeviceToHost,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}

void cg_dev(Ctx *ctx, int op, int prec, const double2 *b, double bscale, double2 *x_out, double tol, int maxit,
            nosh_krylov_result *res, double *hist_host) {
  ensure_work(ctx);
  if (maxit < 0) NOSH_THROW(NOSH_EINVAL, "maxit < 0");
  if (prec != NOSH_PREC_NONE && prec != NOSH_PREC_KEOREG_AMG) NOSH_THROW(NOSH_EINVAL, "unknown preconditioner %d", prec);
  const bool pc = prec != NOSH_PREC_NONE;
  if (pc) amg_ensure(ctx);
  ApplyArgs A = base_args(ctx, nullptr, nullptr, nullptr);
  int epi;
  op_args(ctx, op, A, &epi);
  ctx->hist.ensure((size_t)maxit + 2);
  double2 *Rv = ctx->work[0].p, *Pd = ctx->work[1].p, *AP = ctx->work[2].p, *Zv = ctx->work[12].p, *X = x_out;
  const unsigned grid = (unsigned)ctx->n_chunks;
  const int64_t No = ctx->No;
  FinArgs F = fin_args(ctx, FIN_CG_INIT, 0, tol, maxit, nullptr);
  FinArgs Fnone = fin_args(ctx, FIN_DOT, 0, tol, maxit, nullptr, false);
  if (grid) KLAUNCH(ctx, k_cg_init, grid, TPB, b, bscale, No, Rv, Pd, X, ctx->partials.p, F);
  finalize_launch(ctx, F);
  if (pc) {
    // p_0 = z_0 = M r_0, rho_0 = <r_0, z_0>
    amg_vcycle(ctx, Rv, Zv, nullptr);
    FinArgs Fr = fin_args(ctx, FIN_PCG_INIT_RHO, 0, tol, maxit, nullptr, false);
    if (grid) KLAUNCH(ctx, k_dot, grid, TPB, Rv, Zv, No, ctx->partials.p, Fnone);
    finalize_launch(ctx, Fr);
    if (No) CUDA_CHECK(cudaMemcpyAsync(Pd, Zv, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  KrylovState hs;
  int check = 4;
  for (int h = 1; h <= maxit; h++) {
    halo_exchange(ctx, Pd);
    A.x = Pd;
    A.y = AP;
    A.host_iter = h;
    FinArgs Fa = fin_args(ctx, FIN_CG_PAP, h, tol, maxit, nullptr);
    FinArgs Fb = fin_args(ctx, pc ? FIN_PCG_RR : FIN_CG_RHO, h, tol, maxit, nullptr);
    A.fin = Fa;
    launch_apply(ctx, epi, FUSE_CG, A);
    finalize_launch(ctx, Fa);
    if (grid) KLAUNCH(ctx, k_cg_update, grid, TPB, ctx->kstate.p, h, Pd, AP, Rv, X, No, ctx->partials.p, Fb);
    finalize_launch(ctx, Fb);
    if (pc) {
      // z = M r, beta = <r, z> / rho  (the finalize is gated on iter == h, i.e. host_iter h + 1)
      precond_apply(ctx, prec, Rv, Zv);
      FinArgs Fr = fin_args(ctx, FIN_PCG_RHO, h + 1, tol, maxit, nullptr, false);
      if (grid) KLAUNCH(ctx, k_dot, grid, TPB, Rv, Zv, No, ctx->partials.p, Fnone);
      finalize_launch(ctx, Fr);
    }
    if (grid) KLAUNCH(ctx, k_cg_direction, grid, TPB, ctx->kstate.p, h, pc ? Zv : Rv, Pd, No);
    if (h % check == 0 || h == maxit) {
      if (poll_done(ctx, &hs)) break;
      if (check < 32) check *= 2;
    }
  }
  poll_done(ctx, &hs);
  if (p2p_check_error(ctx)) NOSH_THROW(NOSH_ECOMM, "peer-memory reduction timed out (a rank is not responding)");
  if (res) {
    res->iterations = hs.iter;
    res->converged = hs.converged;
    res->relres = hs.relres;
    res->breakdown = hs.breakdown;
  }
  if (hist_host) {
    CUDA_CHECK(cudaMemcpyAsync(hist_host, ctx->hist.p, sizeof(double) * (hs.iter + 1), cudaMemcpyDeviceToHost,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}

// the linear solve of the Newton / continuation drivers: J x = bscale * b with the ctx's solver choice
void jacobian_solve(Ctx *ctx, const double2 *b, double bscale, double2 *x, double tol, int maxit,
                    nosh_krylov_result *kr) {
  switch (ctx->lin_solver) {
    case NOSH_SOLVER_MINRES:
      minres_dev(ctx, NOSH_OP_JACOBIAN, ctx->precond, b, bscale, x, tol, maxit, kr, nullptr);
      break;
    case NOSH_SOLVER_CG:
      cg_dev(ctx, NOSH_OP_JACOBIAN, ctx->precond, b, bscale, x, tol, maxit, kr, nullptr);
      break;
    case NOSH_SOLVER_GMRES:
      gmres_dev(ctx, NOSH_OP_JACOBIAN, ctx->precond, b, bscale, x, tol, maxit, ctx->gmres_restart, kr, nullptr);
      break;
    default:
      NOSH_THROW(NOSH_EINVAL, "unknown linear solver %d", ctx->lin_solver);
  }
}

// dF/dp = (dK/dp) psi + { c t |psi|^2 psi  (p == "g")  |  c t dV/dp psi }   (nls::computeDFDP_)
void compute_dfdp_dev(Ctx *ctx, int np, const char *const *names, const double *values, const char *pname,
                      double2 *psi, double2 *out) {
  ensure_work(ctx);
  dkeo_fill(ctx, np, names, values, pname);  // dK/dp for the column's own parameter
  ApplyArgs A = base_args(ctx, ctx->dKval.p, psi, out);
  A.cv = ctx->cv.p;
  A.thick = ctx->thick.p;
  if (strcmp(pname, "g") == 0) {  // src/model_evaluator_nls.cpp:665-674
    apply_halo_dev(ctx, EPI_DG, FUSE_NONE, A, psi);
  } else {  // :676-691
    ctx->dvdp.ensure(ctx->No > 0 ? ctx->No : 1);
    potential_dvdp(ctx, pname, ctx->dvdp.p);
    A.V = ctx->dvdp.p;
    apply_halo_dev(ctx, EPI_DV, FUSE_NONE, A, psi);
  }
}

// Newton, full step.  psi: device vector with Nl entries, updated in place.
void newton_dev(Ctx *ctx, int np, const char *const *names, const double *values, double2 *psi, double nl_tol,
                int nl_maxit, double lin_tol, int lin_maxit, nosh_newton_result *res, int32_t *lin_iters,
                double *fnorms) {
  ensure_work(ctx);
  const double g = param_at(np, names, values, "g");
  keo_fill(ctx, np, names, values, false);
  update_potential(ctx, np, names, values);
  double2 *F = ctx->work[6].p, *D = ctx->work[7].p;
  int k = 0, total = 0, lin_failed = 0;
  compute_f_dev(ctx, g, psi, F);
  double fn = sqrt(dot_dev(ctx, F, F));
  if (fnorms) fnorms[0] = fn;
  while (k < nl_maxit && !(fn < nl_tol)) {
    // evalModel(W_op): jacobian_operator::rebuild (KEO refill is a cache hit) + diagonals
    jac_diags_dev(ctx, g, psi);
    // evalModel(W_prec): keo_regularized::rebuild at the current state (src/model_evaluator_nls.cpp:507-522)
    if (ctx->precond != NOSH_PREC_NONE) keoreg_diags_dev(ctx, g, psi);
    nosh_krylov_result kr;
    memset(&kr, 0, sizeof(kr));
    jacobian_solve(ctx, F, -1.0, D, lin_tol, lin_maxit, &kr);
    if (lin_iters) lin_iters[k] = kr.iterations;
    total += kr.iterations;
    if (kr.breakdown) {  // Lanczos breakdown / indefinite preconditioner: D is not a Newton direction
      lin_failed = 2;
      break;
    }
    if (!kr.converged) lin_failed = 1;  // NOX takes the step of an unconverged linear solve, too; recorded
    axpy_dev(ctx, 1.0, D, psi);
    compute_f_dev(ctx, g, psi, F);
    fn = sqrt(dot_dev(ctx, F, F));
    k++;
    if (fnorms) fnorms[k] = fn;
    if (!(fn == fn)) break;  // NaN residual: nothing further to gain
  }
  if (res) {
    res->steps = k;
    res->converged = fn < nl_tol;
    res->total_linear_iterations = total;
    res->fnorm = fn;
    res->linear_solve_status = lin_failed;
  }
}

// observer::observeSolution / continuation_data_saver::saveSolution: hand the accepted step to the caller
static bool observe_step(Ctx *ctx, int step, double param, double energy, double norm, const double2 *psi) {
  if (!ctx->step_observer) return false;
  std::vector<double> h(2 * (size_t)(ctx->No > 0 ? ctx->No : 1));
  if (ctx->No)
    CUDA_CHECK(cudaMemcpyAsync(h.data(), psi, sizeof(double2) * ctx->No, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return ctx->step_observer(ctx->step_observer_user, step, param, energy, norm, h.data(), 2 * ctx->No) != 0;
}

// Natural-parameter continuation with a tangent predictor (LOCA "Natural" stepper + "Tangent"
// predictor of examples/conf.xml:35-47, constant step; the arc-length variant needs bordered solves
// and is a later row).  Step k: p_k = p_0 + k dp; for k > 0 the predictor solves J t = -dF/dp at the
// previous solution and sets psi += dp t; then the Newton corrector runs at p_k.
void continuation_dev(Ctx *ctx, int np, const char *const *names, const double *values, const char *pname,
                      double dp, int nsteps, double2 *psi, double nl_tol, int nl_maxit, double lin_tol,
                      int lin_maxit, nosh_continuation_step *out) {
  ensure_work(ctx);
  int ip = -1;
  for (int i = 0; i < np; i++)
    if (names[i] && strcmp(names[i], pname) == 0) ip = i;
  if (ip < 0) NOSH_THROW(NOSH_EKEY, "continuation parameter \"%s\" missing", pname);
  std::vector<double> vals(values, values + np);
  const double p0 = vals[ip];
  const double volume = weighted_sum_dev(ctx, 0, nullptr, nullptr);  // control_volumes->norm1()
  double2 *T = ctx->work[9].p, *dF = ctx->work[10].p;
  for (int k = 0; k <= nsteps; k++) {
    nosh_continuation_step st;
    memset(&st, 0, sizeof(st));
    if (k > 0) {
      // tangent predictor at the previous solution / parameter
      const double g = param_at(np, names, vals.data(), "g");
      keo_fill(ctx, np, names, vals.data(), false);
      update_potential(ctx, np, names, vals.data());
      jac_diags_dev(ctx, g, psi);
      if (ctx->precond != NOSH_PREC_NONE) keoreg_diags_dev(ctx, g, psi);
      compute_dfdp_dev(ctx, np, names, vals.data(), pname, psi, dF);
      nosh_krylov_result kr;
      jacobian_solve(ctx, dF, -1.0, T, lin_tol, lin_maxit, &kr);
      st.predictor_linear_iterations = kr.iterations;
      axpy_dev(ctx, dp, T, psi);
    }
    vals[ip] = p0 + k * dp;
    nosh_newton_result nr;
    newton_dev(ctx, np, names, vals.data(), psi, nl_tol, nl_maxit, lin_tol, lin_maxit, &nr, nullptr, nullptr);
    st.step = k;
    st.param = vals[ip];
    st.newton_steps = nr.steps;
    st.converged = nr.converged;
    st.linear_iterations = nr.total_linear_iterations;
    st.fnorm = nr.fnorm;
    st.gibbs_energy = weighted_sum_dev(ctx, 2, psi, psi) / volume;
    st.norm = sqrt(weighted_sum_dev(ctx, 1, psi, psi) / volume);
    if (out) out[k] = st;
    const bool stop = nr.converged && observe_step(ctx, k, st.param, st.gibbs_energy, st.norm, psi);
    if (!nr.converged || stop) {
      for (int j = k + 1; j <= nsteps && out; j++) {
        memset(&out[j], 0, sizeof(st));
        out[j].step = -1;
      }
      break;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Pseudo-arclength continuation: the LOCA configuration of the reference's nosh-cont
// (examples/conf.xml:35-75: "Continuation Method" = "Arc Length", "Predictor" = "Tangent", "Step Size"
// "Method" = "Adaptive", aggressiveness 2).  LOCA is not in the reference tree: what is restated
// (oracle/continuation.py) is the textbook bordering form it implements:
//   unknowns (x, p), constraint  g = <xdot, x - x0>/len + pdot (p - p0) - ds = 0   (LOCA's scaled dot
//   product: Euclidean / vector length, parameter scaling theta = 1)
//   Newton on the bordered system by two solves with the same Jacobian:  J a = -F,  J b = -dF/dp,
//   dp = -(g + <xdot,a>/len) / (pdot + <xdot,b>/len),  x += a + dp b,  p += dp
//   tangent after every accepted step:  J t = -dF/dp,  (xdot, pdot) = +-(t, 1)/sqrt(<t,t>/len + 1), sign
//   such that the direction is kept
//   step size: ds *= 1 + aggr ((nl_maxit - its)/(nl_maxit - 1))^2 after success, ds /= 2 after failure.
// Optional, opt->flags (LOCA's defaults "Enable Arc Length Scaling" / "Hit Continuation Bound", which nosh-cont
// inherits; derivations in oracle/continuation.py):
//   NOSH_ARC_SCALING    the scaled dot product is <x,y>/len + s^2 p q; when a new tangent has s |pdot| > c_max,
//                       s is reset such that s |pdot| = c_goal, the tangent renormalised and ds, ds_min, ds_max
//                       multiplied by |pdot_old / pdot_new|; the step sizes of the options are parameter
//                       increments, divided by |pdot| of the first tangent
//   NOSH_ARC_HIT_BOUND  a step whose predictor would leave [min_value, max_value] is shortened to land on the
//                       bound and is the last arc-length step; one natural-continuation step (constant predictor)
//                       to the bound itself ends the run
// ---------------------------------------------------------------------------------------------------------
void arclength_dev(Ctx *ctx, int np, const char *const *names, const double *values, const char *pname,
                   const nosh_arclength_options *opt, double2 *psi, nosh_arclength_step *out, int *nsteps_out) {
  ensure_work(ctx);
  int ip = -1;
  for (int i = 0; i < np; i++)
    if (names[i] && strcmp(names[i], pname) == 0) ip = i;
  if (ip < 0) NOSH_THROW(NOSH_EKEY, "continuation parameter \"%s\" missing", pname);
  if (!(opt->initial_step_size != 0.0) || !(opt->min_step_size > 0.0) || !(opt->max_step_size >= opt->min_step_size) ||
      opt->max_steps < 0 || opt->nl_maxit < 2)
    NOSH_THROW(NOSH_EINVAL, "bad arc-length options");
  const bool scaling = (opt->flags & NOSH_ARC_SCALING) != 0, hit_bound = (opt->flags & NOSH_ARC_HIT_BOUND) != 0;
  const double c_goal = opt->goal_contribution > 0.0 ? opt->goal_contribution : 0.5;
  const double c_max = opt->max_contribution > 0.0 ? opt->max_contribution : 0.8;
  const double sc_min = opt->min_scale > 0.0 ? opt->min_scale : 1e-3;
  if (scaling && !(c_goal < 1.0 && c_max < 1.0)) NOSH_THROW(NOSH_EINVAL, "arc-length contributions must be < 1");
  double sc = scaling && opt->initial_scale > 0.0 ? opt->initial_scale : 1.0;  // parameter scale factor s
  std::vector<double> vals(values, values + np);
  const int64_t No = ctx->No, Nl = ctx->Nl > 0 ? ctx->Nl : 1;
  const double len = 2.0 * (double)ctx->n_global;  // Tpetra vector length of the complex_map
  const double volume = weighted_sum_dev(ctx, 0, nullptr, nullptr);
  DBuf<double2> X0, XD, Av, Bv, Fp, Dx;
  for (DBuf<double2> *v : {&X0, &XD, &Av, &Bv, &Fp, &Dx}) {
    v->alloc(Nl);
    CUDA_CHECK(cudaMemsetAsync(v->p, 0, sizeof(double2) * Nl, ctx->stream));
  }
  double2 *F = ctx->work[6].p;
  const unsigned g1 = (unsigned)cdiv(No > 0 ? No : 1, 256);
  auto lincomb = [&](double a, const double2 *x, double b, const double2 *y, double2 *z) {
    if (No) KLAUNCH(ctx, k_lincomb, g1, 256, a, x, b, y, z, No);
  };
  auto prepare = [&](const double2 *x) {  // K(p), V(p), Jacobian (+ preconditioner) diagonals at x
    const double g = param_at(np, names, vals.data(), "g");
    keo_fill(ctx, np, names, vals.data(), false);
    update_potential(ctx, np, names, vals.data());
    jac_diags_dev(ctx, g, x);
    if (ctx->precond != NOSH_PREC_NONE) keoreg_diags_dev(ctx, g, x);
    return g;
  };
  auto record = [&](int k, int conv, int nsteps, int lin, int pred, double fn, double ds, double pdot) {
    nosh_arclength_step st;
    memset(&st, 0, sizeof(st));
    st.step = k;
    st.converged = conv;
    st.newton_steps = nsteps;
    st.linear_iterations = lin;
    st.predictor_linear_iterations = pred;
    st.param = vals[ip];
    st.fnorm = fn;
    st.step_size = ds;
    st.dparam_ds = pdot;
    st.scale = sc;
    st.gibbs_energy = weighted_sum_dev(ctx, 2, psi, psi) / volume;
    st.norm = sqrt(weighted_sum_dev(ctx, 1, psi, psi) / volume);
    if (out) out[k] = st;
    return conv ? observe_step(ctx, k, st.param, st.gibbs_energy, st.norm, psi) : false;
  };
  // tangent at (psi, p): J t = -dF/dp; returns MINRES iterations, writes (XD, pdot) with the sign that
  // keeps <(XD,pdot)_new, (XD,pdot)_old> > 0 (first call: pdot has the sign of ds); ratio = |pdot before /
  // after a change of the scale factor| (1 without)
  auto tangent = [&](double &pdot, bool first, double sign0, double &ratio) {
    prepare(psi);
    compute_dfdp_dev(ctx, np, names, vals.data(), pname, psi, Fp.p);
    nosh_krylov_result kr;
    jacobian_solve(ctx, Fp.p, -1.0, Bv.p, opt->lin_tol, opt->lin_maxit, &kr);
    const double tt = dot_dev(ctx, Bv.p, Bv.p) / len;
    double pd = 1.0 / sqrt(tt + sc * sc);
    bool flip;
    if (first) {
      flip = sign0 < 0.0;
    } else {  // against the old tangent, in the old scale
      const double along = dot_dev(ctx, Bv.p, XD.p) / len * pd + sc * sc * pd * pdot;
      flip = along < 0.0;
    }
    ratio = 1.0;
    if (scaling) {
      const double c = sc * pd;
      if (c > c_max) {
        double sn = c_goal / pd * sqrt((1.0 - c * c) / (1.0 - c_goal * c_goal));
        if (sn < sc_min) sn = sc_min;
        const double pn = 1.0 / sqrt(tt + sn * sn);
        ratio = pd / pn;
        pd = pn;
        sc = sn;
      }
    }
    if (flip) pd = -pd;
    lincomb(pd, Bv.p, 0.0, Bv.p, XD.p);
    pdot = pd;
    return kr.iterations;
  };

  int done = 0;
  // step 0: solution at the initial parameter value
  nosh_newton_result nr;
  newton_dev(ctx, np, names, vals.data(), psi, opt->nl_tol, opt->nl_maxit, opt->lin_tol, opt->lin_maxit, &nr, nullptr,
             nullptr);
  bool stop = record(0, nr.converged, nr.steps, nr.total_linear_iterations, 0, nr.fnorm, 0.0, 0.0);
  done = 1;
  if (nr.converged && opt->max_steps > 0 && !stop) {
    double ds = opt->initial_step_size, pdot = 0.0, ratio = 1.0;
    double ds_min = opt->min_step_size, ds_max = opt->max_step_size;
    int pred_its = tangent(pdot, true, ds, ratio);
    ds = fabs(ds);
    if (scaling) {  // parameter increments -> arc lengths
      const double u = 1.0 / fabs(pdot);
      ds *= u;
      ds_min *= u;
      ds_max *= u;
    }
    double p0 = vals[ip];
    int reached = 0;
    double bound = 0.0;
    if (No) CUDA_CHECK(cudaMemcpyAsync(X0.p, psi, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int k = 1; k <= opt->max_steps;) {
      // predictor
      bool capped = false;
      if (hit_bound) {
        const double pred = p0 + ds * pdot;
        if (pred > opt->max_value || pred < opt->min_value) {
          bound = pred > opt->max_value ? opt->max_value : opt->min_value;
          ds = (bound - p0) / pdot;
          capped = true;
        }
      }
      lincomb(1.0, X0.p, ds, XD.p, psi);
      vals[ip] = p0 + ds * pdot;
      // corrector: Newton on the bordered system
      int its = 0, lin = 0;
      double nrm = 0.0;
      bool ok = false;
      for (;;) {
        const double g = prepare(psi);
        compute_f_dev(ctx, g, psi, F);
        lincomb(1.0, psi, -1.0, X0.p, Dx.p);
        const double gc = dot_dev(ctx, XD.p, Dx.p) / len + sc * sc * pdot * (vals[ip] - p0) - ds;
        nrm = sqrt(dot_dev(ctx, F, F) + gc * gc);
        if (nrm < opt->nl_tol) {
          ok = true;
          break;
        }
        if (its >= opt->nl_maxit || !(nrm == nrm)) break;
        compute_dfdp_dev(ctx, np, names, vals.data(), pname, psi, Fp.p);
        nosh_krylov_result ka, kb;
        memset(&ka, 0, sizeof(ka));
        memset(&kb, 0, sizeof(kb));
        jacobian_solve(ctx, F, -1.0, Av.p, opt->lin_tol, opt->lin_maxit, &ka);
        jacobian_solve(ctx, Fp.p, -1.0, Bv.p, opt->lin_tol, opt->lin_maxit, &kb);
        lin += ka.iterations + kb.iterations;
        if (ka.breakdown || kb.breakdown) break;  // failed step: halved below
        const double xa = dot_dev(ctx, XD.p, Av.p) / len, xb = dot_dev(ctx, XD.p, Bv.p) / len;
        const double dp = -(gc + xa) / (sc * sc * pdot + xb);
        axpy_dev(ctx, 1.0, Av.p, psi);
        axpy_dev(ctx, dp, Bv.p, psi);
        vals[ip] += dp;
        its++;
      }
      if (!ok) {
        // failed step: halve and retry from the last solution (LOCA: "Failed Step Reduction Factor" 0.5)
        ds *= 0.5;
        if (ds < ds_min) {
          if (No) CUDA_CHECK(cudaMemcpyAsync(psi, X0.p, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
          vals[ip] = p0;
          break;
        }
        continue;
      }
      // accepted
      const double ds_used = ds;
      p0 = vals[ip];
      if (No) CUDA_CHECK(cudaMemcpyAsync(X0.p, psi, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
      const int pred_prev = pred_its;
      pred_its = tangent(pdot, false, 0.0, ratio);
      stop = record(k, 1, its, lin, pred_prev, nrm, ds_used, pdot);
      done = k + 1;
      if (stop) break;
      const double fac = (double)(opt->nl_maxit - its) / (double)(opt->nl_maxit - 1);
      ds *= 1.0 + opt->aggressiveness * fac * fac;
      if (ds > ds_max) ds = ds_max;
      ds *= ratio;
      ds_min *= ratio;
      ds_max *= ratio;
      const bool outside = vals[ip] > opt->max_value || vals[ip] < opt->min_value;
      if (hit_bound && (capped || outside)) {
        if (!capped) bound = vals[ip] > opt->max_value ? opt->max_value : opt->min_value;
        reached = 1;
        break;
      }
      if (outside) break;
      k++;
    }
    if (reached && vals[ip] != bound) {
      // the last step: natural continuation to the bound itself, constant predictor
      const double before = vals[ip];
      vals[ip] = bound;
      newton_dev(ctx, np, names, vals.data(), psi, opt->nl_tol, opt->nl_maxit, opt->lin_tol, opt->lin_maxit, &nr,
                 nullptr, nullptr);
      if (nr.converged) {
        record(done, 1, nr.steps, nr.total_linear_iterations, 0, nr.fnorm, bound - before, pdot);
        done++;
      } else {
        vals[ip] = before;
        if (No) CUDA_CHECK(cudaMemcpyAsync(psi, X0.p, sizeof(double2) * No, cudaMemcpyDeviceToDevice, ctx->stream));
      }
    }
  }
  if (nsteps_out) *nsteps_out = done;
}

}  // namespace nosh
