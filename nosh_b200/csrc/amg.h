// amg.h -- smoothed-aggregation AMG V-cycle for the regularised KEO (amg.cu).
// Replaces the inverse the reference gets from MueLu in keo_regularized::apply /
// rebuildInverse_ (src/keo_regularized.cpp:88-165,290-336).
#pragma once
#include "common.cuh"
namespace nosh {

// real 2x2 block  [[a, b], [c, d]]  acting on (re, im)
struct __align__(32) B22 {
  double a, b, c, d;
};

struct AmgLevel {
  int64_t n = 0;   // nodes (block rows)
  int64_t nb = 0;  // blocks (CSR)
  // block CSR (levels >= 1; level 0 is the ctx's complex SELL matrix + pd0/pd1)
  DBuf<int32_t> rowptr, col;
  DBuf<B22> val;
  // SELL-32 copy used by the apply kernels (levels >= 1)
  int64_t nslices = 0, nstored = 0;
  DBuf<int32_t> slice_off, scol;
  DBuf<B22> sval;
  DBuf<double2> dinv;  // 1 / point diagonal (prolongator smoothing, power iteration)
  DBuf<double2> sinv;  // 1 / absolute row sum: the l1-Jacobi scaling of the smoother, lambda_max(S^-1 A) <= 1
  double lam = 0.0;    // lambda_max(D^-1 A): power-iteration estimate (prolongator damping)
  DBuf<double> offsum;  // level 0: sum_{j != i} (|Re K_ij| + |Im K_ij|) over the owned columns, per state
  double omega = 0.0;  // prolongator damping actually used
  // transfer operators to the next (coarser) level
  int64_t nc = 0;
  DBuf<int32_t> agg;       // n: aggregate of each node
  DBuf<double> p0;         // n: tentative-prolongator weight
  DBuf<int32_t> p_rowptr, p_col;  // P by fine row
  DBuf<B22> p_val;
  DBuf<float4> p_val32;  // level 0 with the mixed-precision cycle: fp32 copy of p_val (a, b, c, d)
  DBuf<int32_t> r_rowptr, r_fine, r_pos;  // P^T by coarse row: fine node, position in p_val
  int64_t p_nnz = 0;
  // V-cycle vectors (n entries; level 0: Nl entries, ghost part stays zero)
  DBuf<double2> b, x, x2, d, r;
};

// one captured V-cycle: the ~25 launches of a cycle for fixed (b, x, gate) pointers as one CUDA graph
struct VcycleGraph {
  const double2 *b;
  double2 *x;
  const KrylovState *gate;
  int degree, coarse_degree, mixed;
  cudaGraphExec_t exec;
  int64_t launches;
};

struct Amg {
  std::vector<AmgLevel *> levels;
  DBuf<double> coarse_inv;  // dense (2 n_last)^2, row-major
  int64_t n_coarse2 = 0;
  double setup_seconds = 0.0;
  std::vector<VcycleGraph> graphs;
  void *store = nullptr;  // one allocation holding every buffer of the levels (build_hierarchy moves them here)
  size_t store_bytes = 0;
  ~Amg() {
    for (auto &g : graphs) cudaGraphExecDestroy(g.exec);
    for (auto *l : levels) delete l;
    coarse_inv.release();
    if (store) cudaFree(store);
  }
};

void amg_free(Ctx *ctx);
// builds the hierarchy for the current regularised KEO (K values + pd0/pd1) unless a valid one is
// kept by the reuse policy
void amg_ensure(Ctx *ctx);
// x = M b  (one V-cycle, zero initial guess).  b: No entries; x: Nl entries (ghost part untouched).
// gate (may be NULL): device Krylov state; all kernels return immediately once gate->done is set.
void amg_vcycle(Ctx *ctx, const double2 *b, double2 *x, const KrylovState *gate);

}  // namespace nosh
