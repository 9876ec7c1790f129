// apply.cuh -- the fused operator-apply kernels (rows a11, a13, a14, a19 of SURVEY.md 8).
//
//   y_i = sum_j K_ij x_j  (+ epilogue)      one pass over the block matrix
//
// complex block matrix (16 B value + 4 B column per block), interleaved (re,im) vectors
// read/written as 128-bit double2.  Epilogues fold what the reference does in a second
// sweep over y into the same kernel:
//   EPI_DIAG : + [[d0r, d1],[d1, d0i]] x_i      jacobian_operator::apply, src/jacobian_operator.cpp:95-100
//   EPI_F    : + c t (V + g |x_i|^2) x_i         nls::compute_f_,        src/model_evaluator_nls.cpp:618-624
//   EPI_DG   : + c t |x_i|^2 x_i                 computeDFDP_ "g",       :665-674
//   EPI_DV   : + c t dV_i x_i                    computeDFDP_ other,     :676-691
// and the Krylov fusions ride on the rows while they are in registers:
//   FUSE_MINRES : input scaled by 1/beta (v = r/beta never materialised), y -= (beta/oldBeta) r1,
//                 chunk partials of <v, y>
//   FUSE_CG     : chunk partials of <p, A p>
// and the smoother steps of the AMG V-cycle (amg.cu) on the finest level:
//   FUSE_RESID  : y = b - A x
//   FUSE_CHEB   : t = D^-1 (b - A x);  d = c1 d + c2 t;  y = x + d     (one Chebyshev/Jacobi step)
//
// Two storage layouts behind the same slots (chosen by measurement):
//   SELL-32 : one thread per row, 32-row slices stored column-major -> every matrix load is a
//             fully coalesced 512 B (values) / 128 B (columns) warp access, no shuffles
//   CSR     : LPR lanes per row + xor-shuffle reduction
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace nosh {

enum { EPI_NONE = 0, EPI_DIAG = 1, EPI_F = 2, EPI_DG = 3, EPI_DV = 4 };
enum { FUSE_NONE = 0, FUSE_AXPBY = 1, FUSE_MINRES = 2, FUSE_CG = 3, FUSE_RESID = 4, FUSE_CHEB = 5 };

struct ApplyArgs {
  int64_t No;
  int64_t nslices;
  const int32_t *rowptr;     // CSR: No+1
  const int32_t *slice_off;  // SELL: nslices+1
  const int32_t *sell_row;   // SELL-32-sigma: row stored at a SELL position (NULL: identity); mesh.cu
  const int32_t *col;
  const double2 *val;
  const float2 *val32;  // optional fp32 copy of val (FUSE_RESID / FUSE_CHEB on SELL-32: the mixed-precision V-cycle)
  const double2 *x;  // Nl entries (owned + ghosts); owned entries only when xg is set
  const double2 *xg; // optional: ghost entries (column c >= No reads xg[c - No]) -- the landing slot of the
                     // peer-memory halo exchange, so the caller's vector needs no ghost room (FUSE_NONE/AXPBY)
  double2 *y;        // No entries
  // epilogues
  const double2 *d0;
  const double *d1;
  const double *cv, *thick, *V;
  double g;
  // FUSE_AXPBY: y = a*(A x) + b*y
  double a, b;
  // Krylov fusions
  const KrylovState *st;
  const double2 *r1;
  double *partials;
  int host_iter;
  // with xg: CTAs with blockIdx.x >= halo_first_block wait for the neighbours' halo flags of exchange
  // halo_epoch before they gather (the chunk list puts the chunks that reference ghosts last)
  const HaloView *halo;
  unsigned long long halo_epoch;
  int halo_first_block;
  // ... and the first push_blocks CTAs first store my boundary entries of x into the neighbours' landing buffers
  // (SELL-32 kernels; 0 = the push was launched separately)
  const int32_t *send_idx;
  int64_t n_send;
  int push_blocks;
  // optional list of chunks to process (interior / boundary split); NULL = all chunks
  const int32_t *chunk_list;
  int n_list;
  // FUSE_MINRES / FUSE_CG on one GPU: the last CTA finishes the reduction (fin.counter != NULL)
  FinArgs fin;
  // FUSE_RESID / FUSE_CHEB
  const double2 *bvec;
  double2 *dvec;        // Chebyshev direction (read if c1 != 0, written if non-NULL)
  const double2 *dinv;  // 1 / point diagonal
  double c1, c2;
  // optional: every variant returns immediately once gate->done is set (V-cycles launched after
  // the Krylov solver has converged)
  const KrylovState *gate;
};

void launch_apply(Ctx *ctx, int epi, int fuse, const ApplyArgs &A);

}  // namespace nosh
