// krylov.h -- solvers and vector utilities on device pointers (krylov.cu)
#pragma once
#include "apply.cuh"
#include "common.cuh"
namespace nosh {
// one operator apply including its halo exchange (overlapped over peer memory); x: owned entries suffice when
// ctx->p2p.ok, otherwise Nl entries
void apply_halo_dev(Ctx *ctx, int epi, int fuse, ApplyArgs &A, double2 *x);
void ensure_work(Ctx *ctx);
double dot_dev(Ctx *ctx, const double2 *x, const double2 *y);
void jac_diags_dev(Ctx *ctx, double g, const double2 *psi);
void keoreg_diags_dev(Ctx *ctx, double g, const double2 *psi);
void axpy_dev(Ctx *ctx, double a, const double2 *x, double2 *y);
void apply_op_dev(Ctx *ctx, int op, double2 *x, double2 *y);
void apply_op_gated_dev(Ctx *ctx, int op, double2 *x, double2 *y, const KrylovState *gate);
void compute_f_dev(Ctx *ctx, double g, double2 *psi, double2 *f);
void minres_dev(Ctx *ctx, int op, int prec, const double2 *b, double bscale, double2 *x_out, double tol, int maxit,
                nosh_krylov_result *res, double *hist_host);
void cg_dev(Ctx *ctx, int op, int prec, const double2 *b, double bscale, double2 *x_out, double tol, int maxit,
            nosh_krylov_result *res, double *hist_host);
void newton_dev(Ctx *ctx, int np, const char *const *names, const double *values, double2 *psi, double nl_tol,
                int nl_maxit, double lin_tol, int lin_maxit, nosh_newton_result *res, int32_t *lin_iters,
                double *fnorms);
double weighted_sum_dev(Ctx *ctx, int mode, const double2 *a, const double2 *b);
void compute_dfdp_dev(Ctx *ctx, int np, const char *const *names, const double *values, const char *pname,
                      double2 *psi, double2 *out);
void continuation_dev(Ctx *ctx, int np, const char *const *names, const double *values, const char *pname,
                      double dp, int nsteps, double2 *psi, double nl_tol, int nl_maxit, double lin_tol,
                      int lin_maxit, nosh_continuation_step *out);
void gmres_dev(Ctx *ctx, int op, int prec, const double2 *b, double bscale, double2 *x_out, double tol, int maxit,
               int restart, nosh_krylov_result *res, double *hist_host);
void arclength_dev(Ctx *ctx, int np, const char *const *names, const double *values, const char *pname,
                   const nosh_arclength_options *opt, double2 *psi, nosh_arclength_step *out, int *nsteps_out);
}  // namespace nosh
