// fvm.h -- generic real-valued FVM matrix / operator cores on the device (fvm.cu)
#pragma once
#include "common.cuh"
namespace nosh {
void fvm_boundary_vertices(Ctx *ctx, int32_t *flags_dev);
// all array arguments are HOST pointers (or NULL); see include/nosh_b200.h:nosh_fvm_matrix_fill
void fvm_matrix_fill(Ctx *ctx, const double *edge_coeff, const double *edge_lhs, const double *edge_rhs,
                     const double *vertex_lhs, const double *vertex_rhs, const int32_t *dmask, const double *dval);
// device pointers
void fvm_apply_dev(Ctx *ctx, bool with_matrix, int vertex_kind, double alpha, const double *u0, const int32_t *mask,
                   int dirichlet_kind, const double *dval, const double *x, double *y);
void fvm_cg_dev(Ctx *ctx, const double *b, double *x, double tol, int maxit, nosh_krylov_result *res);
}  // namespace nosh
