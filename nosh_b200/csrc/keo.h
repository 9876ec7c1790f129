// keo.h -- fields + KEO assembly (keo.cu)
#pragma once
#include "common.cuh"
namespace nosh {
void set_thickness(Ctx *ctx, const double *values, double c);
void set_mvp_explicit(Ctx *ctx, const double *A_host, const double *B);
void set_mvp_constcurl(Ctx *ctx, const double b[3], const double u[3]);
void ensure_alpha(Ctx *ctx);
void keo_fill(Ctx *ctx, int np, const char *const *names, const double *values, bool force);
void dkeo_fill(Ctx *ctx, int np, const char *const *names, const double *values, const char *dname);
void edge_projection(Ctx *ctx, int np, const char *const *names, const double *values, const char *dname,
                     double *a_dev, double *da_dev);
void update_potential(Ctx *ctx, int np, const char *const *names, const double *values);
void potential_dvdp(Ctx *ctx, const char *pname, double *out);
}  // namespace nosh
