// keo.cu -- fields (a5-a8) and assembly of the kinetic-energy operator (a9, a10).
//
// HBM layout: one complex (re,im) double2 per block of the owned rows, in the storage
// order chosen at mesh set-up (block-CSR or SELL-32).  The assembly is one thread per
// edge: it evaluates the edge phase once and writes the two off-diagonal blocks
// K_ij = -alpha e^{-i a}, K_ji = conj(K_ij) straight into their precomputed slots --
// no atomics, no column search (the reference does 16 sumIntoGlobalValues column
// searches per edge, src/parameter_matrix_keo.cpp:162-166).  The diagonal
// K_ii = sum_e alpha_e does not depend on the parameters and is written once, when the
// alpha cache is built.
#include "keo.h"

namespace nosh {

namespace {

inline dim3 grid_for(int64_t n, int tpb = 256) { return dim3((unsigned)cdiv(n > 0 ? n : 1, tpb)); }
#define LAUNCH(ctx, kernel, n, ...)                              \
  do {                                                           \
    kernel<<<grid_for(n), 256, 0, (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                           \
    CUDA_CHECK(cudaGetLastError());                              \
  } while (0)

// A = 0.5 B x X  (examples/state-equippers/plain-gl:22-39)
__global__ void k_A_from_curl(const double *coords, double bx, double by, double bz, int64_t n, double *A) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
  A[3 * i] = 0.5 * (by * z - bz * y);
  A[3 * i + 1] = 0.5 * (bz * x - bx * z);
  A[3 * i + 2] = 0.5 * (bx * y - by * x);
}
// cache_e = 0.5 (A_v0 + A_v1) . (x_v0 - x_v1)   (src/vector_field_explicit_values.cpp:31-48)
__global__ void k_edge_cache_explicit(const double *coords, const double *A, const int32_t *edges,
                                      int64_t E, double *cache) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int a = edges[2 * e], b = edges[2 * e + 1];
  const double ax = 0.5 * (A[3 * a] + A[3 * b]), ay = 0.5 * (A[3 * a + 1] + A[3 * b + 1]),
               az = 0.5 * (A[3 * a + 2] + A[3 * b + 2]);
  const double ex = coords[3 * a] - coords[3 * b], ey = coords[3 * a + 1] - coords[3 * b + 1],
               ez = coords[3 * a + 2] - coords[3 * b + 2];
  cache[e] = ax * ex + ay * ey + az * ez;
}
// cache3_e = 0.5 x_v1 x x_v0  (constantCurl edge cache restated so that
// constantCurl(B) == explicit_values(0.5 B x X))
__global__ void k_edge_cache_constcurl(const double *coords, const int32_t *edges, int64_t E, double *c3) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int a = edges[2 * e], b = edges[2 * e + 1];
  const double x0 = coords[3 * a], y0 = coords[3 * a + 1], z0 = coords[3 * a + 2];
  const double x1 = coords[3 * b], y1 = coords[3 * b + 1], z1 = coords[3 * b + 2];
  c3[3 * e] = 0.5 * (y1 * z0 - z1 * y0);
  c3[3 * e + 1] = 0.5 * (z1 * x0 - x1 * z0);
  c3[3 * e + 2] = 0.5 * (x1 * y0 - y1 * x0);
}
// alpha_e = (covolume_e / length_e) * 0.5 (t_v0 + t_v1)  (src/parameter_matrix_keo.cpp:221-228)
__global__ void k_alpha(const int32_t *edges, const double *len, const double *cov, const double *thick,
                        int64_t E, double *alpha) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double a = cov[e] / len[e];
  alpha[e] = a * 0.5 * (thick[edges[2 * e]] + thick[edges[2 * e + 1]]);
}
// K_ii = sum of alpha over the row's edges in ascending column order -- the order in
// which the reference's edge loop reaches them (pure additions => same bits).
__global__ void k_kdiag(const int32_t *rowptr, const int32_t *edge_of, const double *alpha, int64_t No,
                        double *kdiag) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= No) return;
  double s = 0.0;
  for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
    const int e = edge_of[p];
    if (e >= 0) s += alpha[e];
  }
  kdiag[r] = s;
}
__global__ void k_set_diag(const int32_t *diag_slot, const double *kdiag, int64_t No, double2 *val) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= No) return;
  val[diag_slot[r]] = make_double2(kdiag ? kdiag[r] : 0.0, 0.0);
}

struct FillArgs {
  int64_t E;
  const double *alpha;
  const double *cache;  // E (explicit) or 3E (constcurl)
  const int32_t *slot_ij, *slot_ji;
  double2 *val;
  double mu;
  double rb[3], drb[3];  // rotated curl vector and its theta-derivative
  int dmode;             // dKEO: 0 = d/dmu, 1 = d/dtheta, 2 = derivative is zero
};

// MODE 0: KEO (src/parameter_matrix_keo.cpp:106-166), MODE 1: dKEO/dp
// (src/parameter_matrix_dkeo_dp.cpp:96-137).  CC: constantCurl projection.
template <int MODE, bool CC>
__global__ void __launch_bounds__(256) k_keo_fill(FillArgs A) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.E) return;
  double proj, dproj = 0.0;  // a_e / mu and (d a_e/d theta) / mu
  if (CC) {
    const double c0 = A.cache[3 * e], c1 = A.cache[3 * e + 1], c2 = A.cache[3 * e + 2];
    proj = A.rb[0] * c0 + A.rb[1] * c1 + A.rb[2] * c2;
    if (MODE == 1) dproj = A.drb[0] * c0 + A.drb[1] * c1 + A.drb[2] * c2;
  } else {
    proj = A.cache[e];
  }
  const double a = A.mu * proj;
  double s, c;
  sincos(a, &s, &c);
  const double al = A.alpha[e];
  double v0, v1;
  if (MODE == 0) {
    v0 = -c * al;
    v1 = -s * al;
  } else {
    const double da = A.dmode == 0 ? proj : (A.dmode == 1 ? A.mu * dproj : 0.0);
    v0 = (da * s) * al;
    v1 = (-da * c) * al;
  }
  // real 4x4 block of the reference <=> complex K_ij = v0 - i v1, K_ji = v0 + i v1
  const int sij = A.slot_ij[e], sji = A.slot_ji[e];
  if (sij >= 0) A.val[sij] = make_double2(v0, -v1);
  if (sji >= 0) A.val[sji] = make_double2(v0, v1);
}

__global__ void k_fill_const(double *a, int64_t n, double v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_scale_copy(const double *in, double s, int64_t n, double *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = s * in[i];
}
__global__ void k_projection(FillArgs A, int cc, double *a, double *da) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.E) return;
  double proj, dproj = 0.0;
  if (cc) {
    const double c0 = A.cache[3 * e], c1 = A.cache[3 * e + 1], c2 = A.cache[3 * e + 2];
    proj = A.rb[0] * c0 + A.rb[1] * c1 + A.rb[2] * c2;
    dproj = A.drb[0] * c0 + A.drb[1] * c1 + A.drb[2] * c2;
  } else {
    proj = A.cache[e];
  }
  if (a) a[e] = A.mu * proj;
  if (da) da[e] = A.dmode == 0 ? proj : (A.dmode == 1 ? A.mu * dproj : 0.0);
}

// rotate_ / dRotateDTheta_ of src/vector_field_constant_curl.cpp:144-197 (3 scalars, host)
void rotate_b(const Ctx *ctx, double theta, double rb[3], double drb[3]) {
  const double *b = ctx->cc_b, *u = ctx->cc_u;
  for (int i = 0; i < 3; i++) rb[i] = drb[i] = b[i];
  if (!ctx->cc_has_u) return;
  double s, c;
  sincos(theta, &s, &c);
  const double ub = u[0] * b[0] + u[1] * b[1] + u[2] * b[2];
  const double uxb[3] = {u[1] * b[2] - u[2] * b[1], u[2] * b[0] - u[0] * b[2], u[0] * b[1] - u[1] * b[0]};
  for (int i = 0; i < 3; i++) {
    if (s != 0.0) rb[i] = c * b[i] + s * uxb[i] + ((1.0 - c) * ub) * u[i];
    drb[i] = -s * b[i] + c * uxb[i] + ((1.0 + s) * ub) * u[i];
  }
}

FillArgs make_fill_args(Ctx *ctx, double mu, double theta, double2 *val) {
  FillArgs A;
  A.E = ctx->E;
  A.alpha = ctx->alpha.p;
  A.cache = ctx->ecache.p;
  A.slot_ij = ctx->slot_ij.p;
  A.slot_ji = ctx->slot_ji.p;
  A.val = val;
  A.mu = mu;
  A.dmode = 0;
  rotate_b(ctx, theta, A.rb, A.drb);
  return A;
}

void require_fields(Ctx *ctx) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  if (ctx->mvp_kind == MVP_NONE) NOSH_THROW(NOSH_ESTATE, "no magnetic vector potential set");
  if (!ctx->thick_set) NOSH_THROW(NOSH_ESTATE, "no thickness field set");
}

}  // namespace

void set_thickness(Ctx *ctx, const double *values, double c) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  ctx->thick.alloc(ctx->Nl);
  if (values)
    CUDA_CHECK(cudaMemcpyAsync(ctx->thick.p, values, sizeof(double) * ctx->Nl, cudaMemcpyHostToDevice,
                               ctx->stream));
  else
    LAUNCH(ctx, k_fill_const, ctx->Nl, ctx->thick.p, ctx->Nl, c);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->thick_set = true;
  ctx->alpha_ok = false;
  ctx->keo_filled = ctx->dkeo_filled = false;
  ctx->amg_valid = false;  // a new field invalidates the whole hierarchy, not only its finest diagonal
}

void set_mvp_explicit(Ctx *ctx, const double *A_host, const double *B) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  DBuf<double> A;
  A.alloc(ctx->Nl * 3);
  if (A_host)
    CUDA_CHECK(cudaMemcpyAsync(A.p, A_host, sizeof(double) * ctx->Nl * 3, cudaMemcpyHostToDevice, ctx->stream));
  else
    LAUNCH(ctx, k_A_from_curl, ctx->Nl, ctx->coords.p, B[0], B[1], B[2], ctx->Nl, A.p);
  ctx->ecache.alloc(ctx->E);
  LAUNCH(ctx, k_edge_cache_explicit, ctx->E, ctx->coords.p, A.p, ctx->edges.p, ctx->E, ctx->ecache.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->mvp_kind = MVP_EXPLICIT;
  ctx->cc_has_u = false;
  ctx->keo_filled = ctx->dkeo_filled = false;
  ctx->amg_valid = false;  // a new field invalidates the whole hierarchy, not only its finest diagonal
}

void set_mvp_constcurl(Ctx *ctx, const double b[3], const double u[3]) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  // src/vector_field_constant_curl.cpp:35-44: exact normalisation is required
  if (b[0] * b[0] + b[1] * b[1] + b[2] * b[2] != 1.0)
    NOSH_THROW(NOSH_EINVAL, "Curl vector not normalized: <b,b> = %.17g.", b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  if (u && u[0] * u[0] + u[1] * u[1] + u[2] * u[2] != 1.0)
    NOSH_THROW(NOSH_EINVAL, "Rotation vector not normalized: <u,u> = %.17g.", u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  for (int i = 0; i < 3; i++) {
    ctx->cc_b[i] = b[i];
    ctx->cc_u[i] = u ? u[i] : 0.0;
  }
  ctx->cc_has_u = u != nullptr;
  ctx->ecache.alloc(ctx->E * 3);
  LAUNCH(ctx, k_edge_cache_constcurl, ctx->E, ctx->coords.p, ctx->edges.p, ctx->E, ctx->ecache.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->mvp_kind = MVP_CONSTCURL;
  ctx->keo_filled = ctx->dkeo_filled = false;
  ctx->amg_valid = false;  // a new field invalidates the whole hierarchy, not only its finest diagonal
}

// keo::build_alpha_cache_ (first refill only) + the parameter-independent diagonal
void ensure_alpha(Ctx *ctx) {
  require_fields(ctx);
  if (ctx->alpha_ok) return;
  ctx->alpha.alloc(ctx->E);
  LAUNCH(ctx, k_alpha, ctx->E, ctx->edges.p, ctx->elen.p, ctx->ecov.p, ctx->thick.p, ctx->E, ctx->alpha.p);
  ctx->Kdiag.alloc(ctx->No);
  LAUNCH(ctx, k_kdiag, ctx->No, ctx->rowptr.p, ctx->edge_of.p, ctx->alpha.p, ctx->No, ctx->Kdiag.p);
  ctx->Kval.alloc(ctx->nstored);
  CUDA_CHECK(cudaMemsetAsync(ctx->Kval.p, 0, sizeof(double2) * ctx->nstored, ctx->stream));
  LAUNCH(ctx, k_set_diag, ctx->No, ctx->diag_slot.p, ctx->Kdiag.p, ctx->No, ctx->Kval.p);
  ctx->alpha_ok = true;
  ctx->keo_filled = ctx->dkeo_filled = false;
}

static void mvp_params(Ctx *ctx, int np, const char *const *names, const double *values, double *mu,
                       double *theta) {
  *mu = param_at(np, names, values, "mu");  // set_parameters: params.at("mu")
  *theta = 0.0;
  if (ctx->mvp_kind == MVP_CONSTCURL) *theta = param_at(np, names, values, "theta");
}

void keo_fill(Ctx *ctx, int np, const char *const *names, const double *values, bool force) {
  ensure_alpha(ctx);
  double mu, theta;
  mvp_params(ctx, np, names, values, &mu, &theta);
  if (!force && ctx->keo_filled && ctx->keo_mu == mu && ctx->keo_theta == theta) return;
  FillArgs A = make_fill_args(ctx, mu, theta, ctx->Kval.p);
  if (ctx->mvp_kind == MVP_CONSTCURL)
    LAUNCH(ctx, (k_keo_fill<0, true>), ctx->E, A);
  else
    LAUNCH(ctx, (k_keo_fill<0, false>), ctx->E, A);
  ctx->keo_filled = true;
  ctx->keo_mu = mu;
  ctx->keo_theta = theta;
  ctx->kval_version++;
  ctx->keoreg_version++;  // the regularised KEO shares these values: a kept AMG hierarchy refreshes its finest level
}

static int dmode_of(Ctx *ctx, const char *dname) {
  if (!dname) NOSH_THROW(NOSH_EINVAL, "derivative parameter name is NULL");
  if (strcmp(dname, "mu") == 0) return 0;
  if (ctx->mvp_kind == MVP_CONSTCURL) {
    if (strcmp(dname, "theta") == 0) return 1;
    // src/vector_field_constant_curl.cpp:135-139 throws
    NOSH_THROW(NOSH_EINVAL, "Illegal parameter \"%s\".", dname);
  }
  return 2;  // explicit_values: 0.0 for any other name (src/vector_field_explicit_values.cpp:84-89)
}

void dkeo_fill(Ctx *ctx, int np, const char *const *names, const double *values, const char *dname) {
  ensure_alpha(ctx);
  double mu, theta;
  mvp_params(ctx, np, names, values, &mu, &theta);
  const int dmode = dmode_of(ctx, dname);
  if (ctx->dkeo_filled && ctx->dkeo_mu == mu && ctx->dkeo_theta == theta && ctx->dkeo_name == dname) return;
  if (!ctx->dKval.p) {
    ctx->dKval.alloc(ctx->nstored);
    CUDA_CHECK(cudaMemsetAsync(ctx->dKval.p, 0, sizeof(double2) * ctx->nstored, ctx->stream));
  }
  FillArgs A = make_fill_args(ctx, mu, theta, ctx->dKval.p);
  A.dmode = dmode;
  if (ctx->mvp_kind == MVP_CONSTCURL)
    LAUNCH(ctx, (k_keo_fill<1, true>), ctx->E, A);
  else
    LAUNCH(ctx, (k_keo_fill<1, false>), ctx->E, A);
  ctx->dkeo_filled = true;
  ctx->dkeo_mu = mu;
  ctx->dkeo_theta = theta;
  ctx->dkeo_name = dname;
}

void edge_projection(Ctx *ctx, int np, const char *const *names, const double *values, const char *dname,
                     double *a_dev, double *da_dev) {
  require_fields(ctx);
  double mu, theta;
  mvp_params(ctx, np, names, values, &mu, &theta);
  FillArgs A = make_fill_args(ctx, mu, theta, nullptr);
  A.alpha = nullptr;
  A.dmode = dname ? dmode_of(ctx, dname) : 0;
  LAUNCH(ctx, k_projection, ctx->E, A, ctx->mvp_kind == MVP_CONSTCURL ? 1 : 0, a_dev, dname ? da_dev : nullptr);
}

// scalar potential for the current parameters (a7): src/scalar_field_constant.cpp:43-58,
// src/scalar_field_explicit_values.cpp:33-41
void update_potential(Ctx *ctx, int np, const char *const *names, const double *values) {
  if (ctx->pot_kind == POT_NONE) NOSH_THROW(NOSH_ESTATE, "no scalar potential set");
  ctx->Vcur.ensure(ctx->No);
  if (ctx->pot_kind == POT_CONSTANT) {
    double v = ctx->pot_c;
    const double *p = ctx->pot_param.empty() ? nullptr : find_param(np, names, values, ctx->pot_param.c_str());
    if (p) v = ctx->pot_c + *p;
    LAUNCH(ctx, k_fill_const, ctx->No, ctx->Vcur.p, ctx->No, v);
  } else {
    const double beta = param_at(np, names, values, "beta");
    LAUNCH(ctx, k_scale_copy, ctx->No, ctx->pot_values.p, beta, ctx->No, ctx->Vcur.p);
  }
}

// dV/dp into Vcur-sized buffer `out` (src/scalar_field_constant.cpp:60-75,
// src/scalar_field_explicit_values.cpp:43-62)
void potential_dvdp(Ctx *ctx, const char *pname, double *out) {
  if (ctx->pot_kind == POT_NONE) NOSH_THROW(NOSH_ESTATE, "no scalar potential set");
  if (ctx->pot_kind == POT_CONSTANT) {
    const double v = (!ctx->pot_param.empty() && ctx->pot_param == pname) ? 1.0 : 0.0;
    LAUNCH(ctx, k_fill_const, ctx->No, out, ctx->No, v);
  } else {
    if (strcmp(pname, "beta") == 0)
      LAUNCH(ctx, k_scale_copy, ctx->No, ctx->pot_values.p, 1.0, ctx->No, out);
    else
      LAUNCH(ctx, k_fill_const, ctx->No, out, ctx->No, 0.0);
  }
}

}  // namespace nosh
