// capi.cu -- the extern "C" boundary declared in include/nosh_b200.h.
// Thin: argument checks, host<->device staging, status codes.  Never throws.
#include <cstdlib>

#include "amg.h"
#include "apply.cuh"
#include "comm.h"
#include "fvm.h"
#include "keo.h"
#include "krylov.h"
#include "mesh.h"

struct nosh_ctx : public nosh::Ctx {};

using namespace nosh;

namespace {

#define API_BEGIN(ctx)            \
  if (!(ctx)) return NOSH_EINVAL; \
  try {
#define API_END(ctx)                        \
  }                                         \
  catch (const nosh::Exception &e) {        \
    (ctx)->err = e.msg;                     \
    return e.code;                          \
  }                                         \
  catch (const std::exception &e) {         \
    (ctx)->err = e.what();                  \
    return NOSH_EINVAL;                     \
  }                                         \
  return NOSH_OK;

double wall_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

void require_mesh(Ctx *ctx) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
}

// input vector of 2*No doubles -> device pointer with room for ghosts when needed
// does an operator apply need its input to carry a ghost segment?  With peer memory the ghosts are read from
// the landing buffer (krylov.cu:apply_halo_dev), so a device vector of the caller is used in place.
bool ghost_room(const Ctx *ctx) { return ctx->nranks > 1 && !ctx->p2p.ok; }

void ensure_copy_stream(Ctx *ctx) {
  if (ctx->copy_stream) return;
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
  for (auto &pf : ctx->prefetch) CUDA_CHECK(cudaEventCreateWithFlags(&pf.ready, cudaEventDisableTiming));
  for (auto &e : ctx->out_done) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

double2 *stage_in(Ctx *ctx, const double *p, DBuf<double2> &buf, bool need_ghost_room) {
  if (!p && ctx->No > 0) NOSH_THROW(NOSH_EINVAL, "NULL vector");
  const bool dev = p && is_device_ptr(p);
  if (dev && !need_ghost_room) return (double2 *)p;
  if (!dev && ctx->copy_stream)
    for (auto &pf : ctx->prefetch)
      if (pf.valid && pf.host == p) {  // already on its way (nosh_prefetch): wait for it instead of copying
        pf.valid = false;
        CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, pf.ready, 0));
        return pf.dev.p;
      }
  buf.ensure(ctx->Nl > 0 ? ctx->Nl : 1);
  if (ctx->No > 0)
    CUDA_CHECK(cudaMemcpyAsync(buf.p, p, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               ctx->stream));
  return buf.p;
}
struct OutVec {
  double2 *dev;
  double *user;
  bool host;
  int slot;
};
OutVec stage_out(Ctx *ctx, double *p, DBuf<double2> &buf) {
  if (!p && ctx->No > 0) NOSH_THROW(NOSH_EINVAL, "NULL vector");
  OutVec o;
  o.user = p;
  o.host = !p || !is_device_ptr(p);
  o.slot = 0;
  if (o.host) {
    DBuf<double2> *b = &buf;
    if (ctx->async_output && &buf == &ctx->stage_y) {
      // the previous result may still be travelling to the host from the other buffer: alternate, and do not
      // overwrite a buffer before its own copy has finished
      ensure_copy_stream(ctx);
      o.slot = ctx->out_flip;
      ctx->out_flip ^= 1;
      if (o.slot == 1) b = &ctx->stage_y2;
      CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->out_done[o.slot], 0));
    }
    b->ensure(ctx->Nl > 0 ? ctx->Nl : 1);
    o.dev = b->p;
  } else {
    o.dev = (double2 *)p;
  }
  return o;
}
void finish_out(Ctx *ctx, const OutVec &o) {
  if (o.host && ctx->No > 0) {
    if (ctx->async_output && ctx->copy_stream) {
      // pipelined: the D2H rides on the copy stream behind the producing work; nosh_ctx_synchronize completes it
      CUDA_CHECK(cudaEventRecord(ctx->ev_order, ctx->stream));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_order, 0));
      CUDA_CHECK(cudaMemcpyAsync(o.user, o.dev, sizeof(double2) * ctx->No, cudaMemcpyDeviceToHost, ctx->copy_stream));
      CUDA_CHECK(cudaEventRecord(ctx->out_done[o.slot], ctx->copy_stream));
      return;
    }
    CUDA_CHECK(cudaMemcpyAsync(o.user, o.dev, sizeof(double2) * ctx->No, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}
template <typename T>
void d2h(Ctx *ctx, T *host, const T *dev, size_t n) {
  if (!host || n == 0) return;
  CUDA_CHECK(cudaMemcpyAsync(host, dev, sizeof(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

__global__ void k_gather_vals(const int32_t *pos, const double2 *val, int64_t n, double2 *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = val[pos[i]];
}
__global__ void k_gather_real(const int32_t *pos, const double *val, int64_t n, double *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = val[pos[i]];
}
__global__ void k_widen(const int32_t *in, int64_t n, int64_t *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

void after_mesh(Ctx *ctx) {
  amg_free(ctx);
  halo_setup(ctx);
  ensure_work(ctx);
  p2p_setup(ctx);
  ctx->thick_set = false;
  ctx->mvp_kind = MVP_NONE;
  ctx->alpha_ok = ctx->keo_filled = ctx->dkeo_filled = ctx->jac_ok = ctx->keoreg_ok = false;
  ctx->fvm_filled = false;
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void check_apply_shape(Ctx *ctx, int64_t ldx, int64_t ldy, int nvec) {
  if (nvec < 1) NOSH_THROW(NOSH_EINVAL, "nvec < 1");
  if (nvec > 1 && (ldx < 2 * ctx->No || ldy < 2 * ctx->No)) NOSH_THROW(NOSH_EINVAL, "leading dimension too small");
}

}  // namespace

extern "C" {

const char *nosh_version(void) { return "nosh_b200 0.1 (sm_100a)"; }

nosh_status nosh_ctx_create(int device, void *stream, nosh_ctx **out) {
  if (!out) return NOSH_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return NOSH_ECUDA;  // no CPU fallback
  if (device < 0 || device >= ndev) return NOSH_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return NOSH_ECUDA;
  nosh_ctx *ctx = new nosh_ctx;
  ctx->device = device;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return NOSH_ECUDA;
    }
    ctx->own_stream = true;
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  if (const char *e = getenv("NOSH_B200_LAYOUT")) {
    if (strcmp(e, "csr") == 0) ctx->layout = NOSH_LAYOUT_CSR;
    if (strcmp(e, "sell32") == 0) ctx->layout = NOSH_LAYOUT_SELL32;
  }
  if (const char *e = getenv("NOSH_B200_GROUP")) {
    const long long g = atoll(e);
    if (g >= CHUNK && g % CHUNK == 0) ctx->group_vertices = g;
  }
  if (const char *e = getenv("NOSH_B200_PERSISTENT_MINRES")) ctx->persistent_minres = atoi(e) != 0;
  if (const char *e = getenv("NOSH_B200_PERSISTENT_MGPU")) ctx->persistent_mgpu = atoi(e) != 0;
  if (const char *e = getenv("NOSH_B200_MGPU_FENCE")) ctx->mgpu_fence = atoi(e);
  if (const char *e = getenv("NOSH_B200_MGPU_LEAN")) ctx->mgpu_lean = atoi(e) != 0;
  if (const char *e = getenv("NOSH_B200_AMG_GRAPH")) ctx->amg_graph = atoi(e) != 0;
  if (const char *e = getenv("NOSH_B200_AMG_PANEL_PRODUCTS"))
    if (atoll(e) > 0) ctx->amg_panel_products = atoll(e);
  if (const char *e = getenv("NOSH_B200_SELL_SIGMA")) ctx->sell_sigma = atoi(e) < 0 ? -1 : (atoi(e) != 0);
  *out = ctx;
  return NOSH_OK;
}

void nosh_ctx_destroy(nosh_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  amg_free(ctx);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_order);
    for (auto &pf : ctx->prefetch) cudaEventDestroy(pf.ready);
    for (auto &e : ctx->out_done) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->copy_stream);
  }
  cudaStream_t s = ctx->own_stream ? ctx->stream : nullptr;
  delete ctx;
  if (s) cudaStreamDestroy(s);
}

const char *nosh_last_error(const nosh_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

nosh_status nosh_ctx_set_layout(nosh_ctx *ctx, nosh_layout layout) {
  API_BEGIN(ctx)
  if (ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "layout must be chosen before the mesh is set");
  if (layout != NOSH_LAYOUT_CSR && layout != NOSH_LAYOUT_SELL32) NOSH_THROW(NOSH_EINVAL, "unknown layout");
  ctx->layout = layout;
  API_END(ctx)
}

nosh_status nosh_ctx_set_group_vertices(nosh_ctx *ctx, int64_t g) {
  API_BEGIN(ctx)
  if (ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "group size must be chosen before the mesh is set");
  if (g < CHUNK || g % CHUNK) NOSH_THROW(NOSH_EINVAL, "group_vertices must be a positive multiple of %d", CHUNK);
  ctx->group_vertices = g;
  API_END(ctx)
}

nosh_status nosh_ctx_synchronize(nosh_ctx *ctx) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_stream) CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
  API_END(ctx)
}

nosh_status nosh_prefetch(nosh_ctx *ctx, const double *host_vector) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!host_vector) NOSH_THROW(NOSH_EINVAL, "NULL vector");
  if (is_device_ptr(host_vector) || ctx->No == 0) return NOSH_OK;  // nothing to stage
  ensure_copy_stream(ctx);
  Ctx::Prefetch *slot = nullptr;
  for (auto &pf : ctx->prefetch)
    if (pf.valid && pf.host == host_vector) slot = &pf;  // refresh an entry that was never consumed
  if (!slot) {
    slot = &ctx->prefetch[ctx->prefetch_next];
    ctx->prefetch_next = (ctx->prefetch_next + 1) % 4;
  }
  slot->dev.ensure(ctx->Nl > 0 ? ctx->Nl : 1);
  // the buffer's last consumer was enqueued on the compute stream before this call: stay behind it
  CUDA_CHECK(cudaEventRecord(ctx->ev_order, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_order, 0));
  CUDA_CHECK(cudaMemcpyAsync(slot->dev.p, host_vector, sizeof(double2) * ctx->No, cudaMemcpyHostToDevice, ctx->copy_stream));
  CUDA_CHECK(cudaEventRecord(slot->ready, ctx->copy_stream));
  slot->host = host_vector;
  slot->valid = true;
  API_END(ctx)
}

nosh_status nosh_ctx_set_async_output(nosh_ctx *ctx, int enabled) {
  API_BEGIN(ctx)
  if (!enabled && ctx->copy_stream) CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
  ctx->async_output = enabled != 0;
  if (enabled) ensure_copy_stream(ctx);
  API_END(ctx)
}

nosh_status nosh_comm_unique_id(void *id128) {
  if (!id128) return NOSH_EINVAL;
  try {
    comm_unique_id(id128);
  } catch (const nosh::Exception &e) {
    return e.code;
  }
  return NOSH_OK;
}

nosh_status nosh_partition_range(int64_t n_global, int nranks, int rank, int64_t group_vertices, int64_t *begin,
                                 int64_t *end, int64_t *group_used) {
  if (n_global <= 0 || nranks < 1 || rank < 0 || rank >= nranks || group_vertices < CHUNK ||
      group_vertices % CHUNK)
    return NOSH_EINVAL;
  partition_range(n_global, nranks, rank, group_vertices, begin, end, group_used, nullptr, nullptr, nullptr);
  return NOSH_OK;
}

nosh_status nosh_ctx_comm_init(nosh_ctx *ctx, const void *id128, int rank, int nranks) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  comm_init(ctx, id128, rank, nranks);
  API_END(ctx)
}

nosh_status nosh_ctx_comm_init_host(nosh_ctx *ctx, int rank, int nranks, nosh_allgather_fn allgather, void *user) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  comm_init_host(ctx, rank, nranks, allgather, user);
  API_END(ctx)
}

nosh_status nosh_ctx_get_stat(nosh_ctx *ctx, const char *key, double *value) {
  API_BEGIN(ctx)
  if (!key || !value) NOSH_THROW(NOSH_EINVAL, "NULL argument");
  if (strcmp(key, "p2p") == 0) {
    *value = ctx->p2p.ok ? 1.0 : 0.0;
  } else if (strcmp(key, "n_chunks_interior") == 0) {
    *value = (double)ctx->n_chunks_int;
  } else if (strcmp(key, "n_chunks_boundary") == 0) {
    *value = (double)ctx->n_chunks_bnd;
  } else if (strcmp(key, "n_send") == 0) {
    *value = (double)ctx->n_send;
  } else {
    auto it = ctx->stats.find(key);
    if (it == ctx->stats.end()) NOSH_THROW(NOSH_EKEY, "unknown stat \"%s\"", key);
    *value = it->second;
  }
  API_END(ctx)
}

// all recorded stats as "key=value\n" lines (NUL terminated, truncated to cap)
nosh_status nosh_ctx_list_stats(nosh_ctx *ctx, char *buf, int64_t cap) {
  API_BEGIN(ctx)
  if (!buf || cap < 1) NOSH_THROW(NOSH_EINVAL, "NULL buffer");
  std::string out;
  char line[256];
  for (const auto &kv : ctx->stats) {
    snprintf(line, sizeof(line), "%s=%.9g\n", kv.first.c_str(), kv.second);
    out += line;
  }
  const size_t n = std::min<size_t>(out.size(), (size_t)cap - 1);
  memcpy(buf, out.data(), n);
  buf[n] = 0;
  API_END(ctx)
}

nosh_status nosh_mesh_set(nosh_ctx *ctx, int dim, int64_t nv, const double *coords, int64_t nc,
                          const int32_t *cells) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->has_mesh = false;
  const double t0 = wall_s();
  mesh_from_host(ctx, dim, nv, coords, nc, cells);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->stats["setup.mesh_s"] = wall_s() - t0;
  after_mesh(ctx);
  API_END(ctx)
}

nosh_status nosh_mesh_set_local(nosh_ctx *ctx, int dim, int64_t n_global, int64_t nv_local, const int64_t *vertex_gids,
                                const double *coords, int64_t nc_local, const int32_t *cells) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->has_mesh = false;
  const double t0 = wall_s();
  mesh_from_host_local(ctx, dim, n_global, nv_local, vertex_gids, coords, nc_local, cells);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->stats["setup.mesh_s"] = wall_s() - t0;
  after_mesh(ctx);
  API_END(ctx)
}

nosh_status nosh_mesh_tetgrid(nosh_ctx *ctx, int nx, int ny, int nz, const double lo[3], const double hi[3],
                              double jitter, uint64_t seed) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!lo || !hi) NOSH_THROW(NOSH_EINVAL, "NULL bounds");
  ctx->has_mesh = false;
  const double t0 = wall_s();
  mesh_tetgrid(ctx, nx, ny, nz, lo, hi, jitter, seed);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->stats["setup.mesh_s"] = wall_s() - t0;
  after_mesh(ctx);
  API_END(ctx)
}

nosh_status nosh_mesh_info(const nosh_ctx *ctx, nosh_mesh_info_t *info) {
  if (!ctx || !info) return NOSH_EINVAL;
  if (!ctx->has_mesh) return NOSH_ESTATE;
  info->dim = ctx->dim;
  info->n_global = ctx->n_global;
  info->owned_begin = ctx->vb;
  info->n_owned = ctx->No;
  info->n_ghost = ctx->Ng;
  info->n_cells = ctx->nc;
  info->n_edges = ctx->E;
  info->n_blocks = ctx->nb;
  info->n_stored = ctx->nstored;
  return NOSH_OK;
}

nosh_status nosh_mesh_local_gids(nosh_ctx *ctx, int64_t *gids) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  DBuf<int64_t> w;
  w.alloc(ctx->Nl);
  k_widen<<<(unsigned)cdiv(ctx->Nl, 256), 256, 0, ctx->stream>>>(ctx->gid.p, ctx->Nl, w.p);
  ctx->launches++;
  d2h(ctx, gids, w.p, ctx->Nl);
  API_END(ctx)
}

nosh_status nosh_mesh_get_coords(nosh_ctx *ctx, double *coords) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  d2h(ctx, coords, ctx->coords.p, ctx->Nl * 3);
  API_END(ctx)
}

nosh_status nosh_mesh_get_cells(nosh_ctx *ctx, int32_t *cells) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  d2h(ctx, cells, ctx->cells.p, ctx->nc * (ctx->dim + 1));
  API_END(ctx)
}

nosh_status nosh_mesh_get_edges(nosh_ctx *ctx, int32_t *edges, double *length, double *covolume) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  d2h(ctx, edges, ctx->edges.p, ctx->E * 2);
  d2h(ctx, length, ctx->elen.p, ctx->E);
  d2h(ctx, covolume, ctx->ecov.p, ctx->E);
  API_END(ctx)
}

nosh_status nosh_mesh_get_control_volumes(nosh_ctx *ctx, double *cv) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  d2h(ctx, cv, ctx->cv.p, ctx->No);
  API_END(ctx)
}

nosh_status nosh_set_thickness(nosh_ctx *ctx, const double *values, double c) {
  API_BEGIN(ctx)
  set_thickness(ctx, values, c);
  API_END(ctx)
}

nosh_status nosh_set_potential_constant(nosh_ctx *ctx, double c, const char *param1_name) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ctx->pot_kind = POT_CONSTANT;
  ctx->pot_c = c;
  ctx->pot_param = param1_name ? param1_name : "";
  API_END(ctx)
}

nosh_status nosh_set_potential_values(nosh_ctx *ctx, const double *values) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!values) NOSH_THROW(NOSH_EINVAL, "NULL values");
  ctx->pot_values.alloc(ctx->No);
  CUDA_CHECK(cudaMemcpyAsync(ctx->pot_values.p, values, sizeof(double) * ctx->No, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->pot_kind = POT_VALUES;
  API_END(ctx)
}

nosh_status nosh_set_mvp_explicit(nosh_ctx *ctx, const double *A) {
  API_BEGIN(ctx)
  if (!A) NOSH_THROW(NOSH_EINVAL, "NULL A");
  set_mvp_explicit(ctx, A, nullptr);
  API_END(ctx)
}

nosh_status nosh_set_mvp_explicit_curl(nosh_ctx *ctx, const double B[3]) {
  API_BEGIN(ctx)
  if (!B) NOSH_THROW(NOSH_EINVAL, "NULL B");
  set_mvp_explicit(ctx, nullptr, B);
  API_END(ctx)
}

nosh_status nosh_set_mvp_constcurl(nosh_ctx *ctx, const double b[3], const double u[3]) {
  API_BEGIN(ctx)
  if (!b) NOSH_THROW(NOSH_EINVAL, "NULL b");
  set_mvp_constcurl(ctx, b, u);
  API_END(ctx)
}

nosh_status nosh_get_alpha_cache(nosh_ctx *ctx, double *alpha) {
  API_BEGIN(ctx)
  ensure_alpha(ctx);
  d2h(ctx, alpha, ctx->alpha.p, ctx->E);
  API_END(ctx)
}

nosh_status nosh_get_edge_projection(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                                     const char *dname, double *a, double *da) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  DBuf<double> da_, a_;
  a_.alloc(ctx->E);
  da_.alloc(ctx->E);
  edge_projection(ctx, np, names, values, dname, a_.p, da_.p);
  d2h(ctx, a, a_.p, ctx->E);
  if (dname) d2h(ctx, da, da_.p, ctx->E);
  API_END(ctx)
}

nosh_status nosh_keo_fill(nosh_ctx *ctx, int np, const char *const *names, const double *values) {
  API_BEGIN(ctx)
  // the reference always refills (its cache never hits, src/parameter_object.cpp:17-44); an
  // explicit fill request is honoured even for unchanged parameters
  keo_fill(ctx, np, names, values, true);
  API_END(ctx)
}

nosh_status nosh_dkeo_fill(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                           const char *dname) {
  API_BEGIN(ctx)
  ctx->dkeo_filled = false;
  dkeo_fill(ctx, np, names, values, dname);
  API_END(ctx)
}

nosh_status nosh_matrix_apply(nosh_ctx *ctx, nosh_matrix_id which, const double *X, int64_t ldx, double *Y,
                              int64_t ldy, int nvec, nosh_transp mode, double alpha, double beta) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  (void)mode;  // Hermitian: symmetric in the real layout, every mode is the same operator
  check_apply_shape(ctx, ldx, ldy, nvec);
  const double2 *val = nullptr;
  if (which == NOSH_MAT_KEO) {
    if (!ctx->keo_filled) NOSH_THROW(NOSH_ESTATE, "KEO not filled");
    val = ctx->Kval.p;
  } else if (which == NOSH_MAT_DKEO) {
    if (!ctx->dkeo_filled) NOSH_THROW(NOSH_ESTATE, "dKEO not filled");
    val = ctx->dKval.p;
  } else {
    NOSH_THROW(NOSH_EINVAL, "unknown matrix id");
  }
  ensure_work(ctx);
  for (int v = 0; v < nvec; v++) {
    double2 *x = stage_in(ctx, X + (size_t)v * ldx, ctx->stage_x, ghost_room(ctx));
    OutVec o = stage_out(ctx, Y + (size_t)v * ldy, ctx->stage_y);
    if (o.host && beta != 0.0)
      CUDA_CHECK(cudaMemcpyAsync(o.dev, o.user, sizeof(double2) * ctx->No, cudaMemcpyHostToDevice, ctx->stream));
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
    A.y = o.dev;
    A.a = alpha;
    A.b = beta;
    apply_halo_dev(ctx, EPI_NONE, (alpha == 1.0 && beta == 0.0) ? FUSE_NONE : FUSE_AXPBY, A, x);
    finish_out(ctx, o);
  }
  API_END(ctx)
}

nosh_status nosh_get_block_csr(nosh_ctx *ctx, nosh_matrix_id which, int64_t *rowptr, int32_t *cols,
                               double *vals) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (rowptr) {
    DBuf<int64_t> w;
    w.alloc(ctx->No + 1);
    k_widen<<<(unsigned)cdiv(ctx->No + 1, 256), 256, 0, ctx->stream>>>(ctx->rowptr.p, ctx->No + 1, w.p);
    ctx->launches++;
    d2h(ctx, rowptr, w.p, ctx->No + 1);
  }
  d2h(ctx, cols, ctx->csr_col.p, ctx->nb);
  if (vals) {
    const double2 *val = which == NOSH_MAT_KEO ? ctx->Kval.p : ctx->dKval.p;
    if (which == NOSH_MAT_KEO && !ctx->keo_filled) NOSH_THROW(NOSH_ESTATE, "KEO not filled");
    if (which == NOSH_MAT_DKEO && !ctx->dkeo_filled) NOSH_THROW(NOSH_ESTATE, "dKEO not filled");
    DBuf<double2> g;
    g.alloc(ctx->nb);
    if (ctx->nb) {
      k_gather_vals<<<(unsigned)cdiv(ctx->nb, 256), 256, 0, ctx->stream>>>(ctx->csr_pos.p, val, ctx->nb, g.p);
      ctx->launches++;
    }
    d2h(ctx, (double2 *)vals, g.p, ctx->nb);
  }
  API_END(ctx)
}

nosh_status nosh_jac_rebuild(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                             const double *psi) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  keo_fill(ctx, np, names, values, false);            // keo_->set_parameters (src/jacobian_operator.cpp:135)
  const double g = param_at(np, names, values, "g");  // :156
  update_potential(ctx, np, names, values);
  double2 *x = stage_in(ctx, psi, ctx->stage_x, false);
  jac_diags_dev(ctx, g, x);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_jac_apply(nosh_ctx *ctx, const double *X, int64_t ldx, double *Y, int64_t ldy, int nvec,
                           nosh_transp mode, double alpha, double beta) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  // src/jacobian_operator.cpp:48-59
  if (mode != NOSH_NO_TRANS) NOSH_THROW(NOSH_EINVAL, "Only untransposed applies supported.");
  if (alpha != 1.0) NOSH_THROW(NOSH_EINVAL, "Only alpha==1.0 supported.");
  if (beta != 0.0) NOSH_THROW(NOSH_EINVAL, "Only beta==0.0 supported.");
  check_apply_shape(ctx, ldx, ldy, nvec);
  for (int v = 0; v < nvec; v++) {
    double2 *x = stage_in(ctx, X + (size_t)v * ldx, ctx->stage_x, ghost_room(ctx));
    OutVec o = stage_out(ctx, Y + (size_t)v * ldy, ctx->stage_y);
    apply_op_dev(ctx, NOSH_OP_JACOBIAN, x, o.dev);
    finish_out(ctx, o);
  }
  API_END(ctx)
}

nosh_status nosh_jac_get_diags(nosh_ctx *ctx, double *d0, double *d1b) {
  API_BEGIN(ctx)
  if (!ctx->jac_ok) NOSH_THROW(NOSH_ESTATE, "Jacobian not built");
  d2h(ctx, (double2 *)d0, ctx->jd0.p, ctx->No);
  d2h(ctx, d1b, ctx->jd1.p, ctx->No);
  API_END(ctx)
}

nosh_status nosh_compute_f(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                           const double *psi, double *f) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  keo_fill(ctx, np, names, values, false);            // src/model_evaluator_nls.cpp:536
  const double g = param_at(np, names, values, "g");  // :559
  update_potential(ctx, np, names, values);
  double2 *x = stage_in(ctx, psi, ctx->stage_x, ghost_room(ctx));
  OutVec o = stage_out(ctx, f, ctx->stage_y);
  compute_f_dev(ctx, g, x, o.dev);
  finish_out(ctx, o);
  API_END(ctx)
}

nosh_status nosh_compute_dfdp(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                              const char *pname, const double *psi, double *dfdp) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!pname) NOSH_THROW(NOSH_EINVAL, "NULL parameter name");
  double2 *x = stage_in(ctx, psi, ctx->stage_x, ghost_room(ctx));
  OutVec o = stage_out(ctx, dfdp, ctx->stage_y);
  compute_dfdp_dev(ctx, np, names, values, pname, x, o.dev);
  finish_out(ctx, o);
  API_END(ctx)
}

nosh_status nosh_keoreg_rebuild(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                                const double *psi) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  keo_fill(ctx, np, names, values, false);            // src/keo_regularized.cpp:196
  const double g = param_at(np, names, values, "g");  // :198
  double2 *x = stage_in(ctx, psi, ctx->stage_x, false);
  keoreg_diags_dev(ctx, g, x);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_keoreg_matrix_apply(nosh_ctx *ctx, const double *X, int64_t ldx, double *Y, int64_t ldy,
                                     int nvec) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  check_apply_shape(ctx, ldx, ldy, nvec);
  for (int v = 0; v < nvec; v++) {
    double2 *x = stage_in(ctx, X + (size_t)v * ldx, ctx->stage_x, ghost_room(ctx));
    OutVec o = stage_out(ctx, Y + (size_t)v * ldy, ctx->stage_y);
    apply_op_dev(ctx, NOSH_OP_KEOREG, x, o.dev);
    finish_out(ctx, o);
  }
  API_END(ctx)
}

nosh_status nosh_keoreg_get_diags(nosh_ctx *ctx, double *d0, double *d1b) {
  API_BEGIN(ctx)
  if (!ctx->keoreg_ok) NOSH_THROW(NOSH_ESTATE, "regularised KEO not built");
  d2h(ctx, (double2 *)d0, ctx->pd0.p, ctx->No);
  d2h(ctx, d1b, ctx->pd1.p, ctx->No);
  API_END(ctx)
}

nosh_status nosh_keoreg_apply(nosh_ctx *ctx, const double *X, int64_t ldx, double *Y, int64_t ldy, int nvec,
                              nosh_transp mode, double alpha, double beta) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  // src/keo_regularized.cpp:98-100 (TEUCHOS_ASSERT_EQUALITY on mode, alpha, beta)
  if (mode != NOSH_NO_TRANS) NOSH_THROW(NOSH_EINVAL, "Only untransposed applies supported.");
  if (alpha != 1.0) NOSH_THROW(NOSH_EINVAL, "Only alpha==1.0 supported.");
  if (beta != 0.0) NOSH_THROW(NOSH_EINVAL, "Only beta==0.0 supported.");
  check_apply_shape(ctx, ldx, ldy, nvec);
  ensure_work(ctx);
  amg_ensure(ctx);
  for (int v = 0; v < nvec; v++) {
    double2 *x = stage_in(ctx, X + (size_t)v * ldx, ctx->stage_x, false);
    OutVec o = stage_out(ctx, Y + (size_t)v * ldy, ctx->stage_y);
    double2 *out = o.dev;
    if (!o.host && ctx->nranks > 1) out = ctx->work[11].p;  // the V-cycle needs ghost room
    amg_vcycle(ctx, x, out, nullptr);
    if (out != o.dev && ctx->No)
      CUDA_CHECK(cudaMemcpyAsync(o.dev, out, sizeof(double2) * ctx->No, cudaMemcpyDeviceToDevice, ctx->stream));
    finish_out(ctx, o);
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_amg_set_options(nosh_ctx *ctx, int degree, int coarse_degree, int coarse_max, int max_levels,
                                 int reuse) {
  API_BEGIN(ctx)
  if (coarse_max > 4096) NOSH_THROW(NOSH_EINVAL, "coarse_max > 4096");
  if (max_levels > NOSH_AMG_MAX_LEVELS) NOSH_THROW(NOSH_EINVAL, "max_levels > %d", NOSH_AMG_MAX_LEVELS);
  if (reuse > 1) NOSH_THROW(NOSH_EINVAL, "unknown reuse policy %d", reuse);
  if (degree > 0) ctx->amg_degree = degree;
  if (coarse_degree > 0) ctx->amg_coarse_degree = coarse_degree;
  if (coarse_max > 0) ctx->amg_coarse_max = coarse_max;
  if (max_levels > 0) ctx->amg_max_levels = max_levels;
  if (reuse >= 0) ctx->amg_reuse = reuse;
  amg_free(ctx);
  API_END(ctx)
}

nosh_status nosh_amg_setup(nosh_ctx *ctx) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ensure_work(ctx);
  ctx->amg_valid = false;
  amg_ensure(ctx);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_amg_info(nosh_ctx *ctx, nosh_amg_info_t *info) {
  API_BEGIN(ctx)
  if (!info) NOSH_THROW(NOSH_EINVAL, "NULL info");
  if (!ctx->amg) NOSH_THROW(NOSH_ESTATE, "no AMG hierarchy built");
  memset(info, 0, sizeof(*info));
  info->levels = (int32_t)ctx->amg->levels.size();
  info->degree = ctx->amg_degree;
  info->coarse_degree = ctx->amg_coarse_degree;
  info->setup_seconds = ctx->amg->setup_seconds;
  for (int l = 0; l < info->levels && l < NOSH_AMG_MAX_LEVELS; l++) {
    const AmgLevel &L = *ctx->amg->levels[l];
    info->nodes[l] = L.n;
    info->blocks[l] = L.nb;
    info->p_blocks[l] = L.p_nnz;
    info->lambda_max[l] = L.lam;
  }
  API_END(ctx)
}

static AmgLevel &amg_level(Ctx *ctx, int level) {
  if (!ctx->amg) NOSH_THROW(NOSH_ESTATE, "no AMG hierarchy built");
  if (level < 0 || level >= (int)ctx->amg->levels.size()) NOSH_THROW(NOSH_EINVAL, "no AMG level %d", level);
  return *ctx->amg->levels[level];
}
static void export_csr(Ctx *ctx, int64_t nrows, int64_t nnz, const int32_t *rowptr_dev, const int32_t *col_dev,
                       const B22 *val_dev, int64_t *rowptr, int32_t *cols, double *vals) {
  if (rowptr) {
    DBuf<int64_t> w;
    w.alloc(nrows + 1);
    k_widen<<<(unsigned)cdiv(nrows + 1, 256), 256, 0, ctx->stream>>>(rowptr_dev, nrows + 1, w.p);
    ctx->launches++;
    d2h(ctx, rowptr, w.p, nrows + 1);
  }
  d2h(ctx, cols, col_dev, nnz);
  d2h(ctx, (B22 *)vals, val_dev, nnz);
}

nosh_status nosh_amg_get_aggregates(nosh_ctx *ctx, int level, int32_t *agg) {
  API_BEGIN(ctx)
  AmgLevel &L = amg_level(ctx, level);
  if (!L.agg.p) NOSH_THROW(NOSH_EINVAL, "level %d is the coarsest level", level);
  d2h(ctx, agg, L.agg.p, L.n);
  API_END(ctx)
}

nosh_status nosh_amg_get_matrix(nosh_ctx *ctx, int level, int64_t *rowptr, int32_t *cols, double *vals) {
  API_BEGIN(ctx)
  AmgLevel &L = amg_level(ctx, level);
  if (level == 0) NOSH_THROW(NOSH_EINVAL, "level 0 is the regularised KEO itself (nosh_get_block_csr + nosh_keoreg_get_diags)");
  export_csr(ctx, L.n, L.nb, L.rowptr.p, L.col.p, L.val.p, rowptr, cols, vals);
  API_END(ctx)
}

nosh_status nosh_amg_get_prolongator(nosh_ctx *ctx, int level, int64_t *rowptr, int32_t *cols, double *vals) {
  API_BEGIN(ctx)
  AmgLevel &L = amg_level(ctx, level);
  if (!L.p_rowptr.p) NOSH_THROW(NOSH_EINVAL, "level %d is the coarsest level", level);
  export_csr(ctx, L.n, L.p_nnz, L.p_rowptr.p, L.p_col.p, L.p_val.p, rowptr, cols, vals);
  API_END(ctx)
}

nosh_status nosh_ctx_set_linear_solver(nosh_ctx *ctx, nosh_linear_solver solver, int gmres_restart) {
  API_BEGIN(ctx)
  if (solver != NOSH_SOLVER_MINRES && solver != NOSH_SOLVER_CG && solver != NOSH_SOLVER_GMRES)
    NOSH_THROW(NOSH_EINVAL, "unknown linear solver %d", (int)solver);
  if (gmres_restart > 500) NOSH_THROW(NOSH_EINVAL, "restart length must be in [1, 500]");
  ctx->lin_solver = solver;
  if (gmres_restart > 0) ctx->gmres_restart = gmres_restart;
  API_END(ctx)
}

nosh_status nosh_ctx_set_preconditioner(nosh_ctx *ctx, nosh_precond prec) {
  API_BEGIN(ctx)
  if (prec != NOSH_PREC_NONE && prec != NOSH_PREC_KEOREG_AMG) NOSH_THROW(NOSH_EINVAL, "unknown preconditioner %d", (int)prec);
  ctx->precond = prec;
  API_END(ctx)
}

nosh_status nosh_dot(nosh_ctx *ctx, const double *x, const double *y, double *result) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!result) NOSH_THROW(NOSH_EINVAL, "NULL result");
  const double2 *xd = stage_in(ctx, x, ctx->stage_x, false);
  const double2 *yd = (y == x) ? xd : stage_in(ctx, y, ctx->stage_y, false);
  *result = dot_dev(ctx, xd, yd);
  API_END(ctx)
}

nosh_status nosh_norm2(nosh_ctx *ctx, const double *x, double *result) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!result) NOSH_THROW(NOSH_EINVAL, "NULL result");
  const double2 *xd = stage_in(ctx, x, ctx->stage_x, false);
  *result = sqrt(dot_dev(ctx, xd, xd));
  API_END(ctx)
}

static nosh_status krylov_entry(nosh_ctx *ctx, bool is_minres, nosh_operator_id op, int prec, const double *b,
                                double *x, double tol, int maxit, nosh_krylov_result *res, double *hist) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ensure_work(ctx);
  const double2 *bd = stage_in(ctx, b, ctx->stage_x, false);
  OutVec o = stage_out(ctx, x, ctx->stage_y);
  if (is_minres)
    minres_dev(ctx, op, prec, bd, 1.0, o.dev, tol, maxit, res, hist);
  else
    cg_dev(ctx, op, prec, bd, 1.0, o.dev, tol, maxit, res, hist);
  finish_out(ctx, o);
  API_END(ctx)
}

nosh_status nosh_minres(nosh_ctx *ctx, nosh_operator_id op, const double *b, double *x, double tol, int maxit,
                        nosh_krylov_result *res, double *hist) {
  return krylov_entry(ctx, true, op, NOSH_PREC_NONE, b, x, tol, maxit, res, hist);
}

nosh_status nosh_cg(nosh_ctx *ctx, nosh_operator_id op, const double *b, double *x, double tol, int maxit,
                    nosh_krylov_result *res, double *hist) {
  return krylov_entry(ctx, false, op, NOSH_PREC_NONE, b, x, tol, maxit, res, hist);
}

nosh_status nosh_minres_prec(nosh_ctx *ctx, nosh_operator_id op, nosh_precond prec, const double *b, double *x,
                             double tol, int maxit, nosh_krylov_result *res, double *hist) {
  return krylov_entry(ctx, true, op, (int)prec, b, x, tol, maxit, res, hist);
}

nosh_status nosh_cg_prec(nosh_ctx *ctx, nosh_operator_id op, nosh_precond prec, const double *b, double *x,
                         double tol, int maxit, nosh_krylov_result *res, double *hist) {
  return krylov_entry(ctx, false, op, (int)prec, b, x, tol, maxit, res, hist);
}

nosh_status nosh_gmres(nosh_ctx *ctx, nosh_operator_id op, nosh_precond prec, const double *b, double *x, double tol,
                       int maxit, int restart, nosh_krylov_result *res, double *hist) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ensure_work(ctx);
  const double2 *bd = stage_in(ctx, b, ctx->stage_x, false);
  OutVec o = stage_out(ctx, x, ctx->stage_y);
  gmres_dev(ctx, op, (int)prec, bd, 1.0, o.dev, tol, maxit, restart, res, hist);
  finish_out(ctx, o);
  API_END(ctx)
}

nosh_status nosh_newton(nosh_ctx *ctx, int np, const char *const *names, const double *values, double *psi,
                        double nl_tol, int nl_maxit, double lin_tol, int lin_maxit, nosh_newton_result *res,
                        int32_t *lin_iters, double *fnorms) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ensure_work(ctx);
  if (!psi) NOSH_THROW(NOSH_EINVAL, "NULL psi");
  if (nl_maxit < 0) NOSH_THROW(NOSH_EINVAL, "nl_maxit < 0");
  const bool dev = is_device_ptr(psi);
  double2 *x = ctx->work[8].p;  // Nl entries: the SpMV needs ghost room
  CUDA_CHECK(cudaMemcpyAsync(x, psi, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             ctx->stream));
  newton_dev(ctx, np, names, values, x, nl_tol, nl_maxit, lin_tol, lin_maxit, res, lin_iters, fnorms);
  CUDA_CHECK(cudaMemcpyAsync(psi, x, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_inner_product(nosh_ctx *ctx, const double *phi, const double *psi, double *result) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!result) NOSH_THROW(NOSH_EINVAL, "NULL result");
  const double2 *a = stage_in(ctx, phi, ctx->stage_x, false);
  const double2 *b = (psi == phi) ? a : stage_in(ctx, psi, ctx->stage_y, false);
  const double vol = weighted_sum_dev(ctx, 0, nullptr, nullptr);
  *result = weighted_sum_dev(ctx, 1, a, b) / vol;
  API_END(ctx)
}

nosh_status nosh_gibbs_energy(nosh_ctx *ctx, const double *psi, double *result) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!result) NOSH_THROW(NOSH_EINVAL, "NULL result");
  const double2 *a = stage_in(ctx, psi, ctx->stage_x, false);
  const double vol = weighted_sum_dev(ctx, 0, nullptr, nullptr);
  *result = weighted_sum_dev(ctx, 2, a, a) / vol;
  API_END(ctx)
}

nosh_status nosh_continuation(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                              const char *pname, double dp, int nsteps, double *psi, double nl_tol, int nl_maxit,
                              double lin_tol, int lin_maxit, nosh_continuation_step *steps) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ensure_work(ctx);
  if (!psi || !pname) NOSH_THROW(NOSH_EINVAL, "NULL argument");
  if (nsteps < 0 || nl_maxit < 0) NOSH_THROW(NOSH_EINVAL, "negative step count");
  const bool dev = is_device_ptr(psi);
  double2 *x = ctx->work[8].p;
  CUDA_CHECK(cudaMemcpyAsync(x, psi, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             ctx->stream));
  continuation_dev(ctx, np, names, values, pname, dp, nsteps, x, nl_tol, nl_maxit, lin_tol, lin_maxit, steps);
  CUDA_CHECK(cudaMemcpyAsync(psi, x, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_continuation_arclength(nosh_ctx *ctx, int np, const char *const *names, const double *values,
                                        const char *pname, const nosh_arclength_options *opt, double *psi,
                                        nosh_arclength_step *steps, int32_t *n_records) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  ensure_work(ctx);
  if (!psi || !pname || !opt) NOSH_THROW(NOSH_EINVAL, "NULL argument");
  const bool dev = is_device_ptr(psi);
  double2 *x = ctx->work[8].p;
  CUDA_CHECK(cudaMemcpyAsync(x, psi, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             ctx->stream));
  int n = 0;
  arclength_dev(ctx, np, names, values, pname, opt, x, steps, &n);
  if (n_records) *n_records = n;
  CUDA_CHECK(cudaMemcpyAsync(psi, x, sizeof(double2) * ctx->No, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END(ctx)
}

nosh_status nosh_scratch_vector(nosh_ctx *ctx, int slot, double **dev_ptr) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (slot < 0 || slot >= 8 || !dev_ptr) NOSH_THROW(NOSH_EINVAL, "bad scratch slot");
  const size_t n = ctx->Nl > 0 ? ctx->Nl : 1;
  if (ctx->scratch[slot].n < n) {
    ctx->scratch[slot].alloc(n);
    CUDA_CHECK(cudaMemsetAsync(ctx->scratch[slot].p, 0, sizeof(double2) * n, ctx->stream));
  }
  *dev_ptr = (double *)ctx->scratch[slot].p;
  API_END(ctx)
}

// ---- generic FVM cores (fvm.cu) ----------------------------------------------------------------------
namespace {
// real vectors of n_owned doubles: host pointers are staged through the ctx's (complex-sized) staging buffers
const double *real_in(Ctx *ctx, const double *p, DBuf<double2> &buf) {
  if (!p) NOSH_THROW(NOSH_EINVAL, "NULL vector");
  if (is_device_ptr(p)) return p;
  buf.ensure(ctx->Nl > 0 ? ctx->Nl : 1);
  CUDA_CHECK(cudaMemcpyAsync(buf.p, p, sizeof(double) * ctx->No, cudaMemcpyHostToDevice, ctx->stream));
  return (const double *)buf.p;
}
struct RealOut {
  double *dev, *user;
  bool host;
};
RealOut real_out(Ctx *ctx, double *p, DBuf<double2> &buf) {
  if (!p) NOSH_THROW(NOSH_EINVAL, "NULL vector");
  RealOut o{p, p, !is_device_ptr(p)};
  if (o.host) {
    buf.ensure(ctx->Nl > 0 ? ctx->Nl : 1);
    o.dev = (double *)buf.p;
  }
  return o;
}
void real_finish(Ctx *ctx, const RealOut &o) {
  if (o.host) {
    CUDA_CHECK(cudaMemcpyAsync(o.user, o.dev, sizeof(double) * ctx->No, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}
}  // namespace

nosh_status nosh_mesh_boundary_vertices(nosh_ctx *ctx, int32_t *flags) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (!flags) NOSH_THROW(NOSH_EINVAL, "NULL flags");
  DBuf<int32_t> f;
  f.alloc(ctx->No);
  fvm_boundary_vertices(ctx, f.p);
  d2h(ctx, flags, f.p, ctx->No);
  API_END(ctx)
}

nosh_status nosh_fvm_matrix_fill(nosh_ctx *ctx, const double *edge_coeff, const double *edge_lhs, const double *edge_rhs,
                                 const double *vertex_lhs, const double *vertex_rhs, const int32_t *dirichlet_mask,
                                 const double *dirichlet_values, double *rhs) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  fvm_matrix_fill(ctx, edge_coeff, edge_lhs, edge_rhs, vertex_lhs, vertex_rhs, dirichlet_mask, dirichlet_values);
  if (rhs && ctx->No > 0) {
    CUDA_CHECK(cudaMemcpyAsync(rhs, ctx->fvm_rhs.p, sizeof(double) * ctx->No,
                               is_device_ptr(rhs) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  API_END(ctx)
}

nosh_status nosh_fvm_matrix_apply(nosh_ctx *ctx, const double *x, double *y) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  const double *xd = real_in(ctx, x, ctx->stage_x);
  RealOut o = real_out(ctx, y, ctx->stage_y);
  fvm_apply_dev(ctx, true, NOSH_FVM_VERTEX_NONE, 0.0, nullptr, nullptr, NOSH_FVM_DIRICHLET_NONE, nullptr, xd, o.dev);
  real_finish(ctx, o);
  API_END(ctx)
}

nosh_status nosh_fvm_get_csr(nosh_ctx *ctx, int64_t *rowptr, int32_t *cols, double *vals) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (vals && !ctx->fvm_filled) NOSH_THROW(NOSH_ESTATE, "FVM matrix not filled");
  if (rowptr) {
    DBuf<int64_t> w;
    w.alloc(ctx->No + 1);
    k_widen<<<(unsigned)cdiv(ctx->No + 1, 256), 256, 0, ctx->stream>>>(ctx->rowptr.p, ctx->No + 1, w.p);
    ctx->launches++;
    d2h(ctx, rowptr, w.p, ctx->No + 1);
  }
  if (cols) {  // local column ids -> global (one rank: identical)
    d2h(ctx, cols, ctx->csr_col.p, ctx->nb);
  }
  if (vals) {
    DBuf<double> g;
    g.alloc(ctx->nb);
    k_gather_real<<<(unsigned)cdiv(ctx->nb, 256), 256, 0, ctx->stream>>>(ctx->csr_pos.p, ctx->fvm_val.p, ctx->nb, g.p);
    ctx->launches++;
    d2h(ctx, vals, g.p, ctx->nb);
  }
  API_END(ctx)
}

nosh_status nosh_fvm_operator_apply(nosh_ctx *ctx, int with_matrix, nosh_fvm_vertex_core vertex_core, double alpha,
                                    const double *u0, const int32_t *dirichlet_mask, nosh_fvm_dirichlet_kind dirichlet_kind,
                                    const double *dirichlet_values, const double *x, double *y) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (vertex_core == NOSH_FVM_VERTEX_EXP_LINEARIZED && !u0) NOSH_THROW(NOSH_EINVAL, "the linearised core needs u0");
  if (dirichlet_kind != NOSH_FVM_DIRICHLET_NONE && !dirichlet_mask) NOSH_THROW(NOSH_EINVAL, "Dirichlet kind without a mask");
  if (dirichlet_kind == NOSH_FVM_DIRICHLET_VALUE && !dirichlet_values) NOSH_THROW(NOSH_EINVAL, "Dirichlet values missing");
  const int64_t No = ctx->No > 0 ? ctx->No : 1;
  DBuf<int32_t> dm;
  DBuf<double> dv, du0;
  const double *u0d = nullptr;
  if (u0) {
    if (is_device_ptr(u0)) {
      u0d = u0;
    } else {
      du0.alloc(No);
      CUDA_CHECK(cudaMemcpyAsync(du0.p, u0, sizeof(double) * ctx->No, cudaMemcpyHostToDevice, ctx->stream));
      u0d = du0.p;
    }
  }
  if (dirichlet_mask) {
    dm.alloc(No);
    CUDA_CHECK(cudaMemcpyAsync(dm.p, dirichlet_mask, sizeof(int32_t) * ctx->No, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (dirichlet_values) {
    dv.alloc(No);
    CUDA_CHECK(cudaMemcpyAsync(dv.p, dirichlet_values, sizeof(double) * ctx->No, cudaMemcpyHostToDevice, ctx->stream));
  }
  const double *xd = real_in(ctx, x, ctx->stage_x);
  RealOut o = real_out(ctx, y, ctx->stage_y);
  fvm_apply_dev(ctx, with_matrix != 0, vertex_core, alpha, u0d, dirichlet_mask ? dm.p : nullptr, dirichlet_kind,
                dirichlet_values ? dv.p : nullptr, xd, o.dev);
  real_finish(ctx, o);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // dm / dv / du0 go out of scope
  API_END(ctx)
}

nosh_status nosh_fvm_cg(nosh_ctx *ctx, const double *b, double *x, double tol, int maxit, nosh_krylov_result *res) {
  API_BEGIN(ctx)
  require_mesh(ctx);
  if (maxit < 0) NOSH_THROW(NOSH_EINVAL, "maxit < 0");
  const double *bd = real_in(ctx, b, ctx->stage_x);
  RealOut o = real_out(ctx, x, ctx->stage_y);
  fvm_cg_dev(ctx, bd, o.dev, tol, maxit, res);
  real_finish(ctx, o);
  API_END(ctx)
}

nosh_status nosh_ctx_set_step_observer(nosh_ctx *ctx, nosh_step_observer_fn fn, void *user) {
  API_BEGIN(ctx)
  ctx->step_observer = fn;
  ctx->step_observer_user = user;
  API_END(ctx)
}

nosh_status nosh_ctx_set_tuning(nosh_ctx *ctx, const char *key, int value) {
  API_BEGIN(ctx)
  if (!key) NOSH_THROW(NOSH_EINVAL, "NULL key");
  if (strcmp(key, "apply_variant") == 0) {
    if (value < 0 || value > 7) NOSH_THROW(NOSH_EINVAL, "apply_variant must be in [0, 7]");
    ctx->apply_variant = value;
  } else if (strcmp(key, "persistent_minres") == 0) {
    ctx->persistent_minres = value != 0;
  } else if (strcmp(key, "persistent_mgpu") == 0) {
    ctx->persistent_mgpu = value != 0;
  } else if (strcmp(key, "amg_panel_products") == 0) {
    if (value < 1) NOSH_THROW(NOSH_EINVAL, "amg_panel_products must be positive");
    ctx->amg_panel_products = value;
    ctx->amg_valid = false;
  } else if (strcmp(key, "amg_graph") == 0) {
    ctx->amg_graph = value != 0;
  } else if (strcmp(key, "amg_mixed") == 0) {
    ctx->amg_mixed = value != 0;
  } else if (strcmp(key, "mgpu_lean") == 0) {
    ctx->mgpu_lean = value != 0;
  } else if (strcmp(key, "mgpu_fence") == 0) {
    ctx->mgpu_fence = value;
  } else if (strcmp(key, "sell_sigma") == 0) {
    if (ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "sell_sigma must be chosen before the mesh is set");
    if (value < -1 || value > 1) NOSH_THROW(NOSH_EINVAL, "sell_sigma: -1 auto, 0 off, 1 on");
    ctx->sell_sigma = value;
  } else {
    NOSH_THROW(NOSH_EKEY, "unknown tuning key \"%s\"", key);
  }
  API_END(ctx)
}

int64_t nosh_launch_count(const nosh_ctx *ctx) { return ctx ? ctx->launches : 0; }

nosh_status nosh_timer_start(nosh_ctx *ctx) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
  API_END(ctx)
}

nosh_status nosh_timer_stop(nosh_ctx *ctx, float *ms) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
  if (ms) CUDA_CHECK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  API_END(ctx)
}

}  // extern "C"
