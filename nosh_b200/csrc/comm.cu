// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//   * halo exchange of a state vector's boundary entries before an operator apply
//     (replaces the Tpetra Import inside CrsMatrix::apply, call site
//     src/jacobian_operator.cpp:65)
//   * all-reduce of the fixed-order group sums (replaces Teuchos::reduceAll under
//     Tpetra::MultiVector::dot / norm2)
// NCCL is bound lazily with dlopen so that single-GPU use has no NCCL dependency and the
// library picks up whichever libnccl.so.2 the host process already loaded (torch's).
#include <dlfcn.h>

#include <cstdlib>

#include "comm.h"

namespace nosh {

// minimal NCCL ABI (stable since 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclInt32 = 2, ncclInt64 = 4 };
enum { ncclSum = 0 };

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *g_api = nullptr;

NcclApi *nccl_api() {
  if (g_api) return g_api;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) NOSH_THROW(NOSH_ECOMM, "cannot load libnccl.so.2: %s", dlerror());
  NcclApi *a = new NcclApi;
  a->h = h;
#define SYM(field, name)                                                      \
  *(void **)(&a->field) = dlsym(h, name);                                     \
  if (!a->field) NOSH_THROW(NOSH_ECOMM, "libnccl: symbol %s not found", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_api = a;
  return a;
}

#define NCCL_CHECK(api, expr)                                                                   \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != 0) NOSH_THROW(NOSH_ECOMM, "%s failed: %s", #expr, (api)->GetErrorString(_r));     \
  } while (0)

void comm_unique_id(void *id128) {
  NcclApi *a = nccl_api();
  ncclUniqueId id;
  NCCL_CHECK(a, a->GetUniqueId(&id));
  memcpy(id128, &id, 128);
}

void comm_init(Ctx *ctx, const void *id128, int rank, int nranks) {
  if (ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "comm_init must precede the mesh");
  if (nranks < 1 || rank < 0 || rank >= nranks) NOSH_THROW(NOSH_EINVAL, "bad rank/nranks");
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return;
  NcclApi *a = nccl_api();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c;
  NCCL_CHECK(a, a->CommInitRank(&c, nranks, id, rank));
  ctx->nccl = a;
  ctx->comm = c;
  int lo = 0, hi = 0;
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
  for (cudaEvent_t *e : {&ctx->e_b, &ctx->e_halo, &ctx->e_finb, &ctx->e_c})
    CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
}

void p2p_teardown(Ctx *ctx);

void comm_destroy(Ctx *ctx) {
  p2p_teardown(ctx);
  if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  for (cudaEvent_t e : {ctx->e_b, ctx->e_halo, ctx->e_finb, ctx->e_c})
    if (e) cudaEventDestroy(e);
  ctx->e_b = ctx->e_halo = ctx->e_finb = ctx->e_c = nullptr;
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  ctx->stream2 = nullptr;
}

void comm_allreduce_sum(Ctx *ctx, const double *send, double *recv, int64_t n) {
  NcclApi *a = ctx->nccl;
  NCCL_CHECK(a, a->AllReduce(send, recv, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
}

namespace {
__global__ void k_pack(const double2 *vec, const int32_t *idx, int64_t n, double2 *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = vec[idx[i]];
}
// one CTA per chunk of 512 rows: does any stored column of the chunk point into the ghost segment?
__global__ void k_chunk_has_ghost(const int32_t *rowptr, const int32_t *slice_off, const int32_t *col,
                                  int64_t No, int64_t nslices, int sell, int32_t *flag) {
  const int64_t c = blockIdx.x;
  int b, e;
  if (sell) {
    const int64_t s0 = c * (CHUNK / 32), s1 = min((int64_t)nslices, s0 + CHUNK / 32);
    b = slice_off[s0];
    e = slice_off[s1];
  } else {
    const int64_t r0 = c * CHUNK, r1 = min(No, r0 + CHUNK);
    b = rowptr[r0];
    e = rowptr[r1];
  }
  int any = 0;
  for (int p = b + threadIdx.x; p < e; p += blockDim.x) any |= col[p] >= No;
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flag[c] = any;
}
__global__ void k_ghost_to_owner_local(const int32_t *gid_ghost, int64_t n, int64_t owner_vb, int32_t *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)(gid_ghost[i] - owner_vb);
}
}  // namespace

// Build the exchange plan.  Ghosts are sorted by global id and ownership ranges are
// contiguous, so the ghosts owned by one peer form one contiguous run of the ghost segment.
void halo_setup(Ctx *ctx) {
  const int P = ctx->nranks;
  ctx->send_count.assign(P, 0);
  ctx->recv_count.assign(P, 0);
  ctx->send_off.assign(P + 1, 0);
  ctx->recv_off.assign(P + 1, 0);
  ctx->n_send = 0;
  if (P == 1) return;
  NcclApi *a = ctx->nccl;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  // ghost gids on host
  std::vector<int32_t> gg(ctx->Ng);
  if (ctx->Ng)
    CUDA_CHECK(cudaMemcpyAsync(gg.data(), ctx->gid.p + ctx->No, sizeof(int32_t) * ctx->Ng,
                               cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (int64_t i = 0; i < ctx->Ng; i++) {
    int r = 0;
    while (r + 1 < P && gg[i] >= ctx->part_begin[r + 1]) r++;
    ctx->recv_count[r]++;
  }
  for (int r = 0; r < P; r++) ctx->recv_off[r + 1] = ctx->recv_off[r] + ctx->recv_count[r];
  // exchange the count matrix: row r = what rank r receives from every peer
  DBuf<int64_t> dsend, dall;
  dsend.alloc(P);
  dall.alloc((size_t)P * P);
  CUDA_CHECK(cudaMemcpyAsync(dsend.p, ctx->recv_count.data(), sizeof(int64_t) * P, cudaMemcpyHostToDevice,
                             ctx->stream));
  NCCL_CHECK(a, a->AllGather(dsend.p, dall.p, (size_t)P, ncclInt64, comm, ctx->stream));
  std::vector<int64_t> all((size_t)P * P);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), dall.p, sizeof(int64_t) * P * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < P; r++) ctx->send_count[r] = all[(size_t)r * P + ctx->rank];  // what r wants from me
  for (int r = 0; r < P; r++) ctx->send_off[r + 1] = ctx->send_off[r] + ctx->send_count[r];
  ctx->n_send = ctx->send_off[P];
  // tell every owner which of its vertices I need (as owner-local ids)
  DBuf<int32_t> want;
  want.alloc(ctx->Ng);
  for (int r = 0; r < P; r++)
    if (ctx->recv_count[r]) {
      const int64_t n = ctx->recv_count[r];
      k_ghost_to_owner_local<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(
          ctx->gid.p + ctx->No + ctx->recv_off[r], n, ctx->part_begin[r], want.p + ctx->recv_off[r]);
      ctx->launches++;
    }
  ctx->send_idx.alloc(ctx->n_send);
  ctx->send_buf.alloc(ctx->n_send);
  NCCL_CHECK(a, a->GroupStart());
  for (int r = 0; r < P; r++) {
    if (r == ctx->rank) continue;
    if (ctx->recv_count[r])
      NCCL_CHECK(a, a->Send(want.p + ctx->recv_off[r], (size_t)ctx->recv_count[r], ncclInt32, r, comm, ctx->stream));
    if (ctx->send_count[r])
      NCCL_CHECK(a, a->Recv(ctx->send_idx.p + ctx->send_off[r], (size_t)ctx->send_count[r], ncclInt32, r, comm,
                            ctx->stream));
  }
  NCCL_CHECK(a, a->GroupEnd());
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  // interior / boundary chunk lists for the overlapped apply
  const int64_t nch = cdiv(ctx->No, CHUNK);
  ctx->n_chunks_int = ctx->n_chunks_bnd = 0;
  if (nch > 0) {
    DBuf<int32_t> flag;
    flag.alloc(nch);
    k_chunk_has_ghost<<<(unsigned)nch, 256, 0, ctx->stream>>>(ctx->rowptr.p, ctx->slice_off.p, ctx->col.p, ctx->No,
                                                               ctx->nslices, ctx->layout == NOSH_LAYOUT_SELL32,
                                                               flag.p);
    ctx->launches++;
    std::vector<int32_t> hf(nch), li, lb;
    CUDA_CHECK(cudaMemcpyAsync(hf.data(), flag.p, sizeof(int32_t) * nch, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int64_t c = 0; c < nch; c++) (hf[c] ? lb : li).push_back((int32_t)c);
    ctx->n_chunks_int = (int64_t)li.size();
    ctx->n_chunks_bnd = (int64_t)lb.size();
    ctx->chunks_int.alloc(li.size());
    ctx->chunks_bnd.alloc(lb.size());
    if (!li.empty())
      CUDA_CHECK(cudaMemcpyAsync(ctx->chunks_int.p, li.data(), sizeof(int32_t) * li.size(), cudaMemcpyHostToDevice,
                                 ctx->stream));
    if (!lb.empty())
      CUDA_CHECK(cudaMemcpyAsync(ctx->chunks_bnd.p, lb.data(), sizeof(int32_t) * lb.size(), cudaMemcpyHostToDevice,
                                 ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}

// vec: Nl complex entries; fills the ghost segment [No, No+Ng) from the owners.
void halo_exchange(Ctx *ctx, double2 *vec, cudaStream_t stream) {
  const int P = ctx->nranks;
  if (P == 1 || (ctx->Ng == 0 && ctx->n_send == 0)) return;
  if (!stream) stream = ctx->stream;
  NcclApi *a = ctx->nccl;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  if (ctx->n_send) {
    k_pack<<<(unsigned)cdiv(ctx->n_send, 256), 256, 0, stream>>>(vec, ctx->send_idx.p, ctx->n_send,
                                                                  ctx->send_buf.p);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
  }
  NCCL_CHECK(a, a->GroupStart());
  for (int r = 0; r < P; r++) {
    if (r == ctx->rank) continue;
    if (ctx->send_count[r])
      NCCL_CHECK(a, a->Send(ctx->send_buf.p + ctx->send_off[r], (size_t)ctx->send_count[r] * 2, ncclFloat64, r,
                            comm, stream));
    if (ctx->recv_count[r])
      NCCL_CHECK(a, a->Recv(vec + ctx->No + ctx->recv_off[r], (size_t)ctx->recv_count[r] * 2, ncclFloat64, r, comm,
                            stream));
  }
  NCCL_CHECK(a, a->GroupEnd());
}

// -------------------------------------------------------------------------------------------------
// Peer-memory path.  Every rank exports three allocations with CUDA IPC -- its two MINRES r-buffers
// and a small gather/flag block -- and maps its peers'.  A producing kernel can then store halo
// entries and group sums directly into the consumers' HBM over NVLink (no NCCL call, no extra
// launch latency in the Krylov loop); see k_halo_push below and k_finalize (krylov.cu).
// -------------------------------------------------------------------------------------------------
namespace {
struct PushArgs {
  double2 *dst[MAX_RANKS];
  int64_t off[MAX_RANKS + 1];
  int P;
};
__global__ void k_halo_push(const double2 *vec, const int32_t *idx, int64_t n, PushArgs a) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    int r = 0;
    while (r + 1 < a.P && i >= a.off[r + 1]) r++;
    a.dst[r][i - a.off[r]] = vec[idx[i]];  // NVLink store into rank r's ghost segment
  }
  __threadfence_system();
}
}  // namespace

void p2p_teardown(Ctx *ctx) {
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < MAX_RANKS; r++)
      if (ctx->p2p.opened[k][r]) {
        cudaIpcCloseMemHandle(ctx->p2p.opened[k][r]);
        ctx->p2p.opened[k][r] = nullptr;
      }
  ctx->p2p.ok = false;
}

void p2p_setup(Ctx *ctx) {
  p2p_teardown(ctx);
  const int P = ctx->nranks, me = ctx->rank;
  if (P == 1 || P > MAX_RANKS) return;
  if (const char *e = getenv("NOSH_B200_P2P"))
    if (atoi(e) == 0) return;
  NcclApi *a = ctx->nccl;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  P2P &pp = ctx->p2p;
  const size_t nloc = 2 * MAX_GROUPS + MAX_RANKS + 2;
  pp.local.alloc(nloc);
  CUDA_CHECK(cudaMemsetAsync(pp.local.p, 0, sizeof(unsigned long long) * nloc, ctx->stream));
  pp.epoch = 0;
  // my three handles + where each owner's block starts in my vectors
  struct Pack {
    cudaIpcMemHandle_t h[3];
    int64_t base[MAX_RANKS];
  } mine;
  memset(&mine, 0, sizeof(mine));
  void *ptrs[3] = {ctx->work[0].p, ctx->work[1].p, pp.local.p};
  int ok = 1;
  for (int k = 0; k < 3; k++)
    if (cudaIpcGetMemHandle(&mine.h[k], ptrs[k]) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  for (int q = 0; q < P; q++) mine.base[q] = ctx->No + ctx->recv_off[q];
  DBuf<char> dsend, dall;
  dsend.alloc(sizeof(Pack));
  dall.alloc(sizeof(Pack) * P);
  CUDA_CHECK(cudaMemcpyAsync(dsend.p, &mine, sizeof(Pack), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(a, a->AllGather(dsend.p, dall.p, sizeof(Pack), 0 /* ncclInt8 */, comm, ctx->stream));
  std::vector<Pack> all(P);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), dall.p, sizeof(Pack) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < P && ok; r++) {
    void *m[3];
    for (int k = 0; k < 3; k++) {
      if (r == me) {
        m[k] = ptrs[k];
      } else if (cudaIpcOpenMemHandle(&m[k], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      } else {
        pp.opened[k][r] = m[k];
      }
    }
    if (!ok) break;
    pp.R[0][r] = (double2 *)m[0];
    pp.R[1][r] = (double2 *)m[1];
    pp.view.red[r] = (double *)m[2];
    pp.view.flags[r] = (unsigned long long *)m[2] + 2 * MAX_GROUPS;
    pp.ghost_base[r] = all[r].base[me];
  }
  // all ranks must agree (a rank that failed falls back to NCCL => everybody does)
  DBuf<double> f1, f2;
  f1.alloc(1);
  f2.alloc(1);
  const double mine_ok = ok ? 0.0 : 1.0;
  CUDA_CHECK(cudaMemcpyAsync(f1.p, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  comm_allreduce_sum(ctx, f1.p, f2.p, 1);
  double bad = 0.0;
  CUDA_CHECK(cudaMemcpyAsync(&bad, f2.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (bad != 0.0) {
    p2p_teardown(ctx);
    return;
  }
  pp.view.P = P;
  pp.view.me = me;
  pp.view.epoch = 0;
  pp.view.err = (int *)(pp.local.p + 2 * MAX_GROUPS + MAX_RANKS);
  pp.ok = true;
}

// stores my boundary entries of `vec` (= my work[which_r]) into every neighbour's ghost segment
void p2p_halo_push(Ctx *ctx, int which_r, const double2 *vec) {
  if (!ctx->p2p.ok || ctx->n_send == 0) return;
  PushArgs A;
  A.P = ctx->nranks;
  for (int r = 0; r < ctx->nranks; r++) {
    A.dst[r] = ctx->p2p.R[which_r][r] + ctx->p2p.ghost_base[r];
    A.off[r] = ctx->send_off[r];
  }
  A.off[ctx->nranks] = ctx->send_off[ctx->nranks];
  k_halo_push<<<(unsigned)cdiv(ctx->n_send, 256), 256, 0, ctx->stream>>>(vec, ctx->send_idx.p, ctx->n_send, A);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}

int p2p_check_error(Ctx *ctx) {
  if (!ctx->p2p.ok) return 0;
  int e = 0;
  CUDA_CHECK(cudaMemcpyAsync(&e, ctx->p2p.view.err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return e;
}

}  // namespace nosh
