// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//   * halo exchange of a state vector's boundary entries before an operator apply
//     (replaces the Tpetra Import inside CrsMatrix::apply, call site
//     src/jacobian_operator.cpp:65)
//   * all-reduce of the fixed-order group sums (replaces Teuchos::reduceAll under
//     Tpetra::MultiVector::dot / norm2)
// NCCL is bound lazily with dlopen so that single-GPU use has no NCCL dependency and the
// library picks up whichever libnccl.so.2 the host process already loaded (torch's).
#include <dlfcn.h>

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
}

void comm_destroy(Ctx *ctx) {
  if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
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
}

// vec: Nl complex entries; fills the ghost segment [No, No+Ng) from the owners.
void halo_exchange(Ctx *ctx, double2 *vec) {
  const int P = ctx->nranks;
  if (P == 1 || (ctx->Ng == 0 && ctx->n_send == 0)) return;
  NcclApi *a = ctx->nccl;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  if (ctx->n_send) {
    k_pack<<<(unsigned)cdiv(ctx->n_send, 256), 256, 0, ctx->stream>>>(vec, ctx->send_idx.p, ctx->n_send,
                                                                       ctx->send_buf.p);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
  }
  NCCL_CHECK(a, a->GroupStart());
  for (int r = 0; r < P; r++) {
    if (r == ctx->rank) continue;
    if (ctx->send_count[r])
      NCCL_CHECK(a, a->Send(ctx->send_buf.p + ctx->send_off[r], (size_t)ctx->send_count[r] * 2, ncclFloat64, r,
                            comm, ctx->stream));
    if (ctx->recv_count[r])
      NCCL_CHECK(a, a->Recv(vec + ctx->No + ctx->recv_off[r], (size_t)ctx->recv_count[r] * 2, ncclFloat64, r, comm,
                            ctx->stream));
  }
  NCCL_CHECK(a, a->GroupEnd());
}

}  // namespace nosh
