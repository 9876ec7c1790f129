// comm.cu -- multi-GPU plumbing: one process per GPU, peer memory over NVLink/NVSwitch.
//   * halo exchange of a state vector's boundary entries before an operator apply
//     (replaces the Tpetra Import inside CrsMatrix::apply, call site
//     src/jacobian_operator.cpp:65)
//   * all-gather of the fixed-order group sums (replaces Teuchos::reduceAll under
//     Tpetra::MultiVector::dot / norm2)
// Data path: CUDA-IPC mapped peer buffers -- kernels store halo entries and group sums straight into
// the consumers' HBM and signal with epoch flags (k_halo_push / k_halo_wait here, k_finalize and
// k_minres_persistent_mgpu in krylov.cu).  Set-up (counts, index lists, IPC handles) travels either
// through the caller's own communicator (nosh_ctx_comm_init_host: a host all-gather callback, the
// image of the Teuchos::Comm the reference's mesh carries, src/mesh_reader.cpp:53-57) or through NCCL
// (nosh_ctx_comm_init).  NCCL is bound lazily with dlopen, is the data path only when IPC is
// unavailable, and is not touched at all in host-communicator mode or on one GPU.
#include <dlfcn.h>
#include <time.h>

#include <algorithm>

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

static void comm_common_init(Ctx *ctx, int rank, int nranks) {
  if (ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "comm_init must precede the mesh");
  if (nranks < 1 || rank < 0 || rank >= nranks) NOSH_THROW(NOSH_EINVAL, "bad rank/nranks");
  if (nranks > MAX_RANKS && nranks > 1) NOSH_THROW(NOSH_EINVAL, "at most %d ranks (one box)", MAX_RANKS);
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return;
  if (!ctx->stream2) {
    int lo = 0, hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
    for (cudaEvent_t *e : {&ctx->e_b, &ctx->e_halo, &ctx->e_finb, &ctx->e_c})
      CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
}

void comm_init(Ctx *ctx, const void *id128, int rank, int nranks) {
  comm_common_init(ctx, rank, nranks);
  if (nranks == 1) return;
  NcclApi *a = nccl_api();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c;
  NCCL_CHECK(a, a->CommInitRank(&c, nranks, id, rank));
  ctx->nccl = a;
  ctx->comm = c;
}

void comm_init_host(Ctx *ctx, int rank, int nranks, HostAllgather fn, void *user) {
  if (nranks > 1 && !fn) NOSH_THROW(NOSH_EINVAL, "NULL all-gather callback");
  comm_common_init(ctx, rank, nranks);
  ctx->host_ag = nranks > 1 ? fn : nullptr;
  ctx->host_ag_user = user;
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

// set-up only: all-gather of one fixed-size HOST record per rank
void exchange_allgather(Ctx *ctx, const void *send, void *recv, size_t bytes) {
  if (ctx->host_ag) {
    const int rc = ctx->host_ag(ctx->host_ag_user, send, recv, (int64_t)bytes);
    if (rc != 0) NOSH_THROW(NOSH_ECOMM, "host all-gather callback failed (%d)", rc);
    return;
  }
  NcclApi *a = ctx->nccl;
  if (!a || !ctx->comm) NOSH_THROW(NOSH_ECOMM, "no communicator (nosh_ctx_comm_init / nosh_ctx_comm_init_host)");
  DBuf<char> ds, da;
  ds.alloc(bytes);
  da.alloc(bytes * ctx->nranks);
  CUDA_CHECK(cudaMemcpyAsync(ds.p, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(a, a->AllGather(ds.p, da.p, bytes, 0 /* ncclInt8 */, (ncclComm_t)ctx->comm, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(recv, da.p, bytes * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// sum over ranks of n doubles (device pointers).  NCCL when the library owns a communicator; in
// host-communicator mode (GMRES' batched Gram-Schmidt sums, which synchronise with the host anyway)
// through the callback, added in rank order -- every entry is non-zero on one rank only, so exact.
void comm_allreduce_sum(Ctx *ctx, const double *send, double *recv, int64_t n) {
  if (ctx->nccl && ctx->comm) {
    NcclApi *a = ctx->nccl;
    NCCL_CHECK(a, a->AllReduce(send, recv, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return;
  }
  if (!ctx->host_ag) NOSH_THROW(NOSH_ECOMM, "no communicator");
  const int P = ctx->nranks;
  std::vector<double> mine((size_t)n), all((size_t)n * P);
  CUDA_CHECK(cudaMemcpyAsync(mine.data(), send, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  exchange_allgather(ctx, mine.data(), all.data(), sizeof(double) * n);
  for (int64_t i = 0; i < n; i++) {
    double s = 0.0;
    for (int r = 0; r < P; r++) s += all[(size_t)r * n + i];
    mine[i] = s;
  }
  CUDA_CHECK(cudaMemcpyAsync(recv, mine.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

namespace {
double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
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
}  // namespace

// Build the exchange plan.  Ghosts are sorted by global id and ownership ranges are
// contiguous, so the ghosts owned by one peer form one contiguous run of the ghost segment.
// Two all-gathers of host records (counts, then the padded "which of your vertices I need" lists).
void halo_setup(Ctx *ctx) {
  const double t0 = now_s();
  const int P = ctx->nranks, me = ctx->rank;
  ctx->send_count.assign(P, 0);
  ctx->recv_count.assign(P, 0);
  ctx->send_off.assign(P + 1, 0);
  ctx->recv_off.assign(P + 1, 0);
  ctx->n_send = 0;
  ctx->n_chunks_int = ctx->n_chunks_bnd = 0;
  if (P == 1) return;
  // ghost gids on host
  std::vector<int32_t> gg(ctx->Ng);
  if (ctx->Ng)
    CUDA_CHECK(cudaMemcpyAsync(gg.data(), ctx->gid.p + ctx->No, sizeof(int32_t) * ctx->Ng,
                               cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> want(ctx->Ng);  // owner-local ids, grouped by owner (gids are sorted)
  {
    int r = 0;
    for (int64_t i = 0; i < ctx->Ng; i++) {
      while (r + 1 < P && gg[i] >= ctx->part_begin[r + 1]) r++;
      ctx->recv_count[r]++;
      want[i] = (int32_t)(gg[i] - ctx->part_begin[r]);
    }
  }
  if (ctx->recv_count[me]) NOSH_THROW(NOSH_EMESH, "internal: a ghost vertex is owned by its own rank");
  for (int r = 0; r < P; r++) ctx->recv_off[r + 1] = ctx->recv_off[r] + ctx->recv_count[r];
  // count matrix: row r = what rank r receives from every peer
  std::vector<int64_t> all((size_t)P * P);
  exchange_allgather(ctx, ctx->recv_count.data(), all.data(), sizeof(int64_t) * P);
  for (int r = 0; r < P; r++) ctx->send_count[r] = all[(size_t)r * P + me];  // what r wants from me
  for (int r = 0; r < P; r++) ctx->send_off[r + 1] = ctx->send_off[r] + ctx->send_count[r];
  ctx->n_send = ctx->send_off[P];
  int64_t max_ng = 1;
  for (int r = 0; r < P; r++) {
    int64_t s = 0;
    for (int q = 0; q < P; q++) s += all[(size_t)r * P + q];
    max_ng = std::max(max_ng, s);
  }
  // every rank publishes its want list (padded to the longest); I pick the runs addressed to me
  want.resize(max_ng, 0);
  std::vector<int32_t> want_all((size_t)max_ng * P);
  exchange_allgather(ctx, want.data(), want_all.data(), sizeof(int32_t) * max_ng);
  std::vector<int32_t> sidx(ctx->n_send > 0 ? ctx->n_send : 1);
  for (int r = 0; r < P; r++) {
    if (r == me || ctx->send_count[r] == 0) continue;
    int64_t off = 0;
    for (int q = 0; q < me; q++) off += all[(size_t)r * P + q];
    const int32_t *src = want_all.data() + (size_t)r * max_ng + off;
    for (int64_t i = 0; i < ctx->send_count[r]; i++) {
      if (src[i] < 0 || src[i] >= ctx->No) NOSH_THROW(NOSH_ECOMM, "halo plan: rank %d asks for a vertex I do not own", r);
      sidx[ctx->send_off[r] + i] = src[i];
    }
  }
  ctx->send_idx.alloc(ctx->n_send);
  ctx->send_buf.alloc(ctx->n_send);
  if (ctx->n_send)
    CUDA_CHECK(cudaMemcpyAsync(ctx->send_idx.p, sidx.data(), sizeof(int32_t) * ctx->n_send, cudaMemcpyHostToDevice,
                               ctx->stream));
  // the same list sorted by the 512-row chunk of the source vertex (stable), for kernels that push a chunk's
  // boundary entries right after computing them (k_minres_persistent_mgpu, LEAN schedule)
  {
    const int64_t nch = cdiv(ctx->No, CHUNK);
    std::vector<int32_t> cptr(nch + 1, 0), csrc(ctx->n_send > 0 ? ctx->n_send : 1), crank(csrc.size()), coff(csrc.size());
    for (int64_t e = 0; e < ctx->n_send; e++) cptr[sidx[e] / CHUNK + 1]++;
    for (int64_t c = 0; c < nch; c++) cptr[c + 1] += cptr[c];
    std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
    for (int r = 0; r < P; r++)
      for (int64_t e = ctx->send_off[r]; e < ctx->send_off[r + 1]; e++) {
        const int32_t at = fill[sidx[e] / CHUNK]++;
        csrc[at] = sidx[e];
        crank[at] = r;
        coff[at] = (int32_t)(e - ctx->send_off[r]);
      }
    ctx->csend_ptr.alloc(nch + 1);
    ctx->csend_src.alloc(csrc.size());
    ctx->csend_rank.alloc(csrc.size());
    ctx->csend_off.alloc(csrc.size());
    CUDA_CHECK(cudaMemcpyAsync(ctx->csend_ptr.p, cptr.data(), sizeof(int32_t) * (nch + 1), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->csend_src.p, csrc.data(), sizeof(int32_t) * csrc.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->csend_rank.p, crank.data(), sizeof(int32_t) * csrc.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->csend_off.p, coff.data(), sizeof(int32_t) * csrc.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  // interior / boundary chunk lists for the overlapped apply
  const int64_t nch = cdiv(ctx->No, CHUNK);
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
    li.insert(li.end(), lb.begin(), lb.end());
    ctx->chunks_all.alloc(li.size());
    CUDA_CHECK(cudaMemcpyAsync(ctx->chunks_all.p, li.data(), sizeof(int32_t) * li.size(), cudaMemcpyHostToDevice,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // li / lb are about to go out of scope
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->stats["setup.halo_s"] = now_s() - t0;
}

// -------------------------------------------------------------------------------------------------
// Peer-memory path.  Every rank exports four allocations with CUDA IPC -- its two MINRES r-buffers, a
// small gather/flag block and the ghost landing buffer of the stand-alone halo exchange -- and maps
// its peers'.  A producing kernel can then store halo entries and group sums directly into the
// consumers' HBM over NVLink (no NCCL call, no extra launch latency in the Krylov loop).
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

// Stand-alone exchange, producer side: store my boundary entries into every neighbour's landing buffer
// (slot = epoch parity); the CTA that draws the last ticket raises my flag in every neighbour.
__global__ void __launch_bounds__(256) k_halo_push_signal(const double2 *vec, const int32_t *idx, int64_t n,
                                                          const HaloView h, unsigned long long epoch) {
  __shared__ int s_last;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t slot = (int64_t)(epoch & 1ull);
  if (i < n) {
    int r = 0;
    while (r + 1 < h.P && i >= h.send_off[r + 1]) r++;
    h.dst[r][slot * h.slot_stride[r] + (i - h.send_off[r])] = vec[idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(h.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < h.P && h.send_off[threadIdx.x + 1] > h.send_off[threadIdx.x])
      *((volatile unsigned long long *)h.flag_of[threadIdx.x]) = epoch;
    if (threadIdx.x == 0) *h.ticket = 0u;
  }
}

// Consumer side: wait until every neighbour's flag has reached `epoch`; with vec != NULL also copy the
// landed ghosts behind the owned entries (callers that need one contiguous vector).
__global__ void __launch_bounds__(256) k_halo_wait(const HaloView *h, unsigned long long epoch, const double2 *G,
                                                   int64_t ng, double2 *ghost_out) {
  halo_wait_cta(h, epoch);
  if (ghost_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < ng) ghost_out[i] = __ldcg(G + (int64_t)(epoch & 1ull) * ng + i);
  }
}
}  // namespace

void p2p_teardown(Ctx *ctx) {
  for (int k = 0; k < 4; k++)
    for (int r = 0; r < MAX_RANKS; r++)
      if (ctx->p2p.opened[k][r]) {
        cudaIpcCloseMemHandle(ctx->p2p.opened[k][r]);
        ctx->p2p.opened[k][r] = nullptr;
      }
  ctx->p2p.ok = false;
}

void p2p_setup(Ctx *ctx) {
  const double t0 = now_s();
  p2p_teardown(ctx);
  const int P = ctx->nranks, me = ctx->rank;
  if (P == 1 || P > MAX_RANKS) return;
  bool want = true;
  if (const char *e = getenv("NOSH_B200_P2P"))
    if (atoi(e) == 0) want = false;
  if (!want && !ctx->nccl) NOSH_THROW(NOSH_ECOMM, "NOSH_B200_P2P=0 needs the NCCL communicator (nosh_ctx_comm_init)");
  if (!want) return;
  P2P &pp = ctx->p2p;
  const size_t MG2 = 2 * MAX_GROUPS;
  const size_t nloc = MG2 + 2 * MAX_RANKS + 4;
  pp.local.alloc(nloc);
  CUDA_CHECK(cudaMemsetAsync(pp.local.p, 0, sizeof(unsigned long long) * nloc, ctx->stream));
  const int64_t ng = ctx->Ng > 0 ? ctx->Ng : 1;
  pp.ghost.alloc(2 * ng);
  CUDA_CHECK(cudaMemsetAsync(pp.ghost.p, 0, sizeof(double2) * 2 * ng, ctx->stream));
  pp.hepoch = 0;
  // my handles + where each owner's block starts in my vectors / my landing buffer
  struct Pack {
    cudaIpcMemHandle_t h[4];
    int64_t base[MAX_RANKS];
    int64_t goff[MAX_RANKS];
    int64_t ng;
    int64_t ok;
  } mine;
  memset(&mine, 0, sizeof(mine));
  void *ptrs[4] = {ctx->work[0].p, ctx->work[1].p, pp.local.p, pp.ghost.p};
  int ok = 1;
  for (int k = 0; k < 4; k++)
    if (cudaIpcGetMemHandle(&mine.h[k], ptrs[k]) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  for (int q = 0; q < P; q++) {
    mine.base[q] = ctx->No + ctx->recv_off[q];
    mine.goff[q] = ctx->recv_off[q];
  }
  mine.ng = ng;
  mine.ok = ok;
  std::vector<Pack> all(P);
  exchange_allgather(ctx, &mine, all.data(), sizeof(Pack));
  for (int r = 0; r < P; r++) ok = ok && all[r].ok;
  for (int r = 0; r < P && ok; r++) {
    void *m[4];
    for (int k = 0; k < 4; k++) {
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
    pp.view.flags[r] = (unsigned long long *)m[2] + MG2;
    pp.ghost_base[r] = all[r].base[me];
    pp.halo.dst[r] = (double2 *)m[3] + all[r].goff[me];
    pp.halo.slot_stride[r] = all[r].ng;
    pp.halo.flag_of[r] = (unsigned long long *)m[2] + MG2 + MAX_RANKS + 2 + me;
  }
  // all ranks must agree (a rank that failed falls back => everybody does)
  int64_t mine_ok = ok;
  std::vector<int64_t> oks(P);
  exchange_allgather(ctx, &mine_ok, oks.data(), sizeof(int64_t));
  for (int r = 0; r < P; r++) ok = ok && oks[r];
  if (!ok) {
    p2p_teardown(ctx);
    if (!ctx->nccl)
      NOSH_THROW(NOSH_ECOMM, "CUDA IPC peer mapping failed and there is no NCCL communicator to fall back to");
    return;
  }
  double timeout_s = 10.0;
  if (const char *e = getenv("NOSH_B200_P2P_TIMEOUT_S")) timeout_s = atof(e) > 0 ? atof(e) : timeout_s;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
  const long long ticks = (long long)(timeout_s * 1e3 * (khz > 0 ? khz : 2000000));
  pp.view.P = P;
  pp.view.me = me;
  pp.view.err = (int *)(pp.local.p + MG2 + MAX_RANKS);
  pp.view.epoch_ctr = pp.local.p + MG2 + MAX_RANKS + 1;
  pp.view.timeout = ticks;
  pp.halo.P = P;
  pp.halo.me = me;
  pp.halo.my_flags = pp.local.p + MG2 + MAX_RANKS + 2;
  pp.halo.ticket = (unsigned int *)(pp.local.p + MG2 + 2 * MAX_RANKS + 2);
  pp.halo.err = pp.view.err;
  pp.halo.timeout = ticks;
  for (int r = 0; r < P; r++) {
    pp.halo.send_off[r] = ctx->send_off[r];
    pp.halo.recv_cnt[r] = ctx->recv_count[r];
  }
  pp.halo.send_off[P] = ctx->send_off[P];
  pp.halo_dev.alloc(1);
  CUDA_CHECK(cudaMemcpyAsync(pp.halo_dev.p, &pp.halo, sizeof(HaloView), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  pp.ok = true;
  ctx->stats["setup.p2p_s"] = now_s() - t0;
}

// ---- stand-alone halo exchange -----------------------------------------------------------------------
// begin: my boundary entries of `vec` start travelling into the neighbours' landing buffers.
void halo_begin(Ctx *ctx, const double2 *vec, cudaStream_t stream) {
  P2P &pp = ctx->p2p;
  pp.hepoch++;
  if (ctx->n_send == 0) return;
  if (!stream) stream = ctx->stream;
  k_halo_push_signal<<<(unsigned)cdiv(ctx->n_send, 256), 256, 0, stream>>>(vec, ctx->send_idx.p, ctx->n_send, pp.halo,
                                                                          pp.hepoch);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}
// a new exchange whose push is done by the caller's own kernel (apply.cu: halo_push_cta)
unsigned long long halo_next_epoch(Ctx *ctx) { return ++ctx->p2p.hepoch; }
// where the ghosts of the exchange begun last land (Ng entries, ghost order)
const double2 *halo_slot(Ctx *ctx) {
  const int64_t ng = ctx->Ng > 0 ? ctx->Ng : 1;
  return ctx->p2p.ghost.p + (int64_t)(ctx->p2p.hepoch & 1ull) * ng;
}
// end: wait for the neighbours' entries of this exchange.  Returns the landing slot (Ng entries, ghost order);
// with ghost_out != NULL the ghosts are also copied there (vec + No of a contiguous local vector).
const double2 *halo_end(Ctx *ctx, double2 *ghost_out, cudaStream_t stream) {
  P2P &pp = ctx->p2p;
  const int64_t ng = ctx->Ng > 0 ? ctx->Ng : 1;
  const double2 *slot = pp.ghost.p + (int64_t)(pp.hepoch & 1ull) * ng;
  if (ctx->Ng == 0) return slot;
  if (!stream) stream = ctx->stream;
  const unsigned grid = ghost_out ? (unsigned)cdiv(ctx->Ng, 256) : 1u;
  k_halo_wait<<<grid, 256, 0, stream>>>(pp.halo_dev.p, pp.hepoch, pp.ghost.p, ctx->Ng, ghost_out);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
  return slot;
}

// vec: Nl complex entries; fills the ghost segment [No, No+Ng) from the owners.
void halo_exchange(Ctx *ctx, double2 *vec, cudaStream_t stream) {
  const int P = ctx->nranks;
  if (P == 1) return;
  if (!stream) stream = ctx->stream;
  if (ctx->p2p.ok) {
    halo_begin(ctx, vec, stream);
    halo_end(ctx, vec + ctx->No, stream);
    return;
  }
  if (ctx->Ng == 0 && ctx->n_send == 0) return;
  NcclApi *a = ctx->nccl;
  if (!a) NOSH_THROW(NOSH_ECOMM, "halo exchange: neither peer memory nor NCCL available");
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
