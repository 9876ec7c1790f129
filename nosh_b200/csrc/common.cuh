// common.cuh -- context, device buffers, deterministic reductions, load/store helpers.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/nosh_b200.h"

namespace nosh {

// ----------------------------------------------------------------------------
// fixed reduction geometry: every grid reduction is a
// three-level tree keyed by the GLOBAL vertex index, so that its result does not
// depend on how vertices are distributed over GPUs:
//   level 1: one CTA sums one chunk of CHUNK consecutive vertices (fixed order)
//   level 2: one warp sums the chunk partials of one group (group_vertices)
//   level 3: the group sums of the whole mesh are added in a fixed tree
// ----------------------------------------------------------------------------
constexpr int CHUNK = 512;       // vertices per level-1 partial
constexpr int TPB = 256;         // threads per CTA of the vector kernels (2 vertices/thread)
constexpr int MAX_GROUPS = 1024; // level-3 fan-in limit

struct Exception {
  nosh_status code;
  std::string msg;
};

#define NOSH_THROW(code, ...)                           \
  do {                                                  \
    char _b[512];                                       \
    snprintf(_b, sizeof(_b), __VA_ARGS__);              \
    throw ::nosh::Exception{code, std::string(_b)};     \
  } while (0)

#define CUDA_CHECK(expr)                                                                     \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      NOSH_THROW(NOSH_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                 __LINE__);                                                                  \
  } while (0)

// A cudaMalloc'ed block with a first-fit sub-allocator (blocks can be handed back).  Used by the AMG set-up
// (amg.cu), whose hundreds of allocations would otherwise each be a driver call: on this pool's VMs a cudaMalloc /
// cudaFree costs 0.3 ... 10 ms depending on the box, whatever the size.
struct Arena {
  char *base = nullptr;
  size_t cap = 0;
  struct Blk {
    size_t off, size;
    bool free;
  };
  std::vector<Blk> blocks;  // by offset
  size_t high = 0, used = 0;
  void init(size_t bytes) {
    if (cudaMalloc(&base, bytes) != cudaSuccess) {
      cudaGetLastError();
      base = nullptr;
      bytes = 0;
    }
    cap = bytes;
    blocks.assign(1, Blk{0, bytes, true});
  }
  void *alloc(size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    for (size_t i = 0; i < blocks.size(); i++)
      if (blocks[i].free && blocks[i].size >= bytes) {
        if (blocks[i].size > bytes) blocks.insert(blocks.begin() + i + 1, Blk{blocks[i].off + bytes, blocks[i].size - bytes, true});
        blocks[i].size = bytes;
        blocks[i].free = false;
        used += bytes;
        high = std::max(high, used);
        return base + blocks[i].off;
      }
    return nullptr;
  }
  bool owns(const void *p) const { return base && (const char *)p >= base && (const char *)p < base + cap; }
  void release(void *p) {
    const size_t off = (size_t)((char *)p - base);
    for (size_t i = 0; i < blocks.size(); i++)
      if (blocks[i].off == off && !blocks[i].free) {
        blocks[i].free = true;
        used -= blocks[i].size;
        if (i + 1 < blocks.size() && blocks[i + 1].free) {
          blocks[i].size += blocks[i + 1].size;
          blocks.erase(blocks.begin() + i + 1);
        }
        if (i > 0 && blocks[i - 1].free) {
          blocks[i - 1].size += blocks[i].size;
          blocks.erase(blocks.begin() + i);
        }
        return;
      }
  }
  void destroy() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
    blocks.clear();
  }
};
// While set (AMG set-up), DBuf allocations of this thread come from the arena; a DBuf remembers where its memory
// came from, so it may be released -- or moved to permanent storage with rehome() -- at any later time BEFORE the
// arena dies.
inline thread_local Arena *g_dbuf_arena = nullptr;
struct AllocStats {
  double malloc_s = 0.0, free_s = 0.0;
  int64_t malloc_n = 0, free_n = 0;
};
inline thread_local AllocStats g_alloc_stats;
inline double wall_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  Arena *owner = nullptr;  // memory from an arena (returned to it on release)
  bool borrowed = false;   // memory owned by someone else (rehome): never freed here
  DBuf() = default;
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p && !borrowed) {
      if (owner) {
        owner->release(p);
      } else {
        const double t = wall_now();
        cudaFree(p);
        g_alloc_stats.free_s += wall_now() - t;
        g_alloc_stats.free_n++;
      }
    }
    p = nullptr;
    n = 0;
    owner = nullptr;
    borrowed = false;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    if (g_dbuf_arena) {
      p = (T *)g_dbuf_arena->alloc(count * sizeof(T));
      if (p) owner = g_dbuf_arena;
    }
    if (!p) {
      const double t = wall_now();
      CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
      g_alloc_stats.malloc_s += wall_now() - t;
      g_alloc_stats.malloc_n++;
    }
    n = count;
  }
  void ensure(size_t count) {
    if (count > n) alloc(count);
  }
  void swap(DBuf &o) {
    std::swap(p, o.p);
    std::swap(n, o.n);
    std::swap(owner, o.owner);
    std::swap(borrowed, o.borrowed);
  }
  size_t bytes() const { return n * sizeof(T); }
  // copy the contents to dst (device memory someone else owns, at least bytes() long), give the old memory back,
  // point there from now on.  Stream-ordered: the old block may be reused by work enqueued later on `stream`.
  void rehome(void *dst, cudaStream_t stream) {
    if (!p) return;
    CUDA_CHECK(cudaMemcpyAsync(dst, p, bytes(), cudaMemcpyDeviceToDevice, stream));
    const size_t keep = n;
    release();
    p = (T *)dst;
    n = keep;
    borrowed = true;
  }
};

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- MINRES / CG device-resident scalar state -------------------------------
struct KrylovState {
  // recurrences (Belos::MinresIter names)
  double beta1, beta, oldBeta, alpha, dbar, epsln, oldeps, delta, gbar, gamma, cs, sn, phi, phibar;
  // coefficients consumed by the vector kernels
  double inv_beta;      // 1/beta_k                 (v_k = r_k * inv_beta)
  double f_r1;          // beta_k/beta_{k-1}        (0 on the first iteration)
  double f_r2;          // alpha_k/beta_k
  double inv_beta_prev; // 1/beta of the vector W is built from
  double c_w1, c_w2, inv_gamma, phi_w;
  // CG
  double rho, rho_old, pAp, cg_alpha, cg_beta, r0norm;
  double tol;
  double relres;
  int iter;
  int maxit;
  int done;      // 1 => all further kernels are no-ops
  int converged;
  int breakdown;
};

struct NcclApi;  // comm.cu
struct Amg;      // amg.h

constexpr int MAX_RANKS = 8;

// Peer-memory (NVLink/NVSwitch) view of the other ranks' buffers, opened through CUDA IPC:
// lets one kernel push halo entries / group sums straight into a peer's HBM and signal it.
struct P2PView {
  double *red[MAX_RANKS];               // rank r's gather buffer: 2 slots x MAX_GROUPS doubles
  unsigned long long *flags[MAX_RANKS]; // rank r's arrival flags: one per source rank
  int P, me;
  unsigned long long *epoch_ctr;        // local: number of reductions EXECUTED so far (slot = parity); bumped on
                                        // the device only, so launches skipped after convergence do not count
  int *err;                             // local: set on a spin time-out
  long long timeout;                    // spin time-out in clock64 ticks
};
// Stand-alone halo exchange through peer memory (operator applies outside the persistent Krylov loop):
// every rank owns a ghost landing buffer G (2 slots x Ng entries) that its neighbours store into.
struct HaloView {
  double2 *dst[MAX_RANKS];              // where my block starts in rank r's landing buffer (slot 0)
  int64_t slot_stride[MAX_RANKS];       // rank r's Ng (distance between its two slots)
  int64_t send_off[MAX_RANKS + 1];      // my send list, grouped by destination rank
  int64_t recv_cnt[MAX_RANKS];          // how many ghosts rank r sends me
  unsigned long long *flag_of[MAX_RANKS];  // rank r's halo flag for source `me`
  unsigned long long *my_flags;         // my halo flags, one per source rank
  unsigned int *ticket;                 // last-CTA ticket of the push kernel
  int *err;
  long long timeout;
  int P, me;
};
struct P2P {
  bool ok = false;
  DBuf<unsigned long long> local;       // [2*MAX_GROUPS sums | MAX_RANKS flags | err | epoch | MAX_RANKS halo flags | ticket]
  DBuf<double2> ghost;                  // 2 x Ng landing buffer of the stand-alone halo exchange
  DBuf<HaloView> halo_dev;              // device copy of `halo` (read by the boundary CTAs of the apply kernels)
  void *opened[4][MAX_RANKS] = {};
  P2PView view;
  HaloView halo;
  double2 *R[2][MAX_RANKS] = {};        // rank r's MINRES r-buffers (work[0], work[1])
  int64_t ghost_base[MAX_RANKS] = {};   // where my block starts inside rank r's vectors
  unsigned long long hepoch = 0;        // halo exchanges issued (all ranks call them collectively)
};

// Host-side all-gather of fixed-size records, supplied by the caller (MPI_Allgather on the communicator the
// reference's mesh carries, torch.distributed in the tests): recv = P consecutive records in rank order.
typedef int (*HostAllgather)(void *user, const void *send, void *recv, int64_t bytes_per_rank);

enum MvpKind { MVP_NONE = 0, MVP_EXPLICIT = 1, MVP_CONSTCURL = 2 };
enum PotKind { POT_NONE = 0, POT_CONSTANT = 1, POT_VALUES = 2 };

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // multi-GPU overlap: communication / side stream and the events that order it against `stream`
  cudaStream_t stream2 = nullptr;
  cudaEvent_t e_b = nullptr, e_halo = nullptr, e_finb = nullptr, e_c = nullptr;

  // comm
  int rank = 0, nranks = 1;
  NcclApi *nccl = nullptr;
  void *comm = nullptr;  // ncclComm_t
  int (*step_observer)(void *, int, double, double, double, const double *, int64_t) = nullptr;  // continuation drivers
  void *step_observer_user = nullptr;
  HostAllgather host_ag = nullptr;  // set-up exchange through the caller's communicator (no NCCL in the library)
  void *host_ag_user = nullptr;
  std::map<std::string, double> stats;  // set-up timings etc. (nosh_ctx_get_stat)

  // options
  int layout = NOSH_LAYOUT_SELL32;
  int persistent_minres = 1;        // single-GPU unpreconditioned MINRES as one cooperative launch (krylov.cu);
                                    // env NOSH_B200_PERSISTENT_MINRES=0 / tuning key "persistent_minres" turn it off
  int persistent_mgpu = 1;          // multi-GPU persistent loop over peer memory (env NOSH_B200_PERSISTENT_MGPU=0 /
                                    // tuning key "persistent_mgpu" select the multi-launch loop)
  int persist_grid_mgpu = 0;
  int mgpu_lean = 0;                // k_minres_persistent_mgpu: 1 = the two-grid-syncs-per-iteration schedule (krylov.cu;
                                    // bit-identical, measured SLOWER: 1377 vs 1441 it/s on 2 B200), 0 = six syncs
  int mgpu_fence = 2;               // k_minres_persistent_mgpu: who fences system-wide after the halo push (krylov.cu)
  int persist_grid = 0;             // co-resident CTAs of that kernel (occupancy x SMs), computed once
  int apply_variant = 0;            // measurement knob: which k_apply_sell variant the MINRES loop uses (apply.cu)
  int64_t group_vertices = 65536;

  // ---- mesh ----
  bool has_mesh = false;
  int dim = 0;
  int64_t n_global = 0, vb = 0, ve = 0;
  int64_t No = 0, Ng = 0, Nl = 0, nc = 0, E = 0;
  std::vector<int64_t> part_begin;  // nranks+1 global vertex offsets
  DBuf<int32_t> gid;                // Nl global ids (int32: n_global < 2^31)
  DBuf<double> coords;              // Nl x 3
  DBuf<int32_t> cells;              // nc x (dim+1) local ids
  DBuf<int32_t> edges;              // E x 2 local ids, gid(v0) < gid(v1)
  DBuf<double> elen, ecov;          // E
  DBuf<double> cv;                  // No
  // ---- block matrix structure (owned rows) ----
  int64_t nb = 0;                   // CSR blocks
  int64_t nstored = 0;              // stored blocks (incl. SELL padding)
  int64_t nslices = 0;
  DBuf<int32_t> rowptr;             // No+1 (CSR positions)
  DBuf<int32_t> csr_col;            // nb local col ids (CSR order)
  DBuf<int32_t> csr_pos;            // nb: CSR position -> storage position
  DBuf<int32_t> edge_of;            // nb: edge of a CSR position (-1: diagonal)
  DBuf<int32_t> col;                // nstored local col ids (storage order)
  DBuf<int32_t> slice_off;          // nslices+1 (SELL)
  DBuf<int32_t> sell_row, sell_pos; // SELL-32-sigma: row stored at a position / position of a row (mesh.cu)
  bool sell_permuted = false;
  int sell_sigma = -1;              // -1 auto (sort when the padding exceeds 5 %; always with several ranks), 0 off, 1 on
  DBuf<int32_t> slot_ij, slot_ji;   // E storage positions (-1: row not owned)
  DBuf<int32_t> diag_slot;          // No storage positions
  DBuf<double2> Kval, dKval;        // nstored
  DBuf<double> Kdiag;               // No: sum of alpha over incident edges
  // ---- generic FVM matrix (fvm.cu): one double per storage slot of the same graph ----
  DBuf<double> fvm_val, fvm_rhs, fvm_dval;
  DBuf<int32_t> fvm_mask;
  bool fvm_filled = false;
  // ---- fields ----
  DBuf<double> thick;               // Nl
  bool thick_set = false;
  int pot_kind = POT_NONE;
  double pot_c = 0.0;
  std::string pot_param;
  DBuf<double> pot_values;          // No (POT_VALUES)
  DBuf<double> Vcur;                // No: V for the current parameters
  DBuf<double> dvdp;                // No: dV/dp scratch
  int mvp_kind = MVP_NONE;
  DBuf<double> ecache;              // E (explicit) or 3E (constcurl)
  double cc_b[3] = {0, 0, 1}, cc_u[3] = {0, 0, 0};
  bool cc_has_u = false;
  DBuf<double> alpha;               // E
  bool alpha_ok = false;
  // fill caches (the parameter cache the reference meant to have)
  bool keo_filled = false, dkeo_filled = false;
  double keo_mu = 0, keo_theta = 0, dkeo_mu = 0, dkeo_theta = 0;
  std::string dkeo_name;
  // ---- Jacobian / preconditioner diagonals ----
  DBuf<double2> jd0;                // No (d0[2k], d0[2k+1])
  DBuf<double> jd1;                 // No
  bool jac_ok = false;
  DBuf<double2> pd0;
  DBuf<double> pd1;
  bool keoreg_ok = false;
  // ---- preconditioner: AMG V-cycle on the regularised KEO (amg.cu) ----
  Amg *amg = nullptr;
  bool amg_valid = false;     // hierarchy matches the reuse policy
  int amg_degree = 1;         // Chebyshev degree of the pre-/post-smoother on the finest level
  int amg_coarse_degree = 2;  // ... on the coarse levels (cheap there: measured -18 % MINRES iterations for +7 % V-cycle cost)
  int amg_coarse_max = 512;   // nodes at which the hierarchy stops and a dense inverse is used
  int amg_max_levels = 10;
  int amg_reuse = 1;          // nosh_amg_reuse: 0 none, 1 full ("reuse: type" = "full", keo_regularized.cpp:300)
  int64_t amg_panel_products = (int64_t)32 << 20;  // products per row panel of the set-up's sparse products (amg.cu)
  int amg_graph = 1;          // replay the V-cycle as a CUDA graph (amg.cu:amg_vcycle)
  // mixed-precision V-cycle (tuning key "amg_mixed", off by default): the two finest-level smoother passes read an
  // fp32 copy of K (12 B per block instead of 20), restriction / prolongation an fp32 copy of the finest P.  Vectors,
  // diagonals, accumulation and every coarser level stay fp64; the Krylov solver's own operator is never touched.
  int amg_mixed = 0;
  DBuf<float2> Kval32;
  int64_t kval_version = 0, kval32_version = -1;
  bool amg_keep_l0 = false;   // keep the level-0 block CSR copy (parity accessors)
  int64_t keoreg_version = 0, amg_dinv_version = -1;
  int lin_solver = 0;         // nosh_linear_solver of the Newton / continuation drivers (default MINRES)
  int gmres_restart = 300;    // Belos "Num Blocks" default
  int precond = 0;            // nosh_precond the Newton / continuation drivers use for their linear solves
  // ---- work vectors (2*Nl doubles each) ----
  DBuf<double2> work[14];
  DBuf<double2> scratch[8];
  DBuf<double2> stage_x, stage_y;   // host staging
  // pipelined host I/O (capi.cu: nosh_prefetch / nosh_ctx_set_async_output): copies on their own stream
  cudaStream_t copy_stream = nullptr;
  struct Prefetch {
    const double *host = nullptr;
    DBuf<double2> dev;
    cudaEvent_t ready = nullptr;
    bool valid = false;
  } prefetch[4];
  int prefetch_next = 0;
  cudaEvent_t ev_order = nullptr;   // orders the copy stream behind the compute stream
  int async_output = 0;
  DBuf<double2> stage_y2;           // second output staging buffer (async output alternates)
  cudaEvent_t out_done[2] = {nullptr, nullptr};
  int out_flip = 0;
  DBuf<double2> gmres_basis;        // (restart+1) x Nl Arnoldi vectors, allocated by the first nosh_gmres
  // ---- reductions ----
  DBuf<double> partials;            // 2 x n_chunks
  DBuf<double> group_sums;          // 2 x MAX_GROUPS (global group index)
  DBuf<KrylovState> kstate;
  DBuf<unsigned int> ticket;        // last-CTA ticket counter of the in-kernel reductions
  DBuf<double> hist;
  DBuf<double> scalar_out;          // misc device scalars
  int64_t n_chunks = 0;
  int64_t n_groups_global = 0, group_begin = 0, n_groups_local = 0;
  int chunks_per_group = 0;
  // ---- halo exchange ----
  std::vector<int64_t> send_count, recv_count, send_off, recv_off;  // per peer, in vertices
  DBuf<int32_t> send_idx;           // owned local ids to pack, grouped by peer
  DBuf<double2> send_buf;
  int64_t n_send = 0;
  P2P p2p;
  // chunks (512 rows) whose rows reference no ghost column can be applied before the halo lands
  DBuf<int32_t> chunks_int, chunks_bnd;
  DBuf<int32_t> csend_ptr, csend_src, csend_rank, csend_off;  // send list sorted by source chunk (comm.cu)
  DBuf<int32_t> chunks_all;         // interior chunks first, boundary chunks last (one launch, the last CTAs wait)
  int64_t n_chunks_int = 0, n_chunks_bnd = 0;
};

// ---- parameter map helpers ---------------------------------------------------
inline const double *find_param(int np, const char *const *names, const double *values,
                                const char *key) {
  for (int i = 0; i < np; i++)
    if (names[i] && strcmp(names[i], key) == 0) return &values[i];
  return nullptr;
}
inline double param_at(int np, const char *const *names, const double *values, const char *key) {
  const double *p = find_param(np, names, values, key);
  if (!p) NOSH_THROW(NOSH_EKEY, "parameter \"%s\" missing (std::map::at)", key);
  return *p;
}

#ifdef __CUDACC__
// ---- device helpers ------------------------------------------------------------
__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }

// streaming (read-once) 128-bit load: do not pollute L1
__device__ __forceinline__ double2 ld_stream2(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream_f2(const float2 *p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream_i32(const int *p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Producer side of the stand-alone halo exchange, called by ALL threads of the first `nblocks` CTAs of a grid:
// store my boundary entries of `vec` into the neighbours' landing buffers (slot = epoch parity) over NVLink;
// the CTA that draws the last ticket raises my flag in every neighbour.
__device__ __forceinline__ void halo_push_cta(const HaloView *h, const double2 *vec, const int32_t *idx, int64_t n,
                                              unsigned long long epoch, int nblocks) {
  __shared__ int s_last;
  const int64_t slot = (int64_t)(epoch & 1ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)nblocks * blockDim.x) {
    int r = 0;
    while (r + 1 < h->P && i >= h->send_off[r + 1]) r++;
    h->dst[r][slot * h->slot_stride[r] + (i - h->send_off[r])] = vec[idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(h->ticket, 1u);
    s_last = (t == (unsigned int)nblocks - 1u);
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < h->P && h->send_off[threadIdx.x + 1] > h->send_off[threadIdx.x])
      *((volatile unsigned long long *)h->flag_of[threadIdx.x]) = epoch;
    if (threadIdx.x == 0) *h->ticket = 0u;
  }
}

// Consumer side of the stand-alone halo exchange, called by ALL threads of a CTA: wait until every neighbour's
// flag has reached `epoch` (its entries of this exchange have landed in my buffer), then acquire.
__device__ __forceinline__ void halo_wait_cta(const HaloView *h, unsigned long long epoch) {
  if ((int)threadIdx.x < h->P && h->recv_cnt[threadIdx.x] > 0) {
    const volatile unsigned long long *f = (const volatile unsigned long long *)&h->my_flags[threadIdx.x];
    const long long t0 = clock64();
    while (*f < epoch) {
      if (clock64() - t0 > h->timeout) {
        *h->err = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

// Deterministic CTA sum (fixed tree: xor-shuffle inside each warp, then the warp
// sums in warp order).  Result valid in thread 0.  NW = warps per CTA.
template <int NW>
__device__ __forceinline__ double block_sum(double v, double *smem /* NW doubles */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NW; i++) s += smem[i];
  }
  __syncthreads();
  return s;
}
#endif

}  // namespace nosh
