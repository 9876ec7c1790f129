// amg.cu -- smoothed-aggregation AMG V-cycle for the regularised KEO, built and applied on the device.
//
// Reference: keo_regularized::apply is ONE V-cycle of a MueLu hierarchy built with
// "number of equations" = 2 and "reuse: type" = "full" (src/keo_regularized.cpp:88-165,290-336).
// MueLu is not in the reference tree (parity unpinned); what is implemented here is the published
// smoothed-aggregation method with MueLu's documented defaults and a deterministic aggregation,
// restated on the CPU in oracle/amg.py -- the two are compared by tests/test_gpu_amg.py.
//
//   level 0     : the ctx's complex SELL-32 KEO + per-vertex 2x2 blocks (pd0, pd1); smoother steps are
//                 epilogues of the fused apply kernel (apply.cu FUSE_RESID / FUSE_CHEB)
//   level >= 1  : real 2x2-block SELL-32 matrices, one thread per block row
//   aggregation : MIS-2 by hashed priority (rounds of neighbourhood maxima), integer only => identical
//                 to the oracle's sequential greedy sweep
//   P           : (I - 4/3 / lambda_max D^-1 A) P0,   A_c = P^T A P   by expand-sort-compress SpGEMM
//                 (stable radix sort => fixed summation order => run-to-run identical bits)
//   smoother    : Chebyshev of degree `amg_degree` (finest level, default 1) / `amg_coarse_degree` (coarse
//                 levels, default 2) on S^-1 A over [1/20, 1], S = absolute row sums (l1-Jacobi
//                 scaling: lambda_max(S^-1 A) <= 1 is a true bound row by row, so the V-cycle stays positive
//                 definite on any mesh); degree 1 = damped l1-Jacobi
//   coarsest    : dense inverse (host Cholesky at set-up), warp-per-row matvec
// No atomics on floating-point data anywhere; restriction uses an explicit transpose index.
#include "amg.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cub/cub.cuh>

#include "apply.cuh"

namespace nosh {

namespace {

constexpr int POWER_ITS = 10;
constexpr double SA_DAMPING = 4.0 / 3.0;
constexpr double CHEB_RATIO = 20.0;
constexpr int ST_OUT = 0, ST_UND = 1, ST_IN = 2;
constexpr int64_t SMALL_LEVEL = 65536;  // rows below which a level uses the lanes-per-row kernels

inline dim3 grid_for(int64_t n, int tpb = 256) { return dim3((unsigned)cdiv(n > 0 ? n : 1, tpb)); }

#define ALAUNCH(ctx, kernel, n, ...)                             \
  do {                                                           \
    kernel<<<grid_for(n), 256, 0, (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                           \
    CUDA_CHECK(cudaGetLastError());                              \
  } while (0)

// All memory of the set-up comes from ONE cudaMalloc'ed arena (common.cuh: Arena, first fit, blocks can be handed
// back): the temporaries below (TBuf) and, through g_dbuf_arena, every DBuf -- the hierarchy's own buffers included,
// which build_hierarchy moves into one exact-size allocation at the end.  The expand-sort-compress products
// allocate and free hundreds of buffers; cudaMalloc / cudaFree per buffer cost 0.3 ... 10 ms each depending on the
// box (cudaFree synchronises), and the stream-ordered pool (cudaMallocAsync) -- used in round 1 -- grows through the
// virtual-memory API, which on this pool's VMs takes 0.05 ... 1.7 s for 8 GB (profiles/r2_alloc_probe.jsonl).  All
// work is on one stream, so a block handed back by one temporary may be given to the next at once: the kernels that
// used it were enqueued earlier (the ordering cudaFreeAsync gives).  A request the arena cannot serve falls back to
// cudaMalloc / cudaFree.
thread_local Arena *g_arena = nullptr;
thread_local size_t g_fallback_bytes = 0;
template <typename T>
struct TBuf {
  T *p = nullptr;
  size_t n = 0;
  TBuf() = default;
  TBuf(const TBuf &) = delete;
  TBuf &operator=(const TBuf &) = delete;
  ~TBuf() { release(); }
  void release() {
    if (p) {
      if (g_arena && g_arena->owns(p)) g_arena->release(p);
      else cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    p = g_arena ? (T *)g_arena->alloc(count * sizeof(T)) : nullptr;
    if (!p) {
      CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
      g_fallback_bytes += count * sizeof(T);
    }
    n = count;
  }
  void ensure(size_t count) {
    if (count > n) alloc(count);
  }
  void swap(TBuf &o) {
    std::swap(p, o.p);
    std::swap(n, o.n);
  }
};

struct Temp {
  TBuf<char> buf;
  void *get(size_t bytes) {
    buf.ensure(bytes);
    return buf.p;
  }
};

template <typename T>
T fetch(Ctx *ctx, const T *dptr) {
  T h;
  CUDA_CHECK(cudaMemcpyAsync(&h, dptr, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return h;
}

void check_count(int64_t n, const char *what) {
  if (n >= (int64_t)2147483647) NOSH_THROW(NOSH_EINVAL, "AMG set-up: %s needs %lld items (>= 2^31)", what, (long long)n);
}

void sort_pairs_u64(Ctx *ctx, Temp &tmp, TBuf<uint64_t> &keys, TBuf<uint32_t> &vals, int64_t n) {
  check_count(n, "sort");
  if (n == 0) return;
  TBuf<uint64_t> k2;
  TBuf<uint32_t> v2;
  k2.alloc(n);
  v2.alloc(n);
  cub::DoubleBuffer<uint64_t> dk(keys.p, k2.p);
  cub::DoubleBuffer<uint32_t> dv(vals.p, v2.p);
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n, 0, 64, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(t, bytes, dk, dv, (int)n, 0, 64, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (dk.Current() != keys.p) keys.swap(k2);
  if (dv.Current() != vals.p) vals.swap(v2);
}

void exclusive_scan_i32(Ctx *ctx, Temp &tmp, const int32_t *in, int32_t *out, int64_t n) {
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, (int)n, ctx->stream));
}
void exclusive_scan_i64(Ctx *ctx, Temp &tmp, const int64_t *in, int64_t *out, int64_t n) {
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, (int)n, ctx->stream));
}

// ---- 2x2 block arithmetic ---------------------------------------------------------------------
__host__ __device__ __forceinline__ B22 b22(double a, double b, double c, double d) {
  B22 r;
  r.a = a; r.b = b; r.c = c; r.d = d;
  return r;
}
__device__ __forceinline__ B22 mul(const B22 &x, const B22 &y) {
  return b22(x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d);
}
__device__ __forceinline__ B22 tmul(const B22 &x, const B22 &y) {  // x^T y
  return b22(x.a * y.a + x.c * y.c, x.a * y.b + x.c * y.d, x.b * y.a + x.d * y.c, x.b * y.b + x.d * y.d);
}
__device__ __forceinline__ void acc(B22 &s, const B22 &v) {
  s.a += v.a; s.b += v.b; s.c += v.c; s.d += v.d;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// 30-bit hashed priority of node v on `level` (oracle/amg.py:priority)
__device__ __forceinline__ uint64_t amg_priority(uint32_t v, int level) {
  return mix64((uint64_t)v + 0x9E3779B97F4A7C15ull * (uint64_t)(level + 1)) >> 34;
}
__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- level 0: generic block CSR copy of the regularised KEO (owned columns only) ------------------
__global__ void k_l0_count(const int32_t *rowptr, const int32_t *col, int64_t No, int32_t *cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > No) return;
  int c = 0;
  if (i < No)
    for (int p = rowptr[i]; p < rowptr[i + 1]; p++) c += col[p] < No;
  cnt[i] = c;
}
__global__ void k_l0_fill(const int32_t *rowptr, const int32_t *col, const int32_t *csr_pos, const double2 *K,
                          const double2 *pd0, const double *pd1, int64_t No, const int32_t *orow, int32_t *ocol,
                          B22 *oval) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= No) return;
  int o = orow[i];
  for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const int j = col[p];
    if (j >= No) continue;
    const double2 k = K[csr_pos[p]];
    B22 v = b22(k.x, -k.y, k.y, k.x);  // complex block as a real 2x2 (src/parameter_matrix_keo.cpp:150-160)
    if (j == i) {                      // + keo_regularized diagonal block (src/keo_regularized.cpp:233-259)
      const double2 d0 = pd0[i];
      const double d1 = pd1[i];
      v.a += d0.x; v.b += d1; v.c += d1; v.d += d0.y;
    }
    ocol[o] = j;
    oval[o] = v;
    o++;
  }
}
// level 0, per state: off-diagonal absolute row sums of the owned block of K, straight from the SELL-32
// storage (both scalar rows of a complex block row have the same one).  They depend on mu through
// |cos a| + |sin a|, so they are refreshed with the diagonal whenever the matrix changes -- a hierarchy
// kept across a continuation run must not smooth with stale bounds.
__global__ void k_l0_offsum(const int32_t *slice_off, const int32_t *sell_pos, const int32_t *col, const double2 *K,
                            int64_t No, double *offsum) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= No) return;
  const int64_t q = sell_pos ? (int64_t)sell_pos[i] : i;  // SELL position of row i (SELL-32-sigma, mesh.cu)
  const int base = slice_off[q >> 5], end = slice_off[(q >> 5) + 1];
  double s = 0.0;
  for (int p = base + (int)(q & 31); p < end; p += 32) {
    const int c = col[p];
    if (c != i && c < No) {
      const double2 k = K[p];
      s += fabs(k.x) + fabs(k.y);
    }
  }
  offsum[i] = s;
}
// level 0, per state: 1 / diagonal and 1 / absolute row sum of the regularised KEO
__global__ void k_l0_dinv(const double2 *K, const int32_t *diag_slot, const double2 *pd0, const double *pd1,
                          const double *offsum, int64_t No, double2 *dinv, double2 *sinv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= No) return;
  const double kd = K[diag_slot[i]].x;
  const double2 d0 = pd0[i];
  dinv[i] = make_double2(1.0 / (kd + d0.x), 1.0 / (kd + d0.y));
  const double off = offsum[i] + fabs(pd1[i]);
  sinv[i] = make_double2(1.0 / (off + fabs(kd + d0.x)), 1.0 / (off + fabs(kd + d0.y)));
}
__global__ void k_dinv(const int32_t *rowptr, const int32_t *col, const B22 *val, int64_t n, double2 *dinv,
                       double2 *sinv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 d = make_double2(1.0, 1.0);
  double s0 = 0.0, s1 = 0.0;
  for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const B22 v = val[p];
    if (col[p] == i) d = make_double2(1.0 / v.a, 1.0 / v.d);
    s0 += fabs(v.a) + fabs(v.b);
    s1 += fabs(v.c) + fabs(v.d);
  }
  dinv[i] = d;
  sinv[i] = make_double2(1.0 / s0, 1.0 / s1);
}

// ---- MIS-2 aggregation -------------------------------------------------------------------------
__global__ void k_fill_i32(int32_t *a, int64_t n, int32_t v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_mis_pack(int64_t n, int level, const int32_t *state, uint64_t *T) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  T[v] = ((uint64_t)state[v] << 62) | (amg_priority((uint32_t)v, level) << 32) | (uint64_t)(uint32_t)v;
}
__global__ void k_mis_max(int64_t n, const int32_t *rowptr, const int32_t *col, const uint64_t *Tin, uint64_t *Tout) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  uint64_t m = Tin[v];
  for (int p = rowptr[v]; p < rowptr[v + 1]; p++) {
    const uint64_t t = Tin[col[p]];
    m = t > m ? t : m;
  }
  Tout[v] = m;
}
__global__ void k_mis_update(int64_t n, const uint64_t *T0, const uint64_t *T2, int32_t *state, int32_t *undecided) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n || state[v] != ST_UND) return;
  const uint64_t t2 = T2[v];
  if (t2 == T0[v]) state[v] = ST_IN;
  else if ((int)(t2 >> 62) == ST_IN) state[v] = ST_OUT;
  else atomicAdd(undecided, 1);
}
__global__ void k_root_flag(int64_t n, const int32_t *state, int32_t *flag) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v <= n) flag[v] = (v < n && state[v] == ST_IN) ? 1 : 0;
}
__global__ void k_agg_phase12(int64_t n, const int32_t *rowptr, const int32_t *col, const int32_t *state,
                              const int32_t *rootnum, int32_t *agg2) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  int a = -1;
  if (state[v] == ST_IN) {
    a = rootnum[v];
  } else {
    for (int p = rowptr[v]; p < rowptr[v + 1]; p++) {
      const int u = col[p];
      if (state[u] == ST_IN) a = rootnum[u];  // at most one root is adjacent (roots are > 2 apart)
    }
  }
  agg2[v] = a;
}
__global__ void k_agg_phase3(int64_t n, const int32_t *rowptr, const int32_t *col, int level, const int32_t *agg2,
                             int32_t *agg) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  int a = agg2[v];
  if (a < 0) {
    uint64_t best = 0;
    for (int p = rowptr[v]; p < rowptr[v + 1]; p++) {
      const int u = col[p];
      if (agg2[u] < 0) continue;
      const uint64_t k = (amg_priority((uint32_t)u, level) << 32) | (uint64_t)(uint32_t)u;
      if (a < 0 || k > best) {
        best = k;
        a = agg2[u];
      }
    }
  }
  agg[v] = a;
}

// ---- tentative prolongator weights ---------------------------------------------------------------
__global__ void k_member_keys(int64_t n, const int32_t *agg, uint64_t *keys, uint32_t *vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = ((uint64_t)(uint32_t)agg[i] << 32) | (uint64_t)(uint32_t)i;
  vals[i] = (uint32_t)i;
}
__global__ void k_rowptr_from_keys(const uint64_t *keys, int64_t nkeys, int64_t nrows, int32_t *rowptr) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > nrows) return;
  rowptr[r] = (int32_t)lower_bound_u64(keys, nkeys, (uint64_t)r << 32);
}
// s_c[I] = sqrt(sum_{i in I} s_i^2), members in ascending order (s == NULL: s_i = 1)
__global__ void k_agg_weight(int64_t nc, const uint64_t *keys, const int32_t *mrow, const double *s, double *sc) {
  const int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (I >= nc) return;
  double sum = 0.0;
  for (int p = mrow[I]; p < mrow[I + 1]; p++) {
    const double si = s ? s[(uint32_t)(keys[p] & 0xFFFFFFFFull)] : 1.0;
    sum = __dadd_rn(sum, __dmul_rn(si, si));
  }
  sc[I] = sqrt(sum);
}
__global__ void k_p0(int64_t n, const int32_t *agg, const double *s, const double *sc, double *p0) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  p0[i] = (s ? s[i] : 1.0) / sc[agg[i]];
}
__global__ void k_p0_csr(int64_t n, const int32_t *agg, const double *p0, int32_t *rowptr, int32_t *col, B22 *val) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  rowptr[i] = (int32_t)i;
  if (i < n) {
    col[i] = agg[i];
    val[i] = b22(p0[i], 0.0, 0.0, p0[i]);
  }
}

// ---- expand-sort-compress SpGEMM ---------------------------------------------------------------------
struct Csr {
  int64_t n;
  const int32_t *rowptr, *col;
  const B22 *val;
};
// C = A * B: products of row i = sum over its entries p of |B row col[p]|
__global__ void k_mm_count(Csr A, Csr B, int64_t *cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > A.n) return;
  int64_t c = 0;
  if (i < A.n)
    for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; p++) {
      const int k = A.col[p];
      c += B.rowptr[k + 1] - B.rowptr[k];
    }
  cnt[i] = c;
}
// rows [r0, r1) of the product (one panel): positions are relative to off[r0]
__global__ void k_mm_expand(Csr A, Csr B, const int64_t *off, int64_t r0, int64_t r1, uint64_t *keys, uint32_t *idx,
                            B22 *vals) {
  const int64_t i = r0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= r1) return;
  int64_t t = off[i] - off[r0];
  for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; p++) {
    const int k = A.col[p];
    const B22 a = A.val[p];
    for (int q = B.rowptr[k]; q < B.rowptr[k + 1]; q++, t++) {
      keys[t] = ((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)B.col[q];
      idx[t] = (uint32_t)t;
      vals[t] = mul(a, B.val[q]);
    }
  }
}
// C = A^T * B for A, B with the same rows: products of row i = |A row i| * |B row i|, key (A.col, B.col)
__global__ void k_atb_count(Csr A, Csr B, int64_t *cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > A.n) return;
  cnt[i] = i < A.n ? (int64_t)(A.rowptr[i + 1] - A.rowptr[i]) * (B.rowptr[i + 1] - B.rowptr[i]) : 0;
}
__global__ void k_atb_expand(Csr A, Csr B, const int64_t *off, int64_t r0, int64_t r1, uint64_t *keys, uint32_t *idx,
                             B22 *vals) {
  const int64_t i = r0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= r1) return;
  int64_t t = off[i] - off[r0];
  for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; p++) {
    const B22 a = A.val[p];
    const uint64_t hi = (uint64_t)(uint32_t)A.col[p] << 32;
    for (int q = B.rowptr[i]; q < B.rowptr[i + 1]; q++, t++) {
      keys[t] = hi | (uint64_t)(uint32_t)B.col[q];
      idx[t] = (uint32_t)t;
      vals[t] = tmul(a, B.val[q]);
    }
  }
}
__global__ void k_head_flags(const uint64_t *keys, int64_t n, int32_t *head) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  head[i] = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}
__global__ void k_unique_starts(const uint64_t *keys, const int32_t *head, const int32_t *excl, int64_t n,
                                int64_t nuniq, uint64_t *ukeys, int32_t *start) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == n) start[nuniq] = (int32_t)n;
  if (i >= n || !head[i]) return;
  ukeys[excl[i]] = keys[i];
  start[excl[i]] = (int32_t)i;
}
// one thread per unique key: sum of its products in emission order (the sort is stable)
__global__ void k_segment_sum(const uint64_t *ukeys, const int32_t *start, const uint32_t *idx, const B22 *vals,
                              int64_t nuniq, int32_t *col, B22 *out) {
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= nuniq) return;
  B22 s = b22(0.0, 0.0, 0.0, 0.0);
  for (int p = start[u]; p < start[u + 1]; p++) acc(s, vals[idx[p]]);
  col[u] = (int32_t)(ukeys[u] & 0xFFFFFFFFull);
  out[u] = s;
}

struct CsrOut {
  DBuf<int32_t> rowptr, col;
  DBuf<B22> val;
  int64_t n = 0, nnz = 0;
  Csr view() const { return Csr{n, rowptr.p, col.p, val.p}; }
};

// second-level merge of the panels of A^T B: the compressed entries of all panels, concatenated, as "products"
__global__ void k_merge_keys(const uint64_t *ukeys, int64_t n, uint64_t *keys, uint32_t *idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = ukeys[i];
    idx[i] = (uint32_t)i;
  }
}
__global__ void k_panel_cols(const uint64_t *ukeys, int64_t n, int32_t *col) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) col[i] = (int32_t)(ukeys[i] & 0xFFFFFFFFull);
}

// sort `total` (key, index) pairs and sum the products of every unique key in emission order:
// out: nuniq, ukeys (nuniq), usum (nuniq)
void sort_compress(Ctx *ctx, Temp &tmp, TBuf<uint64_t> &keys, TBuf<uint32_t> &idx, const B22 *vals, int64_t total,
                   int64_t &nuniq, TBuf<uint64_t> &ukeys, TBuf<B22> &usum) {
  sort_pairs_u64(ctx, tmp, keys, idx, total);
  TBuf<int32_t> head, excl, start, ucol;
  head.alloc(total + 1);
  excl.alloc(total + 1);
  ALAUNCH(ctx, k_head_flags, total + 1, keys.p, total, head.p);
  exclusive_scan_i32(ctx, tmp, head.p, excl.p, total + 1);
  nuniq = fetch(ctx, excl.p + total);
  ukeys.alloc(nuniq);
  start.alloc(nuniq + 1);
  ucol.alloc(nuniq);
  usum.alloc(nuniq);
  ALAUNCH(ctx, k_unique_starts, total + 1, keys.p, head.p, excl.p, total, nuniq, ukeys.p, start.p);
  ALAUNCH(ctx, k_segment_sum, nuniq, ukeys.p, start.p, idx.p, vals, nuniq, ucol.p, usum.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// Expand (already counted) -> sort -> compress, in ROW PANELS of at most ~panel_products products each: the
// temporaries of one panel (44 B per product + the sort's double buffers) are reused by the next, so a product
// with 460M terms (A P on 8M vertices) needs ~4 GB of the temporaries' arena instead of ~29 GB.  rows_disjoint: every output row
// is produced by one panel only (C = A B): the panels' results are simply concatenated, bit-identical to one
// panel.  Otherwise (C = A^T B) the compressed panels are merged by one more, small, sort-compress pass.
template <typename ExpandFn>
void esc_finish(Ctx *ctx, Temp &tmp, TBuf<int64_t> &cnt, int64_t nrows_in, int64_t nrows_out, bool rows_disjoint,
                ExpandFn expand, CsrOut &C) {
  TBuf<int64_t> off;
  off.alloc(nrows_in + 1);
  exclusive_scan_i64(ctx, tmp, cnt.p, off.p, nrows_in + 1);
  cnt.release();
  const int64_t total = fetch(ctx, off.p + nrows_in);
  const int64_t budget = ctx->amg_panel_products > 0 ? ctx->amg_panel_products : ((int64_t)32 << 20);
  const int64_t npanels = std::max<int64_t>(1, std::min<int64_t>(cdiv(total, budget), nrows_in > 0 ? nrows_in : 1));
  std::vector<TBuf<uint64_t>> pk(npanels);
  std::vector<TBuf<B22>> pv(npanels);
  std::vector<int64_t> pn(npanels, 0);
  int64_t nsum = 0;
  for (int64_t q = 0; q < npanels; q++) {
    const int64_t r0 = nrows_in * q / npanels, r1 = nrows_in * (q + 1) / npanels;
    const int64_t ptotal = fetch(ctx, off.p + r1) - fetch(ctx, off.p + r0);
    check_count(ptotal, "a panel of a sparse product");
    if (ptotal == 0) continue;
    TBuf<uint64_t> keys;
    TBuf<uint32_t> idx;
    TBuf<B22> vals;
    keys.alloc(ptotal);
    idx.alloc(ptotal);
    vals.alloc(ptotal);
    expand(off.p, r0, r1, keys.p, idx.p, vals.p);
    sort_compress(ctx, tmp, keys, idx, vals.p, ptotal, pn[q], pk[q], pv[q]);
    nsum += pn[q];
  }
  off.release();
  check_count(nsum, "a sparse product");
  // concatenate the panels (row-ordered for rows_disjoint)
  TBuf<uint64_t> ukeys;
  TBuf<B22> uval;
  int64_t nuniq = nsum;
  if (npanels == 1) {
    ukeys.swap(pk[0]);
    uval.swap(pv[0]);
  } else {
    ukeys.alloc(nsum);
    uval.alloc(nsum);
    int64_t at = 0;
    for (int64_t q = 0; q < npanels; q++) {
      if (!pn[q]) continue;
      CUDA_CHECK(cudaMemcpyAsync(ukeys.p + at, pk[q].p, sizeof(uint64_t) * pn[q], cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(uval.p + at, pv[q].p, sizeof(B22) * pn[q], cudaMemcpyDeviceToDevice, ctx->stream));
      at += pn[q];
      pk[q].release();
      pv[q].release();
    }
    if (!rows_disjoint) {
      TBuf<uint64_t> keys, mk;
      TBuf<uint32_t> idx;
      TBuf<B22> mv;
      keys.alloc(nsum);
      idx.alloc(nsum);
      ALAUNCH(ctx, k_merge_keys, nsum, ukeys.p, nsum, keys.p, idx.p);
      sort_compress(ctx, tmp, keys, idx, uval.p, nsum, nuniq, mk, mv);
      ukeys.swap(mk);
      uval.swap(mv);
    }
  }
  C.n = nrows_out;
  C.nnz = nuniq;
  C.col.alloc(nuniq);
  C.val.alloc(nuniq);
  C.rowptr.alloc(nrows_out + 1);
  ALAUNCH(ctx, k_panel_cols, nuniq, ukeys.p, nuniq, C.col.p);
  if (nuniq) CUDA_CHECK(cudaMemcpyAsync(C.val.p, uval.p, sizeof(B22) * nuniq, cudaMemcpyDeviceToDevice, ctx->stream));
  ALAUNCH(ctx, k_rowptr_from_keys, nrows_out + 1, ukeys.p, nuniq, nrows_out, C.rowptr.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void spgemm(Ctx *ctx, Temp &tmp, const Csr &A, const Csr &B, CsrOut &C) {
  TBuf<int64_t> cnt;
  cnt.alloc(A.n + 1);
  ALAUNCH(ctx, k_mm_count, A.n + 1, A, B, cnt.p);
  esc_finish(ctx, tmp, cnt, A.n, A.n, true,
             [&](const int64_t *off, int64_t r0, int64_t r1, uint64_t *keys, uint32_t *idx, B22 *vals) {
               ALAUNCH(ctx, k_mm_expand, r1 - r0, A, B, off, r0, r1, keys, idx, vals);
             },
             C);
}
void spgemm_atb(Ctx *ctx, Temp &tmp, const Csr &A, const Csr &B, int64_t ncols_a, CsrOut &C) {
  TBuf<int64_t> cnt;
  cnt.alloc(A.n + 1);
  ALAUNCH(ctx, k_atb_count, A.n + 1, A, B, cnt.p);
  esc_finish(ctx, tmp, cnt, A.n, ncols_a, false,
             [&](const int64_t *off, int64_t r0, int64_t r1, uint64_t *keys, uint32_t *idx, B22 *vals) {
               ALAUNCH(ctx, k_atb_expand, r1 - r0, A, B, off, r0, r1, keys, idx, vals);
             },
             C);
}

// P = P0 - omega D^-1 (A P0), on the pattern of T = A P0
__global__ void k_smooth_p(int64_t n, const int32_t *rowptr, const int32_t *col, B22 *val /* in: T, out: P */,
                           const int32_t *agg, const double *p0, const double2 *dinv, double omega) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 di = dinv[i];
  const double wx = omega * di.x, wy = omega * di.y;
  const int a = agg[i];
  const double p = p0[i];
  for (int q = rowptr[i]; q < rowptr[i + 1]; q++) {
    const B22 t = val[q];
    B22 v = b22(-(wx * t.a), -(wx * t.b), -(wy * t.c), -(wy * t.d));
    if (col[q] == a) {
      v.a += p;
      v.d += p;
    }
    val[q] = v;
  }
}
__global__ void k_transpose_keys(int64_t n, const int32_t *rowptr, const int32_t *col, uint64_t *keys, uint32_t *vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int q = rowptr[i]; q < rowptr[i + 1]; q++) {
    keys[q] = ((uint64_t)(uint32_t)col[q] << 32) | (uint64_t)(uint32_t)i;
    vals[q] = (uint32_t)q;
  }
}
__global__ void k_transpose_fill(const uint64_t *keys, const uint32_t *vals, int64_t nnz, int32_t *r_fine,
                                 int32_t *r_pos) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nnz) return;
  r_fine[q] = (int32_t)(keys[q] & 0xFFFFFFFFull);
  r_pos[q] = (int32_t)vals[q];
}

// ---- SELL-32 conversion of a block CSR ------------------------------------------------------------
__global__ void k_slice_width(const int32_t *rowptr, int64_t n, int64_t nslices, int32_t *width32) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s > nslices) return;
  int w = 0;
  if (s < nslices)
    for (int64_t r = s * 32; r < s * 32 + 32 && r < n; r++) w = max(w, rowptr[r + 1] - rowptr[r]);
  width32[s] = w * 32;
}
__global__ void k_sell_fill(const int32_t *rowptr, const int32_t *col, const B22 *val, const int32_t *slice_off,
                            int64_t n, int64_t nslices, int32_t *scol, B22 *sval) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nslices * 32) return;
  const int base = slice_off[r >> 5] + (int)(r & 31);
  const int w = (slice_off[(r >> 5) + 1] - slice_off[r >> 5]) >> 5;
  int k = 0;
  if (r < n)
    for (int p = rowptr[r]; p < rowptr[r + 1]; p++, k++) {
      scol[base + 32 * k] = col[p];
      sval[base + 32 * k] = val[p];
    }
  for (; k < w; k++) {  // padding: value 0, the row's own column
    scol[base + 32 * k] = (int32_t)(r < n ? r : 0);
    sval[base + 32 * k] = b22(0.0, 0.0, 0.0, 0.0);
  }
}

// ---- apply kernels of the coarse levels -------------------------------------------------------------------
enum { M_APPLY = 0, M_RESID = 1, M_CHEB = 2 };
struct LevelArgs {
  int64_t n, nslices;
  const int32_t *slice_off, *scol;
  const B22 *sval;
  const double2 *x;
  double2 *y;
  const double2 *b;
  double2 *d;
  const double2 *dinv;
  double c1, c2;
  const KrylovState *gate;
};
// one thread per block row; M_APPLY: y = A x; M_RESID: y = b - A x;
// M_CHEB: t = D^-1 (b - A x), d = c1 d + c2 t, y = x + d
template <int MODE>
__global__ void __launch_bounds__(256) k_level_apply(const LevelArgs L) {
  if (L.gate && L.gate->done) return;
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t slice = row >> 5;
  if (slice >= L.nslices) return;
  const int lane = threadIdx.x & 31;
  double2 s = make_double2(0.0, 0.0);
  const int pend = __ldg(L.slice_off + slice + 1);
  int p = __ldg(L.slice_off + slice) + lane;
  for (; p + 32 < pend; p += 64) {
    const int c0 = __ldg(L.scol + p), c1 = __ldg(L.scol + p + 32);
    const B22 v0 = L.sval[p], v1 = L.sval[p + 32];
    const double2 x0 = __ldg(L.x + c0), x1 = __ldg(L.x + c1);
    s.x += v0.a * x0.x + v0.b * x0.y;
    s.y += v0.c * x0.x + v0.d * x0.y;
    s.x += v1.a * x1.x + v1.b * x1.y;
    s.y += v1.c * x1.x + v1.d * x1.y;
  }
  for (; p < pend; p += 32) {
    const int c0 = __ldg(L.scol + p);
    const B22 v0 = L.sval[p];
    const double2 x0 = __ldg(L.x + c0);
    s.x += v0.a * x0.x + v0.b * x0.y;
    s.y += v0.c * x0.x + v0.d * x0.y;
  }
  if (row >= L.n) return;
  if (MODE == M_RESID) {
    const double2 bb = L.b[row];
    s = make_double2(bb.x - s.x, bb.y - s.y);
  } else if (MODE == M_CHEB) {
    const double2 bb = L.b[row], di = L.dinv[row], xi = L.x[row];
    double2 dd = make_double2(L.c2 * (di.x * (bb.x - s.x)), L.c2 * (di.y * (bb.y - s.y)));
    if (L.c1 != 0.0) {
      const double2 dold = L.d[row];
      dd.x += L.c1 * dold.x;
      dd.y += L.c1 * dold.y;
    }
    if (L.d) L.d[row] = dd;
    s = make_double2(xi.x + dd.x, xi.y + dd.y);
  }
  L.y[row] = s;
}

// Small levels (a few thousand rows of ~50 blocks): one thread per row is latency bound, so LPR lanes
// share a row of the block CSR (consecutive lanes read consecutive 32-byte blocks) and combine with a
// fixed xor tree.  Same modes as k_level_apply.
struct LevelCsrArgs {
  int64_t n;
  const int32_t *rowptr, *col;
  const B22 *val;
  const double2 *x;
  double2 *y;
  const double2 *b;
  double2 *d;
  const double2 *dinv;
  double c1, c2;
  const KrylovState *gate;
};
template <int MODE, int LPR>
__global__ void __launch_bounds__(256) k_level_apply_csr(const LevelCsrArgs L) {
  if (L.gate && L.gate->done) return;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = t / LPR;
  const int sl = (int)(t % LPR);
  double2 s = make_double2(0.0, 0.0);
  if (row < L.n) {
    const int e = __ldg(L.rowptr + row + 1);
    for (int p = __ldg(L.rowptr + row) + sl; p < e; p += LPR) {
      const B22 v = L.val[p];
      const double2 xv = __ldg(L.x + __ldg(L.col + p));
      s.x += v.a * xv.x + v.b * xv.y;
      s.y += v.c * xv.x + v.d * xv.y;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
  }
  if (row >= L.n || sl != 0) return;
  if (MODE == M_RESID) {
    const double2 bb = L.b[row];
    s = make_double2(bb.x - s.x, bb.y - s.y);
  } else if (MODE == M_CHEB) {
    const double2 bb = L.b[row], di = L.dinv[row], xi = L.x[row];
    double2 dd = make_double2(L.c2 * (di.x * (bb.x - s.x)), L.c2 * (di.y * (bb.y - s.y)));
    if (L.c1 != 0.0) {
      const double2 dold = L.d[row];
      dd.x += L.c1 * dold.x;
      dd.y += L.c1 * dold.y;
    }
    if (L.d) L.d[row] = dd;
    s = make_double2(xi.x + dd.x, xi.y + dd.y);
  }
  L.y[row] = s;
}

// d = c2 D^-1 b,  x = d   (first Chebyshev step from a zero initial guess)
__global__ void k_cheb_first(int64_t n, const double2 *b, const double2 *dinv, double c2, double2 *d, double2 *x,
                             const KrylovState *gate) {
  if (gate && gate->done) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 bb = b[i], di = dinv[i];
  const double2 v = make_double2(c2 * (di.x * bb.x), c2 * (di.y * bb.y));
  if (d) d[i] = v;
  x[i] = v;
}

__device__ __forceinline__ B22 load_b22(const B22 *p) { return *p; }
__device__ __forceinline__ B22 load_b22(const float4 *p) {
  const float4 f = *p;
  B22 r;
  r.a = f.x;
  r.b = f.y;
  r.c = f.z;
  r.d = f.w;
  return r;
}
__global__ void k_b22_to_f32(int64_t n, const B22 *in, float4 *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const B22 v = in[i];
  out[i] = make_float4((float)v.a, (float)v.b, (float)v.c, (float)v.d);
}
__global__ void k_c64_to_c32(int64_t n, const double2 *in, float2 *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 v = in[i];
  out[i] = make_float2((float)v.x, (float)v.y);
}

// b_c[I] = sum_i P_iI^T r_i : LPR lanes per coarse node, fixed lane-strided order + xor tree
template <int LPR, typename PV>
__global__ void __launch_bounds__(256) k_restrict(int64_t nc, const int32_t *r_rowptr, const int32_t *r_fine,
                                                  const int32_t *r_pos, const PV *p_val, const double2 *r,
                                                  double2 *bc, const KrylovState *gate) {
  if (gate && gate->done) return;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t I = t / LPR;
  const int sl = (int)(t % LPR);
  double2 s = make_double2(0.0, 0.0);
  if (I < nc) {
    const int e1 = r_rowptr[I + 1];
    for (int e = r_rowptr[I] + sl; e < e1; e += LPR) {
      const B22 P = load_b22(p_val + r_pos[e]);
      const double2 v = __ldg(r + r_fine[e]);
      s.x += P.a * v.x + P.c * v.y;
      s.y += P.b * v.x + P.d * v.y;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
  }
  if (I < nc && sl == 0) bc[I] = s;
}
// x_i += sum_J P_iJ xc_J
template <typename PV>
__global__ void __launch_bounds__(256) k_prolong(int64_t n, const int32_t *p_rowptr, const int32_t *p_col,
                                                 const PV *p_val, const double2 *xc, double2 *x,
                                                 const KrylovState *gate) {
  if (gate && gate->done) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 s = x[i];
  for (int q = p_rowptr[i]; q < p_rowptr[i + 1]; q++) {
    const B22 P = load_b22(p_val + q);
    const double2 v = __ldg(xc + p_col[q]);
    s.x += P.a * v.x + P.b * v.y;
    s.y += P.c * v.x + P.d * v.y;
  }
  x[i] = s;
}
// x = Inv b, one warp per row of the dense inverse
__global__ void __launch_bounds__(256) k_dense_apply(int64_t n2, const double *inv, const double *b, double *x,
                                                     const KrylovState *gate) {
  if (gate && gate->done) return;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n2) return;
  double s = 0.0;
  for (int64_t j = lane; j < n2; j += 32) s += inv[row * n2 + j] * b[j];
  s = warp_sum(s);
  if (lane == 0) x[row] = s;
}
__global__ void k_copy2(int64_t n, const double2 *in, double2 *out, const KrylovState *gate) {
  if (gate && gate->done) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// ---- power iteration helpers (set-up only; deterministic two-stage sums) ---------------------------------
__global__ void k_start_vector(int64_t n, double2 *x) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const uint64_t z = mix64((uint64_t)(2 * i + c) + 0x51ED270B7A2F3C15ull);
    u[c] = __dadd_rn(__dmul_rn((double)(z >> 11), 2.0 / 9007199254740992.0), -1.0);
  }
  x[i] = make_double2(u[0], u[1]);
}
__global__ void __launch_bounds__(256) k_dot_partial(int64_t n, const double2 *x, const double2 *y, double *partials) {
  __shared__ double red[8];
  const int64_t i0 = (int64_t)blockIdx.x * 512 + threadIdx.x, i1 = i0 + 256;
  double c = 0.0;
  if (i0 < n) c = x[i0].x * y[i0].x + x[i0].y * y[i0].y;
  if (i1 < n) c += x[i1].x * y[i1].x + x[i1].y * y[i1].y;
  const double s = block_sum<8>(c, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_sum_partials(int64_t np, const double *partials, double *out) {
  __shared__ double red[32];
  double c = 0.0;
  for (int64_t i = threadIdx.x; i < np; i += 1024) c += partials[i];
  const double s = block_sum<32>(c, red);
  if (threadIdx.x == 0) out[0] = s;
}
__global__ void k_scale2(int64_t n, double a, const double2 *in, double2 *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_double2(a * in[i].x, a * in[i].y);
}

double local_dot(Ctx *ctx, DBuf<double> &scratch, int64_t n, const double2 *x, const double2 *y) {
  const int64_t np = cdiv(n, 512);
  scratch.ensure(np + 1);
  k_dot_partial<<<(unsigned)np, 256, 0, ctx->stream>>>(n, x, y, scratch.p + 1);
  k_sum_partials<<<1, 1024, 0, ctx->stream>>>(np, scratch.p + 1, scratch.p);
  ctx->launches += 2;
  CUDA_CHECK(cudaGetLastError());
  return fetch(ctx, scratch.p);
}

// ---- level operations used by both set-up and the V-cycle ---------------------------------------------------
void level_apply(Ctx *ctx, AmgLevel &L, int lev, int mode, const double2 *x, double2 *y, const double2 *b, double2 *d,
                 double c1, double c2, const KrylovState *gate, const double2 *scale) {
  if (L.n == 0) return;
  if (lev == 0) {
    ApplyArgs A;
    memset(&A, 0, sizeof(A));
    A.No = ctx->No;
    A.nslices = ctx->nslices;
    A.rowptr = ctx->rowptr.p;
    A.slice_off = ctx->slice_off.p;
    A.sell_row = ctx->sell_permuted ? ctx->sell_row.p : nullptr;
    A.col = ctx->col.p;
    A.val = ctx->Kval.p;
    if (ctx->amg_mixed && mode != M_APPLY) A.val32 = ctx->Kval32.p;
    A.x = x;
    A.y = y;
    A.a = 1.0;
    A.d0 = ctx->pd0.p;
    A.d1 = ctx->pd1.p;
    A.bvec = b;
    A.dvec = d;
    A.dinv = scale;
    A.c1 = c1;
    A.c2 = c2;
    A.gate = gate;
    launch_apply(ctx, EPI_DIAG, mode == M_APPLY ? FUSE_NONE : mode == M_RESID ? FUSE_RESID : FUSE_CHEB, A);
    return;
  }
  if (L.n <= SMALL_LEVEL) {
    constexpr int LPR = 8;
    LevelCsrArgs A{L.n, L.rowptr.p, L.col.p, L.val.p, x, y, b, d, scale, c1, c2, gate};
    const unsigned grid = (unsigned)cdiv(L.n * LPR, 256);
    if (mode == M_APPLY) k_level_apply_csr<M_APPLY, LPR><<<grid, 256, 0, ctx->stream>>>(A);
    else if (mode == M_RESID) k_level_apply_csr<M_RESID, LPR><<<grid, 256, 0, ctx->stream>>>(A);
    else k_level_apply_csr<M_CHEB, LPR><<<grid, 256, 0, ctx->stream>>>(A);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  LevelArgs A{L.n, L.nslices, L.slice_off.p, L.scol.p, L.sval.p, x, y, b, d, scale, c1, c2, gate};
  const unsigned grid = (unsigned)cdiv(L.nslices * 32, 256);
  if (mode == M_APPLY) k_level_apply<M_APPLY><<<grid, 256, 0, ctx->stream>>>(A);
  else if (mode == M_RESID) k_level_apply<M_RESID><<<grid, 256, 0, ctx->stream>>>(A);
  else k_level_apply<M_CHEB><<<grid, 256, 0, ctx->stream>>>(A);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}

// Rayleigh-quotient estimate of lambda_max(D^-1 A) after POWER_ITS power iterations (oracle/amg.py:lambda_max)
double estimate_lambda(Ctx *ctx, AmgLevel &L, int lev, DBuf<double> &scratch) {
  const int64_t n = L.n;
  double2 *x = L.x.p, *y = L.d.p, *junk = L.x2.p, *zero = L.r.p;
  CUDA_CHECK(cudaMemsetAsync(zero, 0, sizeof(double2) * n, ctx->stream));
  ALAUNCH(ctx, k_start_vector, n, n, x);
  double nrm = sqrt(local_dot(ctx, scratch, n, x, x));
  ALAUNCH(ctx, k_scale2, n, n, 1.0 / nrm, x, x);
  double lam = 0.0;
  for (int it = 0; it < POWER_ITS; it++) {
    // y = D^-1 A x  ==  Chebyshev step with b = 0, c1 = 0, c2 = -1 (d receives the result)
    level_apply(ctx, L, lev, M_CHEB, x, junk, zero, y, 0.0, -1.0, nullptr, L.dinv.p);
    lam = local_dot(ctx, scratch, n, x, y);
    nrm = sqrt(local_dot(ctx, scratch, n, y, y));
    ALAUNCH(ctx, k_scale2, n, n, 1.0 / nrm, y, x);
  }
  return lam;
}

void l0_refresh_diag(Ctx *ctx, AmgLevel &L) {
  ALAUNCH(ctx, k_l0_offsum, ctx->No, ctx->slice_off.p, ctx->sell_permuted ? ctx->sell_pos.p : nullptr, ctx->col.p,
          ctx->Kval.p, ctx->No, L.offsum.p);
  ALAUNCH(ctx, k_l0_dinv, ctx->No, ctx->Kval.p, ctx->diag_slot.p, ctx->pd0.p, ctx->pd1.p, L.offsum.p, ctx->No,
          L.dinv.p, L.sinv.p);
}

void alloc_level_vectors(Ctx *ctx, AmgLevel &L, int lev) {
  const int64_t n = (lev == 0 ? ctx->Nl : L.n) > 0 ? (lev == 0 ? ctx->Nl : L.n) : 1;
  for (DBuf<double2> *v : {&L.b, &L.x, &L.x2, &L.d, &L.r}) {
    v->alloc(n);
    CUDA_CHECK(cudaMemsetAsync(v->p, 0, sizeof(double2) * n, ctx->stream));
  }
}

void build_sell(Ctx *ctx, Temp &tmp, AmgLevel &L) {
  const int64_t ns = cdiv(L.n, 32);
  L.nslices = ns;
  TBuf<int32_t> w32;
  w32.alloc(ns + 1);
  ALAUNCH(ctx, k_slice_width, ns + 1, L.rowptr.p, L.n, ns, w32.p);
  L.slice_off.alloc(ns + 1);
  exclusive_scan_i32(ctx, tmp, w32.p, L.slice_off.p, ns + 1);
  L.nstored = fetch(ctx, L.slice_off.p + ns);
  L.scol.alloc(L.nstored);
  L.sval.alloc(L.nstored);
  ALAUNCH(ctx, k_sell_fill, ns * 32, L.rowptr.p, L.col.p, L.val.p, L.slice_off.p, L.n, ns, L.scol.p, L.sval.p);
}

// MIS-2 aggregation of the block graph (rowptr, col) of `n` nodes; fills L.agg, returns #aggregates
int64_t aggregate(Ctx *ctx, Temp &tmp, AmgLevel &L, int lev, const int32_t *rowptr, const int32_t *col) {
  const int64_t n = L.n;
  TBuf<int32_t> state, counter, flag, rootnum, agg2;
  TBuf<uint64_t> T0, T1, T2;
  state.alloc(n);
  counter.alloc(1);
  T0.alloc(n);
  T1.alloc(n);
  T2.alloc(n);
  ALAUNCH(ctx, k_fill_i32, n, state.p, n, ST_UND);
  for (int round = 0; round < 1000; round++) {
    CUDA_CHECK(cudaMemsetAsync(counter.p, 0, sizeof(int32_t), ctx->stream));
    ALAUNCH(ctx, k_mis_pack, n, n, lev, state.p, T0.p);
    ALAUNCH(ctx, k_mis_max, n, n, rowptr, col, T0.p, T1.p);
    ALAUNCH(ctx, k_mis_max, n, n, rowptr, col, T1.p, T2.p);
    ALAUNCH(ctx, k_mis_update, n, n, T0.p, T2.p, state.p, counter.p);
    if (fetch(ctx, counter.p) == 0) break;
  }
  flag.alloc(n + 1);
  rootnum.alloc(n + 1);
  ALAUNCH(ctx, k_root_flag, n + 1, n, state.p, flag.p);
  exclusive_scan_i32(ctx, tmp, flag.p, rootnum.p, n + 1);
  const int64_t nc = fetch(ctx, rootnum.p + n);
  agg2.alloc(n);
  L.agg.alloc(n);
  ALAUNCH(ctx, k_agg_phase12, n, n, rowptr, col, state.p, rootnum.p, agg2.p);
  ALAUNCH(ctx, k_agg_phase3, n, n, rowptr, col, lev, agg2.p, L.agg.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return nc;
}

// dense SPD inverse on the host (Cholesky); returns false if a pivot is not positive
bool dense_spd_inverse(std::vector<double> &a, int64_t n) {
  // a pivot below this is rounding noise of a singular matrix (e.g. the pure Laplacian for g = mu = 0)
  double amax = 0.0;
  for (int64_t j = 0; j < n; j++) amax = std::max(amax, std::fabs(a[j * n + j]));
  const double tiny = 1e-13 * amax * (double)(n > 0 ? n : 1);
  // A = L L^T (lower, in place)
  for (int64_t j = 0; j < n; j++) {
    double d = a[j * n + j];
    for (int64_t k = 0; k < j; k++) d -= a[j * n + k] * a[j * n + k];
    if (!(d > tiny)) return false;
    d = std::sqrt(d);
    a[j * n + j] = d;
    for (int64_t i = j + 1; i < n; i++) {
      double s = a[i * n + j];
      for (int64_t k = 0; k < j; k++) s -= a[i * n + k] * a[j * n + k];
      a[i * n + j] = s / d;
    }
  }
  // Linv (lower, in place)
  for (int64_t j = 0; j < n; j++) {
    a[j * n + j] = 1.0 / a[j * n + j];
    for (int64_t i = j + 1; i < n; i++) {
      double s = 0.0;
      for (int64_t k = j; k < i; k++) s -= a[i * n + k] * a[k * n + j];
      a[i * n + j] = s / a[i * n + i];
    }
  }
  // inv = Linv^T Linv (symmetric)
  std::vector<double> inv((size_t)n * n);
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = 0; j <= i; j++) {
      double s = 0.0;
      for (int64_t k = i; k < n; k++) s += a[k * n + i] * a[k * n + j];
      inv[i * n + j] = inv[j * n + i] = s;
    }
  a.swap(inv);
  return true;
}

void build_coarse_inverse(Ctx *ctx, Amg &H) {
  AmgLevel &L = *H.levels.back();
  const int64_t n = L.n, n2 = 2 * n;
  if (n > 4096) NOSH_THROW(NOSH_ESTATE, "AMG: coarsest level has %lld nodes (coarsening stalled)", (long long)n);
  std::vector<int32_t> rp(n + 1), col(L.nb);
  std::vector<B22> val(L.nb);
  CUDA_CHECK(cudaMemcpyAsync(rp.data(), L.rowptr.p, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(col.data(), L.col.p, sizeof(int32_t) * L.nb, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(val.data(), L.val.p, sizeof(B22) * L.nb, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  std::vector<double> a((size_t)n2 * n2, 0.0);
  for (int64_t i = 0; i < n; i++)
    for (int p = rp[i]; p < rp[i + 1]; p++) {
      const int64_t j = col[p];
      a[(2 * i) * n2 + 2 * j] = val[p].a;
      a[(2 * i) * n2 + 2 * j + 1] = val[p].b;
      a[(2 * i + 1) * n2 + 2 * j] = val[p].c;
      a[(2 * i + 1) * n2 + 2 * j + 1] = val[p].d;
    }
  // symmetrise (the Galerkin product is symmetric up to rounding)
  for (int64_t i = 0; i < n2; i++)
    for (int64_t j = 0; j < i; j++) a[i * n2 + j] = a[j * n2 + i] = 0.5 * (a[i * n2 + j] + a[j * n2 + i]);
  if (!dense_spd_inverse(a, n2))
    NOSH_THROW(NOSH_ESTATE,
               "AMG: the regularised KEO is not positive definite (to rounding) on the coarsest level "
               "(g <= 0 and mu == 0?)");
  H.n_coarse2 = n2;
  H.coarse_inv.alloc((size_t)n2 * n2);
  CUDA_CHECK(cudaMemcpyAsync(H.coarse_inv.p, a.data(), sizeof(double) * n2 * n2, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// every device buffer a level keeps
template <class F>
void for_each_buffer(AmgLevel &L, F &&f) {
  f(L.rowptr); f(L.col); f(L.val); f(L.slice_off); f(L.scol); f(L.sval); f(L.dinv); f(L.sinv); f(L.offsum);
  f(L.agg); f(L.p0); f(L.p_rowptr); f(L.p_col); f(L.p_val); f(L.p_val32); f(L.r_rowptr); f(L.r_fine); f(L.r_pos);
  f(L.b); f(L.x); f(L.x2); f(L.d); f(L.r);
}

void build_levels(Ctx *ctx, Amg *H, std::chrono::steady_clock::time_point t0) {
  // NOSH_B200_AMG_TIMING=1: phase times of the set-up on stderr
  static const bool timing = [] {
    const char *e = getenv("NOSH_B200_AMG_TIMING");
    return e && atoi(e) != 0;
  }();
  (void)t0;
  auto tlast = std::chrono::steady_clock::now();
  // phase times: always recorded (nosh_ctx_get_stat "amg.setup.<phase>" = seconds summed over the levels; a
  // stream synchronisation per phase is noise against the set-up), printed with NOSH_B200_AMG_TIMING=1
  auto tick = [&](const char *what, int lev) {
    cudaStreamSynchronize(ctx->stream);
    const auto now = std::chrono::steady_clock::now();
    const double sec = std::chrono::duration<double>(now - tlast).count();
    ctx->stats[std::string("amg.setup.") + what] += sec;
    if (timing) fprintf(stderr, "[amg setup] level %d %-22s %8.1f ms\n", lev, what, 1e3 * sec);
    tlast = now;
  };
  Temp tmp;
  DBuf<double> scratch;
  const int64_t No = ctx->No;
  // ---- level 0: block CSR copy for the set-up products ----------------------------------------
  AmgLevel *L = new AmgLevel();
  H->levels.push_back(L);
  L->n = No;
  {
    TBuf<int32_t> cnt;
    cnt.alloc(No + 1);
    ALAUNCH(ctx, k_l0_count, No + 1, ctx->rowptr.p, ctx->csr_col.p, No, cnt.p);
    L->rowptr.alloc(No + 1);
    exclusive_scan_i32(ctx, tmp, cnt.p, L->rowptr.p, No + 1);
    L->nb = fetch(ctx, L->rowptr.p + No);
    L->col.alloc(L->nb);
    L->val.alloc(L->nb);
    ALAUNCH(ctx, k_l0_fill, No, ctx->rowptr.p, ctx->csr_col.p, ctx->csr_pos.p, ctx->Kval.p, ctx->pd0.p, ctx->pd1.p,
            No, L->rowptr.p, L->col.p, L->val.p);
  }
  L->dinv.alloc(No > 0 ? No : 1);
  L->sinv.alloc(No > 0 ? No : 1);
  L->offsum.alloc(No > 0 ? No : 1);
  l0_refresh_diag(ctx, *L);
  alloc_level_vectors(ctx, *L, 0);
  tick("level-0 block CSR", 0);
  DBuf<double> s_cur;  // null-space weights of the current level (level 0: all ones => not allocated)
  for (int lev = 0;; lev++) {
    L = H->levels[lev];
    if (L->n <= ctx->amg_coarse_max || lev == ctx->amg_max_levels - 1) break;
    const int64_t n = L->n;
    const int64_t nc = aggregate(ctx, tmp, *L, lev, L->rowptr.p, L->col.p);
    if (nc >= n) break;
    L->nc = nc;
    tick("aggregation", lev);
    L->lam = estimate_lambda(ctx, *L, lev, scratch);
    L->omega = SA_DAMPING / L->lam;
    tick("power iteration", lev);
    // tentative prolongator weights
    DBuf<double> sc;
    sc.alloc(nc);
    {
      TBuf<uint64_t> mkeys;
      TBuf<uint32_t> mvals;
      TBuf<int32_t> mrow;
      mkeys.alloc(n);
      mvals.alloc(n);
      mrow.alloc(nc + 1);
      ALAUNCH(ctx, k_member_keys, n, n, L->agg.p, mkeys.p, mvals.p);
      sort_pairs_u64(ctx, tmp, mkeys, mvals, n);
      ALAUNCH(ctx, k_rowptr_from_keys, nc + 1, mkeys.p, n, nc, mrow.p);
      ALAUNCH(ctx, k_agg_weight, nc, nc, mkeys.p, mrow.p, lev == 0 ? nullptr : s_cur.p, sc.p);
      L->p0.alloc(n);
      ALAUNCH(ctx, k_p0, n, n, L->agg.p, lev == 0 ? nullptr : s_cur.p, sc.p, L->p0.p);
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    // T = A P0, then P = P0 - omega D^-1 T in place
    CsrOut P;
    {
      CsrOut P0;
      P0.n = n;
      P0.nnz = n;
      P0.rowptr.alloc(n + 1);
      P0.col.alloc(n);
      P0.val.alloc(n);
      ALAUNCH(ctx, k_p0_csr, n + 1, n, L->agg.p, L->p0.p, P0.rowptr.p, P0.col.p, P0.val.p);
      spgemm(ctx, tmp, Csr{n, L->rowptr.p, L->col.p, L->val.p}, P0.view(), P);
    }
    ALAUNCH(ctx, k_smooth_p, n, n, P.rowptr.p, P.col.p, P.val.p, L->agg.p, L->p0.p, L->dinv.p, L->omega);
    tick("P = (I - w D^-1 A) P0", lev);
    // A_c = P^T (A P)
    AmgLevel *C = new AmgLevel();
    H->levels.push_back(C);
    C->n = nc;
    {
      CsrOut AP, Ac;
      spgemm(ctx, tmp, Csr{n, L->rowptr.p, L->col.p, L->val.p}, P.view(), AP);
      tick("A P", lev);
      spgemm_atb(ctx, tmp, P.view(), AP.view(), nc, Ac);
      tick("P^T (A P)", lev);
      C->nb = Ac.nnz;
      C->rowptr.swap(Ac.rowptr);
      C->col.swap(Ac.col);
      C->val.swap(Ac.val);
    }
    // keep P and its transpose index
    L->p_nnz = P.nnz;
    L->p_rowptr.swap(P.rowptr);
    L->p_col.swap(P.col);
    L->p_val.swap(P.val);
    {
      TBuf<uint64_t> tkeys;
      TBuf<uint32_t> tvals;
      tkeys.alloc(L->p_nnz);
      tvals.alloc(L->p_nnz);
      ALAUNCH(ctx, k_transpose_keys, n, n, L->p_rowptr.p, L->p_col.p, tkeys.p, tvals.p);
      sort_pairs_u64(ctx, tmp, tkeys, tvals, L->p_nnz);
      L->r_rowptr.alloc(nc + 1);
      L->r_fine.alloc(L->p_nnz);
      L->r_pos.alloc(L->p_nnz);
      ALAUNCH(ctx, k_rowptr_from_keys, nc + 1, tkeys.p, L->p_nnz, nc, L->r_rowptr.p);
      ALAUNCH(ctx, k_transpose_fill, L->p_nnz, tkeys.p, tvals.p, L->p_nnz, L->r_fine.p, L->r_pos.p);
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    // the level-0 CSR copy is only needed for the products above (the apply uses the ctx's SELL matrix)
    if (lev == 0 && !ctx->amg_keep_l0) {
      L->col.release();
      L->val.release();
    }
    C->dinv.alloc(nc);
    C->sinv.alloc(nc);
    ALAUNCH(ctx, k_dinv, nc, C->rowptr.p, C->col.p, C->val.p, nc, C->dinv.p, C->sinv.p);
    build_sell(ctx, tmp, *C);
    alloc_level_vectors(ctx, *C, lev + 1);
    s_cur.swap(sc);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    tick("transpose, SELL, vectors", lev);
  }
  build_coarse_inverse(ctx, *H);
  tick("dense coarse inverse", (int)H->levels.size() - 1);
}

void build_hierarchy(Ctx *ctx) {
  const auto t0 = std::chrono::steady_clock::now();
  amg_free(ctx);
  Amg *H = new Amg();
  ctx->amg = H;
  for (auto it = ctx->stats.begin(); it != ctx->stats.end();)
    it = it->first.rfind("amg.setup.", 0) == 0 ? ctx->stats.erase(it) : std::next(it);
  // the arena: one panel of the expand-sort-compress products (32M products x ~112 B), the compressed results of
  // the largest product, the level-0 block-CSR copy the products read and the hierarchy itself
  Arena arena;
  {
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    size_t want = ((size_t)4 << 30) + (size_t)ctx->nb * 90;  // measured high-water mark at 8M vertices: 12.6 GB = 105 B/block
    want = std::min(want, (size_t)ctx->nb * 800 + ((size_t)64 << 20));  // small problems: a few hundred bytes per block
    if (const char *e = getenv("NOSH_B200_AMG_ARENA_MB")) want = (size_t)atoll(e) << 20;
    want = std::min(want, free_b / 2);
    const double t = wall_now();
    arena.init(want);
    ctx->stats["amg.setup.arena"] = wall_now() - t;
    ctx->stats["amg.arena_bytes"] = (double)arena.cap;
  }
  const AllocStats a0 = g_alloc_stats;
  g_arena = g_dbuf_arena = &arena;
  g_fallback_bytes = 0;
  auto leave = [&] {
    g_arena = g_dbuf_arena = nullptr;
    cudaStreamSynchronize(ctx->stream);
    ctx->stats["amg.arena_high_bytes"] = (double)arena.high;
    ctx->stats["amg.arena_fallback_bytes"] = (double)g_fallback_bytes;
    ctx->stats["amg.alloc.malloc_calls"] = (double)(g_alloc_stats.malloc_n - a0.malloc_n);
    ctx->stats["amg.alloc.malloc_s"] = g_alloc_stats.malloc_s - a0.malloc_s;
    ctx->stats["amg.alloc.free_calls"] = (double)(g_alloc_stats.free_n - a0.free_n);
    ctx->stats["amg.alloc.free_s"] = g_alloc_stats.free_s - a0.free_s;
    const double t = wall_now();
    arena.destroy();
    ctx->stats["amg.setup.arena release"] = wall_now() - t;
  };
  try {
    build_levels(ctx, H, t0);
    // the hierarchy leaves the arena: one allocation of the exact size, every buffer copied into it
    const double t = wall_now();
    g_arena = g_dbuf_arena = nullptr;
    size_t total = 0;
    auto count = [&](auto &buf) {
      if (buf.p && buf.owner == &arena) total += (buf.bytes() + 511) & ~(size_t)511;
    };
    for (auto *L : H->levels) for_each_buffer(*L, count);
    count(H->coarse_inv);
    if (total) {
      const double tm = wall_now();
      CUDA_CHECK(cudaMalloc(&H->store, total));
      g_alloc_stats.malloc_s += wall_now() - tm;
      g_alloc_stats.malloc_n++;
    }
    H->store_bytes = total;
    size_t off = 0;
    auto move = [&](auto &buf) {
      if (buf.p && buf.owner == &arena) {
        const size_t bytes = (buf.bytes() + 511) & ~(size_t)511;
        buf.rehome((char *)H->store + off, ctx->stream);
        off += bytes;
      }
    };
    for (auto *L : H->levels) for_each_buffer(*L, move);
    move(H->coarse_inv);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (arena.used != 0) NOSH_THROW(NOSH_ESTATE, "internal: AMG set-up left %zu bytes of its arena in use", arena.used);
    ctx->stats["amg.store_bytes"] = (double)total;
    ctx->stats["amg.setup.move to permanent storage"] = wall_now() - t;
  } catch (...) {
    // whatever was built points into the arena: drop it before the arena goes
    amg_free(ctx);
    leave();
    throw;
  }
  leave();
  ctx->amg_valid = true;
  ctx->amg_dinv_version = ctx->keoreg_version;
  H->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

struct Cheb {
  double theta, delta, sigma;
  Cheb() {  // spectrum of S^-1 A, S = absolute row sums: lambda_max <= 1 is a true bound, so no boost factor
    const double lmax = 1.0, lmin = lmax / CHEB_RATIO;
    theta = 0.5 * (lmax + lmin);
    delta = 0.5 * (lmax - lmin);
    sigma = theta / delta;
  }
};

// One V-cycle on level `lev` with right-hand side b (n entries) and zero initial guess.  The result is
// written to `out` if given (level 0), else to one of the level's own buffers; returns where it is.
double2 *vcycle_level(Ctx *ctx, Amg &H, int lev, const double2 *b, double2 *out, const KrylovState *gate) {
  AmgLevel &L = *H.levels[lev];
  const int64_t n = L.n;
  if (lev == (int)H.levels.size() - 1) {
    double2 *x = out ? out : L.x.p;
    const int64_t n2 = H.n_coarse2;
    if (n2 > 0) {
      k_dense_apply<<<(unsigned)cdiv(n2 * 32, 256), 256, 0, ctx->stream>>>(n2, H.coarse_inv.p, (const double *)b,
                                                                           (double *)x, gate);
      ctx->launches++;
      CUDA_CHECK(cudaGetLastError());
    }
    return x;
  }
  const int deg = lev == 0 ? ctx->amg_degree : ctx->amg_coarse_degree;
  const Cheb ch;
  // 2*deg - 1 buffer flips follow the first write: start so that the last one lands in `out`
  double2 *xa = L.x.p, *xb = out ? out : L.x2.p;
  double2 *dvec = deg > 1 ? L.d.p : nullptr;
  // pre-smoothing from x = 0
  ALAUNCH(ctx, k_cheb_first, n, n, b, L.sinv.p, 1.0 / ch.theta, dvec, xa, gate);
  double rho = 1.0 / ch.sigma;
  for (int k = 1; k < deg; k++) {
    const double rho_new = 1.0 / (2.0 * ch.sigma - rho);
    level_apply(ctx, L, lev, M_CHEB, xa, xb, b, dvec, rho_new * rho, 2.0 * rho_new / ch.delta, gate, L.sinv.p);
    std::swap(xa, xb);
    rho = rho_new;
  }
  // coarse-grid correction
  level_apply(ctx, L, lev, M_RESID, xa, L.r.p, b, nullptr, 0.0, 0.0, gate, nullptr);
  AmgLevel &C = *H.levels[lev + 1];
  const bool p32 = lev == 0 && ctx->amg_mixed && L.p_val32.p;
  if (C.n <= SMALL_LEVEL)  // long rows of P^T, few of them: a full warp per coarse node
    k_restrict<32, B22><<<(unsigned)cdiv(C.n * 32, 256), 256, 0, ctx->stream>>>(C.n, L.r_rowptr.p, L.r_fine.p, L.r_pos.p,
                                                                               L.p_val.p, L.r.p, C.b.p, gate);
  else if (p32)
    k_restrict<8, float4><<<(unsigned)cdiv(C.n * 8, 256), 256, 0, ctx->stream>>>(C.n, L.r_rowptr.p, L.r_fine.p, L.r_pos.p,
                                                                                L.p_val32.p, L.r.p, C.b.p, gate);
  else
    k_restrict<8, B22><<<(unsigned)cdiv(C.n * 8, 256), 256, 0, ctx->stream>>>(C.n, L.r_rowptr.p, L.r_fine.p, L.r_pos.p,
                                                                             L.p_val.p, L.r.p, C.b.p, gate);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
  const double2 *xc = vcycle_level(ctx, H, lev + 1, C.b.p, nullptr, gate);
  if (p32 && C.n > SMALL_LEVEL)
    ALAUNCH(ctx, k_prolong<float4>, n, n, L.p_rowptr.p, L.p_col.p, L.p_val32.p, xc, xa, gate);
  else
    ALAUNCH(ctx, k_prolong<B22>, n, n, L.p_rowptr.p, L.p_col.p, L.p_val.p, xc, xa, gate);
  // post-smoothing
  rho = 1.0 / ch.sigma;
  level_apply(ctx, L, lev, M_CHEB, xa, xb, b, dvec, 0.0, 1.0 / ch.theta, gate, L.sinv.p);
  std::swap(xa, xb);
  for (int k = 1; k < deg; k++) {
    const double rho_new = 1.0 / (2.0 * ch.sigma - rho);
    level_apply(ctx, L, lev, M_CHEB, xa, xb, b, dvec, rho_new * rho, 2.0 * rho_new / ch.delta, gate, L.sinv.p);
    std::swap(xa, xb);
    rho = rho_new;
  }
  return xa;
}

}  // namespace

void amg_free(Ctx *ctx) {
  delete ctx->amg;
  ctx->amg = nullptr;
  ctx->amg_valid = false;
}

// the fp32 copies the mixed-precision cycle reads: K whenever it was refilled, the finest P once per hierarchy
static void mixed_refresh(Ctx *ctx) {
  if (!ctx->amg_mixed || ctx->amg->levels.size() < 2) return;
  if (ctx->kval32_version != ctx->kval_version || ctx->Kval32.n < (size_t)ctx->nstored) {
    ctx->Kval32.ensure(ctx->nstored > 0 ? ctx->nstored : 1);
    if (ctx->nstored) ALAUNCH(ctx, k_c64_to_c32, ctx->nstored, ctx->nstored, ctx->Kval.p, ctx->Kval32.p);
    ctx->kval32_version = ctx->kval_version;
  }
  AmgLevel &L = *ctx->amg->levels[0];
  if (!L.p_val32.p && L.p_nnz > 0) {
    L.p_val32.alloc(L.p_nnz);
    ALAUNCH(ctx, k_b22_to_f32, L.p_nnz, L.p_nnz, L.p_val.p, L.p_val32.p);
  }
}

void amg_ensure(Ctx *ctx) {
  if (!ctx->has_mesh) NOSH_THROW(NOSH_ESTATE, "no mesh set");
  if (!ctx->keoreg_ok || !ctx->keo_filled)
    NOSH_THROW(NOSH_ESTATE, "regularised KEO not built (call nosh_keoreg_rebuild first)");
  if (ctx->layout != NOSH_LAYOUT_SELL32) NOSH_THROW(NOSH_EUNSUPPORTED, "the AMG preconditioner needs the SELL-32 layout");
  if (ctx->amg && ctx->amg_valid) {
    if (ctx->amg_dinv_version != ctx->keoreg_version) {
      // "reuse: type" = "full": the hierarchy is kept; only the finest-level matrix (read live from the
      // ctx) and its diagonal follow the new state
      l0_refresh_diag(ctx, *ctx->amg->levels[0]);
      ctx->amg_dinv_version = ctx->keoreg_version;
    }
    mixed_refresh(ctx);
    return;
  }
  build_hierarchy(ctx);
  mixed_refresh(ctx);
}

static void vcycle_launches(Ctx *ctx, const double2 *b, double2 *x, const KrylovState *gate) {
  Amg &H = *ctx->amg;
  if (ctx->nranks > 1 && H.levels.size() > 1) {
    // block-Jacobi over ranks: smooth on buffers whose ghost part is zero, then copy out
    AmgLevel &L = *H.levels[0];
    const double2 *res = vcycle_level(ctx, H, 0, b, L.b.p, gate);
    ALAUNCH(ctx, k_copy2, ctx->No, ctx->No, res, x, gate);
  } else {
    vcycle_level(ctx, H, 0, b, x, gate);
  }
}

// The cycle is ~25 launches, most of them 5-90 us long: issued one by one the launch gaps cost ~0.35 ms of a 2.9 ms
// preconditioned iteration (profiles/r1_launches_amg_minres_n200_iters10.csv: 2.5 ms of kernels).  All arguments
// are fixed for given (b, x, gate) pointers -- the Krylov loops alternate between two pairs -- so every distinct
// cycle is captured once into a CUDA graph and replayed (NOSH_B200_AMG_GRAPH=0 / tuning key "amg_graph" = 0:
// plain launches).  The graphs die with the hierarchy (amg_free).
void amg_vcycle(Ctx *ctx, const double2 *b, double2 *x, const KrylovState *gate) {
  Amg &H = *ctx->amg;
  if (ctx->No == 0) return;
  if (!ctx->amg_graph) {
    vcycle_launches(ctx, b, x, gate);
    return;
  }
  for (auto &g : H.graphs)
    if (g.b == b && g.x == x && g.gate == gate && g.degree == ctx->amg_degree && g.coarse_degree == ctx->amg_coarse_degree &&
        g.mixed == ctx->amg_mixed) {
      CUDA_CHECK(cudaGraphLaunch(g.exec, ctx->stream));
      ctx->launches += g.launches;
      return;
    }
  if (H.graphs.size() >= 8) {  // more pointer pairs than any solver here uses: stop caching, launch directly
    vcycle_launches(ctx, b, x, gate);
    return;
  }
  const int64_t l0 = ctx->launches;
  cudaGraph_t graph = nullptr;
  CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  try {
    vcycle_launches(ctx, b, x, gate);
  } catch (...) {
    cudaStreamEndCapture(ctx->stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  CUDA_CHECK(cudaStreamEndCapture(ctx->stream, &graph));
  VcycleGraph g{b, x, gate, ctx->amg_degree, ctx->amg_coarse_degree, ctx->amg_mixed, nullptr, ctx->launches - l0};
  const cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) NOSH_THROW(NOSH_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  H.graphs.push_back(g);
  CUDA_CHECK(cudaGraphLaunch(g.exec, ctx->stream));
}

}  // namespace nosh
