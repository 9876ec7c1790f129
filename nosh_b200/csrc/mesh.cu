// mesh.cu -- device-side mesh pipeline (rows a1-a4 of SURVEY.md section 8):
//   candidate cells (global ids) -> local cells + ghosts -> unique edges, cell->edge,
//   edge->cell and vertex->cell incidence (radix sorts) -> edge lengths, FVM edge
//   coefficients, control volumes (atomic-free, fixed summation order) -> block-CSR
//   graph of the owned rows + assembly slots -> storage layout (CSR or SELL-32).
// Reference: src/mesh.cpp:629-691,787-881, src/mesh_tetra.cpp:41-271, src/mesh_tri.cpp:44-210.
// The sorts/scans/selects use CUB (one-time set-up, integer work); every geometric and
// structural kernel is hand written.
#include <cub/cub.cuh>

#include "geom.cuh"
#include "mesh.h"

#include "comm.h"

namespace nosh {

namespace {

constexpr uint32_t SENT32 = 0xFFFFFFFFu;
constexpr uint64_t SENT64 = 0xFFFFFFFFFFFFFFFFull;

__device__ __constant__ int c_tet_pair[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
__device__ __constant__ int c_tri_pair[3][2] = {{0, 1}, {0, 2}, {1, 2}};
// Kuhn tetrahedra of the unit cube (corner masks x=1,y=2,z=4), one per axis permutation
__device__ __constant__ int c_kuhn[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7},
                                            {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};

inline dim3 grid_for(int64_t n, int tpb = 256) { return dim3((unsigned)cdiv(n > 0 ? n : 1, tpb)); }

#define LAUNCH(ctx, kernel, n, ...)                                        \
  do {                                                                     \
    kernel<<<grid_for(n), 256, 0, (ctx)->stream>>>(__VA_ARGS__);           \
    (ctx)->launches++;                                                     \
    CUDA_CHECK(cudaGetLastError());                                        \
  } while (0)

// ---- CUB wrappers -----------------------------------------------------------------
struct Temp {
  DBuf<char> buf;
  void *get(size_t bytes) {
    buf.ensure(bytes);
    return buf.p;
  }
};

void sort_pairs_u64(Ctx *ctx, Temp &tmp, DBuf<uint64_t> &keys, DBuf<uint32_t> &vals, int64_t n,
                    int end_bit = 64) {
  // payloads are 32-bit positions: up to 2^32 - 2 items (64M vertices = 2.3e9 (cell, edge) pairs fit)
  if (n >= (int64_t)4294967295ll) NOSH_THROW(NOSH_EINVAL, "mesh too large for one GPU (sort of %lld items)", (long long)n);
  DBuf<uint64_t> k2;
  DBuf<uint32_t> v2;
  k2.alloc(n);
  v2.alloc(n);
  cub::DoubleBuffer<uint64_t> dk(keys.p, k2.p);
  cub::DoubleBuffer<uint32_t> dv(vals.p, v2.p);
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, end_bit, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(t, bytes, dk, dv, n, 0, end_bit, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (dk.Current() != keys.p) keys.swap(k2);
  if (dv.Current() != vals.p) vals.swap(v2);
}

void sort_pairs_u32(Ctx *ctx, Temp &tmp, DBuf<uint32_t> &keys, DBuf<uint32_t> &vals, int64_t n) {
  if (n >= (int64_t)2147483647) NOSH_THROW(NOSH_EINVAL, "mesh too large for one GPU (sort of %lld items)", (long long)n);
  DBuf<uint32_t> k2, v2;
  k2.alloc(n);
  v2.alloc(n);
  cub::DoubleBuffer<uint32_t> dk(keys.p, k2.p);
  cub::DoubleBuffer<uint32_t> dv(vals.p, v2.p);
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n, 0, 32, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(t, bytes, dk, dv, (int)n, 0, 32, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (dk.Current() != keys.p) keys.swap(k2);
  if (dv.Current() != vals.p) vals.swap(v2);
}

void sort_keys_u32(Ctx *ctx, Temp &tmp, DBuf<uint32_t> &keys, int64_t n) {
  DBuf<uint32_t> k2;
  k2.alloc(n);
  cub::DoubleBuffer<uint32_t> dk(keys.p, k2.p);
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, dk, (int)n, 0, 32, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(t, bytes, dk, (int)n, 0, 32, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (dk.Current() != keys.p) keys.swap(k2);
}

void inclusive_scan_i32(Ctx *ctx, Temp &tmp, const int32_t *in, int32_t *out, int64_t n) {
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceScan::InclusiveSum(t, bytes, in, out, n, ctx->stream));
}

void exclusive_scan_i32(Ctx *ctx, Temp &tmp, const int32_t *in, int32_t *out, int64_t n) {
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream));
  void *t = tmp.get(bytes);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, (int)n, ctx->stream));
}

template <typename T>
T fetch(Ctx *ctx, const T *dptr) {
  T h;
  CUDA_CHECK(cudaMemcpyAsync(&h, dptr, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return h;
}

// ---- small device helpers -----------------------------------------------------------
__device__ __forceinline__ int lower_bound_u32(const uint32_t *a, int n, uint32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int to_local(uint32_t g, int64_t vb, int64_t ve, int64_t No,
                                        const uint32_t *ghost, int Ng) {
  if ((int64_t)g >= vb && (int64_t)g < ve) return (int)((int64_t)g - vb);
  return (int)No + lower_bound_u32(ghost, Ng, g);
}

// ---- tetgrid generator ----------------------------------------------------------------
__device__ __forceinline__ double jitter_unit(uint64_t seed, uint64_t gid, int comp) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (gid * 3ull + (uint64_t)(comp + 1));
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const double u = __dmul_rn((double)(z >> 11), 1.0 / 9007199254740992.0);
  return __dadd_rn(__dmul_rn(2.0, u), -1.0);
}

struct GridDesc {
  int n[3];
  double lo[3], h[3], jh[3];
  uint64_t seed;
};

// explicit _rn intrinsics: no FMA contraction, so the coordinates are bit-identical to
// the numpy restatement used by the parity tests (oracle/meshgen.py:tetgrid).
__global__ void k_tetgrid_coords(const int32_t *gid, int64_t Nl, GridDesc g, double *coords) {
  const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (l >= Nl) return;
  const int64_t v = gid[l];
  const int ijk[3] = {(int)(v % g.n[0]), (int)((v / g.n[0]) % g.n[1]),
                      (int)(v / ((int64_t)g.n[0] * g.n[1]))};
  bool interior = true;
#pragma unroll
  for (int d = 0; d < 3; d++) interior = interior && ijk[d] > 0 && ijk[d] < g.n[d] - 1;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double base = __dadd_rn(g.lo[d], __dmul_rn((double)ijk[d], g.h[d]));
    const double disp = __dmul_rn(g.jh[d], jitter_unit(g.seed, (uint64_t)v, d));
    coords[3 * l + d] = interior ? __dadd_rn(base, disp) : base;
  }
}

// hex cells [hex_begin, hex_begin+nhex) (x fastest), 6 Kuhn tets each, global vertex ids
__global__ void k_tetgrid_cells(int64_t hex_begin, int64_t nhex, int nx, int ny, int4 *cells) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nhex * 6) return;
  const int64_t c = hex_begin + t / 6;
  const int k6 = (int)(t % 6);
  const int64_t ci = c % (nx - 1), cj = (c / (nx - 1)) % (ny - 1), ck = c / ((int64_t)(nx - 1) * (ny - 1));
  const int64_t v0 = ci + nx * (cj + (int64_t)ny * ck);
  int v[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int m = c_kuhn[k6][i];
    v[i] = (int)(v0 + (m & 1) + (int64_t)nx * (((m >> 1) & 1) + (int64_t)ny * ((m >> 2) & 1)));
  }
  cells[t] = make_int4(v[0], v[1], v[2], v[3]);
}

// ---- validation of caller-supplied connectivity (the generator's cells are valid by construction) ----
// err: 1 = vertex index outside [0, nv); 2 = a vertex repeated inside one cell.  first: lowest offending cell.
template <int NVC>
__global__ void k_validate_cells(const int32_t *cellsG, int64_t nc, int64_t nv, int *err, unsigned long long *first) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int32_t v[NVC];
  int bad = 0;
#pragma unroll
  for (int i = 0; i < NVC; i++) {
    v[i] = cellsG[c * NVC + i];
    if (v[i] < 0 || (int64_t)v[i] >= nv) bad = 1;
  }
  if (!bad) {
#pragma unroll
    for (int i = 0; i < NVC; i++)
#pragma unroll
      for (int j = i + 1; j < NVC; j++)
        if (v[i] == v[j]) bad = 2;
  }
  if (bad) {
    atomicMax(err, bad == 1 ? 2 : 1);  // an out-of-range index outranks a repeated vertex
    atomicMin(first, (unsigned long long)c);
  }
}

// ---- localisation ----------------------------------------------------------------------
template <int NVC>
__global__ void k_flag_cells(const int32_t *cellsG, int64_t nc, int64_t vb, int64_t ve, int32_t *flag) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int f = 0;
#pragma unroll
  for (int i = 0; i < NVC; i++) {
    const int64_t g = cellsG[c * NVC + i];
    f |= (g >= vb && g < ve);
  }
  flag[c] = f;
}
template <int NVC>
__global__ void k_compact_cells(const int32_t *cellsG, const int32_t *flag, const int32_t *pos,
                                int64_t nc, int32_t *out) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc || !flag[c]) return;
  const int64_t o = pos[c];
#pragma unroll
  for (int i = 0; i < NVC; i++) out[o * NVC + i] = cellsG[c * NVC + i];
}
__global__ void k_ghost_keys(const int32_t *cellsG, int64_t n, int64_t vb, int64_t ve, uint32_t *keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = cellsG[i];
  keys[i] = (g >= vb && g < ve) ? SENT32 : (uint32_t)g;
}
__global__ void k_head_flags_u32(const uint32_t *keys, int64_t n, int32_t *head) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (keys[i] != SENT32 && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}
__global__ void k_scatter_unique_u32(const uint32_t *keys, const int32_t *head, const int32_t *incl,
                                     int64_t n, uint32_t *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !head[i]) return;
  out[incl[i] - 1] = keys[i];
}
__global__ void k_fill_gid(int64_t vb, int64_t No, const uint32_t *ghost, int64_t Ng, int32_t *gid) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < No) gid[i] = (int32_t)(vb + i);
  else if (i < No + Ng) gid[i] = (int32_t)ghost[i - No];
}
__global__ void k_localize(const int32_t *cellsG, int64_t n, int64_t vb, int64_t ve, int64_t No,
                           const uint32_t *ghost, int Ng, int32_t *cellsL) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  cellsL[i] = to_local((uint32_t)cellsG[i], vb, ve, No, ghost, Ng);
}
__global__ void k_gather_coords(const double *gcoords, const int32_t *gid, int64_t Nl, double *coords) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= Nl) return;
  const int64_t g = gid[i];
  coords[3 * i] = gcoords[3 * g];
  coords[3 * i + 1] = gcoords[3 * g + 1];
  coords[3 * i + 2] = gcoords[3 * g + 2];
}
// partitioned ingestion (mesh_from_host_local): the caller's vertex list is sorted by global id once; the
// coordinates of my owned + ghost vertices are found by binary search
struct LocalCoords {
  const uint32_t *sorted_gid;  // nv_local, ascending
  const uint32_t *sorted_idx;  // position in the caller's arrays
  int64_t nv_local;
  const double *coords;        // nv_local x 3, caller's order
};
__global__ void k_gather_coords_local(LocalCoords lc, const int32_t *gid, int64_t Nl, double *coords, int *err) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= Nl) return;
  const uint32_t g = (uint32_t)gid[i];
  int64_t lo = 0, hi = lc.nv_local;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (lc.sorted_gid[mid] < g) lo = mid + 1; else hi = mid;
  }
  if (lo >= lc.nv_local || lc.sorted_gid[lo] != g) {
    atomicMax(err, 3);
    coords[3 * i] = coords[3 * i + 1] = coords[3 * i + 2] = 0.0;
    return;
  }
  const int64_t j = lc.sorted_idx[lo];
  coords[3 * i] = lc.coords[3 * j];
  coords[3 * i + 1] = lc.coords[3 * j + 1];
  coords[3 * i + 2] = lc.coords[3 * j + 2];
}
__global__ void k_gids_to_keys(const int64_t *gids, int64_t n, int64_t n_global, uint32_t *keys, uint32_t *vals,
                               int *err) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = gids[i];
  if (g < 0 || g >= n_global) atomicMax(err, 1);
  keys[i] = (uint32_t)g;
  vals[i] = (uint32_t)i;
}
__global__ void k_check_unique_sorted(const uint32_t *keys, int64_t n, int *err) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i + 1 < n && keys[i] == keys[i + 1]) atomicMax(err, 2);
}
__global__ void k_cells_to_global(const int32_t *cells_local, const int64_t *gids, int64_t n, int32_t *cellsG) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) cellsG[i] = (int32_t)gids[cells_local[i]];
}

// ---- edges --------------------------------------------------------------------------------
// one key per (cell, local edge): (gmin << 32 | gmax) if an endpoint is owned, else sentinel
template <int NVC, int NE>
__global__ void k_edge_keys(const int32_t *cellsL, const int32_t *gid, int64_t nc, int64_t No,
                            uint64_t *keys, uint32_t *vals) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nc * NE) return;
  const int64_t c = t / NE;
  const int le = (int)(t % NE);
  const int a = cellsL[c * NVC + (NE == 6 ? c_tet_pair[le][0] : c_tri_pair[le][0])];
  const int b = cellsL[c * NVC + (NE == 6 ? c_tet_pair[le][1] : c_tri_pair[le][1])];
  uint32_t ga = (uint32_t)gid[a], gb = (uint32_t)gid[b];
  if (ga > gb) { const uint32_t s = ga; ga = gb; gb = s; }
  keys[t] = (a < No || b < No) ? (((uint64_t)ga << 32) | gb) : SENT64;
  vals[t] = (uint32_t)t;
}
__global__ void k_head_flags_u64(const uint64_t *keys, int64_t n, int32_t *head) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (keys[i] != SENT64 && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}
// sorted position i -> edge id (incl[i]-1); writes edges, incidence pointers, cell_edges
__global__ void k_edges_from_sorted(const uint64_t *keys, const uint32_t *vals, const int32_t *head,
                                    const int32_t *incl, int64_t n, int64_t vb, int64_t ve, int64_t No,
                                    const uint32_t *ghost, int Ng, int32_t *edges, uint32_t *inc_ptr) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  if (k == SENT64 || !head[i]) return;
  const int e = incl[i] - 1;
  edges[2 * e] = to_local((uint32_t)(k >> 32), vb, ve, No, ghost, Ng);
  edges[2 * e + 1] = to_local((uint32_t)(k & 0xFFFFFFFFu), vb, ve, No, ghost, Ng);
  inc_ptr[e] = (uint32_t)i;
}
__global__ void k_count_nonsent(const uint64_t *keys, int64_t n, uint32_t *out) {
  out[0] = (uint32_t)lower_bound_u64(keys, n, SENT64);
}
__global__ void k_edge_length(const double *coords, const int32_t *edges, int64_t E, double *len) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  len[e] = norm3(ldv3(coords, edges[2 * e]) - ldv3(coords, edges[2 * e + 1]));
}

// ---- per-cell FVM edge coefficients (a2) ---------------------------------------------------
// One thread per cell.  The cell's edges are ordered by their global (gmin,gmax) key --
// the order of the reference's handle-sorted Range (src/mesh.cpp:636-666) -- before the
// volume / LU steps, so pivoting sees the matrix the reference sees.
__global__ void k_cell_coeff_tet(const double *coords, const int32_t *cellsL, const int32_t *gid,
                                 int64_t nc, double *coef /* nc x 6 */, int *err) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int v[4];
  V3 x[4];
  uint32_t g[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[i] = cellsL[4 * c + i];
    x[i] = ldv3(coords, v[i]);
    g[i] = (uint32_t)gid[v[i]];
  }
  uint64_t key[6];
  int ord[6];
#pragma unroll
  for (int le = 0; le < 6; le++) {
    uint32_t ga = g[c_tet_pair[le][0]], gb = g[c_tet_pair[le][1]];
    if (ga > gb) { const uint32_t s = ga; ga = gb; gb = s; }
    key[le] = ((uint64_t)ga << 32) | gb;
    ord[le] = le;
  }
  // insertion sort of 6 (key, le)
  for (int i = 1; i < 6; i++) {
    const uint64_t k = key[i];
    const int o = ord[i];
    int j = i - 1;
    while (j >= 0 && key[j] > k) {
      key[j + 1] = key[j];
      ord[j + 1] = ord[j];
      j--;
    }
    key[j + 1] = k;
    ord[j + 1] = o;
  }
  V3 e[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const int le = ord[i];
    const int a = c_tet_pair[le][0], b = c_tet_pair[le][1];
    // x_{v0} - x_{v1}, v0 = endpoint with the smaller global id (src/mesh_tetra.cpp:65-67)
    e[i] = (g[a] < g[b]) ? (x[a] - x[b]) : (x[b] - x[a]);
  }
  // volume from three non-coplanar edges, with the retry of src/mesh_tetra.cpp:122-128
  double al = dot3(e[0], cross3(e[1], e[2]));
  if (fabs(al) / norm3(e[0]) / norm3(e[1]) / norm3(e[2]) < 1.0e-5) {
    al = dot3(e[0], cross3(e[1], e[3]));
    if (fabs(al) / norm3(e[0]) / norm3(e[1]) / norm3(e[3]) < 1.0e-5) atomicExch(err, 1);
  }
  const double vol = fabs(al) / 6.0;
  double cf[6];
  if (!edge_coefficients<6>(e, vol, cf)) atomicExch(err, 1);
#pragma unroll
  for (int i = 0; i < 6; i++) coef[6 * c + ord[i]] = cf[i];
}

__global__ void k_cell_coeff_tri(const double *coords, const int32_t *cellsL, const int32_t *gid,
                                 int64_t nc, double *coef /* nc x 3 */, int *err) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  V3 x[3];
  uint32_t g[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int v = cellsL[3 * c + i];
    x[i] = ldv3(coords, v);
    g[i] = (uint32_t)gid[v];
  }
  uint64_t key[3];
  int ord[3];
#pragma unroll
  for (int le = 0; le < 3; le++) {
    uint32_t ga = g[c_tri_pair[le][0]], gb = g[c_tri_pair[le][1]];
    if (ga > gb) { const uint32_t s = ga; ga = gb; gb = s; }
    key[le] = ((uint64_t)ga << 32) | gb;
    ord[le] = le;
  }
  for (int i = 1; i < 3; i++) {
    const uint64_t k = key[i];
    const int o = ord[i];
    int j = i - 1;
    while (j >= 0 && key[j] > k) {
      key[j + 1] = key[j];
      ord[j + 1] = ord[j];
      j--;
    }
    key[j + 1] = k;
    ord[j + 1] = o;
  }
  V3 e[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int a = c_tri_pair[ord[i]][0], b = c_tri_pair[ord[i]][1];
    e[i] = (g[a] < g[b]) ? (x[a] - x[b]) : (x[b] - x[a]);
  }
  const double vol = 0.5 * norm3(cross3(e[0], e[1]));  // src/mesh_tri.cpp:121
  double cf[3];
  if (!edge_coefficients<3>(e, vol, cf)) atomicExch(err, 1);
#pragma unroll
  for (int i = 0; i < 3; i++) coef[3 * c + ord[i]] = cf[i];
}

// covolume_e = sum over the edge's cells, in ascending cell order, of coef * length
// (src/mesh_tetra.cpp:95-99); products and sums individually rounded like the scalar loop.
__global__ void k_edge_covolume(const double *coef, const uint32_t *inc, const uint32_t *inc_ptr,
                                const double *len, int64_t E, double *cov) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double l = len[e];
  double s = 0.0;
  for (uint32_t p = inc_ptr[e]; p < inc_ptr[e + 1]; p++) s = __dadd_rn(s, __dmul_rn(coef[inc[p]], l));
  cov[e] = s;
}

// ---- control volumes (a3) -------------------------------------------------------------------
__global__ void k_cell_cv_tet(const double *coords, const int32_t *cellsL, int64_t nc,
                              double *contrib /* nc x 4 */, int *err) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  V3 x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = ldv3(coords, cellsL[4 * c + i]);
  bool ok = true;
  const V3 cc = tet_circumcenter(x, ok);
  double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int le = 0; le < 6; le++) {
    const int e0 = c_tet_pair[le][0], e1 = c_tet_pair[le][1];
    int o0 = -1, o1 = -1;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (i != e0 && i != e1) { if (o0 < 0) o0 = i; else o1 = i; }
    const double len = norm3(x[e1] - x[e0]);
    const double cov = covolume3d(cc, x[e0], x[e1], x[o0], x[o1], ok);
    const double pyr = 0.5 * len * cov / 3;  // src/mesh_tetra.cpp:263
    s[e0] += pyr;
    s[e1] += pyr;
  }
  if (!ok) atomicExch(err, 2);
#pragma unroll
  for (int i = 0; i < 4; i++) contrib[4 * c + i] = s[i];
}
__global__ void k_cell_cv_tri(const double *coords, const int32_t *cellsL, int64_t nc,
                              double *contrib /* nc x 3 */, int *err) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= nc) return;
  V3 x[3];
#pragma unroll
  for (int i = 0; i < 3; i++) x[i] = ldv3(coords, cellsL[3 * c + i]);
  bool ok = true;
  const V3 cc = tri_circumcenter(x[0], x[1], x[2], ok);
  double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int le = 0; le < 3; le++) {
    const int e0 = c_tri_pair[le][0], e1 = c_tri_pair[le][1], o = 3 - e0 - e1;
    const double len = norm3(x[e1] - x[e0]);
    const double cov = covolume2d(cc, x[e0], x[e1], x[o]);
    const double pyr = 0.5 * len * cov / 2;  // src/mesh.cpp:971
    s[e0] += pyr;
    s[e1] += pyr;
  }
  if (!ok) atomicExch(err, 2);
#pragma unroll
  for (int i = 0; i < 3; i++) contrib[3 * c + i] = s[i];
}
__global__ void k_vertex_keys(const int32_t *cellsL, int64_t n, int64_t No, uint32_t *keys, uint32_t *vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = cellsL[i];
  keys[i] = v < No ? (uint32_t)v : SENT32;
  vals[i] = (uint32_t)i;
}
// cv[v] = sum of the contributions of v's cells in ascending cell order
__global__ void k_vertex_gather(const uint32_t *keys, const uint32_t *vals, int64_t n,
                                const double *contrib, int64_t No, double *cv) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= No) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (uint32_t)v) lo = mid + 1; else hi = mid;
  }
  double s = 0.0;
  for (int64_t p = lo; p < n && keys[p] == (uint32_t)v; p++) s += contrib[vals[p]];
  cv[v] = s;
}

// ---- block graph (a4) -------------------------------------------------------------------------
// entries: 2 per edge + 1 diagonal per owned row; key = row << 32 | global col id
__global__ void k_graph_keys(const int32_t *edges, const int32_t *gid, int64_t E, int64_t No,
                             uint64_t *keys, uint32_t *vals) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < 2 * E) {
    const int64_t e = t >> 1;
    const int d = (int)(t & 1);
    const int r = edges[2 * e + d], cidx = edges[2 * e + 1 - d];
    keys[t] = r < No ? (((uint64_t)r << 32) | (uint32_t)gid[cidx]) : SENT64;
    vals[t] = (uint32_t)t;  // 2e + d
  } else if (t < 2 * E + No) {
    const int64_t r = t - 2 * E;
    keys[t] = ((uint64_t)r << 32) | (uint32_t)gid[r];
    vals[t] = SENT32;
  }
}
__global__ void k_rowptr(const uint64_t *keys, int64_t n, int64_t No, int32_t *rowptr) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > No) return;
  rowptr[r] = (int32_t)lower_bound_u64(keys, n, (uint64_t)r << 32);
}
__global__ void k_graph_fill(const uint64_t *keys, const uint32_t *vals, int64_t nb, int64_t vb,
                             int64_t ve, int64_t No, const uint32_t *ghost, int Ng, int32_t *csr_col,
                             int32_t *edge_of, int32_t *slot_ij, int32_t *slot_ji, int32_t *diag_slot) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nb) return;
  const uint64_t k = keys[p];
  const int row = (int)(k >> 32);
  csr_col[p] = to_local((uint32_t)(k & 0xFFFFFFFFu), vb, ve, No, ghost, Ng);
  const uint32_t v = vals[p];
  if (v == SENT32) {
    diag_slot[row] = (int32_t)p;
    edge_of[p] = -1;
  } else {
    const int e = (int)(v >> 1);
    edge_of[p] = e;
    if (v & 1) slot_ji[e] = (int32_t)p; else slot_ij[e] = (int32_t)p;
  }
}

// ---- storage layout ---------------------------------------------------------------------------
__global__ void k_slice_width(const int32_t *rowptr, int64_t No, int64_t nslices, int32_t *width32) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  int w = 0;
  for (int64_t r = s * 32; r < s * 32 + 32 && r < No; r++) w = max(w, rowptr[r + 1] - rowptr[r]);
  width32[s] = w * 32;
}
// SELL-32-sigma (sigma = CHUNK = 512): inside every window of 512 consecutive rows -- one CTA of the apply
// kernels, one level-1 chunk of the reductions -- the rows are stored in order of decreasing length (stable),
// so that the 32 rows of a slice have (nearly) equal length and the padding of meshes with strongly varying
// vertex valence disappears.  sell_row[q] = row stored at position q (>= No: a padding lane),
// sell_pos[r] = position of row r.  The permutation depends on row lengths and on the window only, both
// properties of the GLOBAL numbering (partitions are chunk aligned), so it is the same for any number of GPUs.
__global__ void __launch_bounds__(CHUNK) k_sell_sort_window(const int32_t *rowptr, int64_t No, int32_t *sell_row,
                                                           int32_t *sell_pos) {
  __shared__ int len[CHUNK];
  const int t = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * CHUNK + t;
  const int mine = i < No ? rowptr[i + 1] - rowptr[i] : -1;
  len[t] = mine;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < CHUNK; j++) rank += (len[j] > mine) || (len[j] == mine && j < t);
  sell_row[(int64_t)blockIdx.x * CHUNK + rank] = (int32_t)min(i, (int64_t)2147483647);
  if (i < No) sell_pos[i] = (int32_t)((int64_t)blockIdx.x * CHUNK + rank);
}
__global__ void k_slice_width_perm(const int32_t *rowptr, const int32_t *sell_row, int64_t No, int64_t nslices,
                                   int32_t *width32) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  int w = 0;
  for (int l = 0; l < 32; l++) {
    const int64_t r = sell_row[s * 32 + l];
    if (r < No) w = max(w, rowptr[r + 1] - rowptr[r]);
  }
  width32[s] = w * 32;
}
__global__ void k_sell_init(int64_t nstored, int64_t No, const int32_t *slice_off, int64_t nslices,
                            const int32_t *sell_row, int32_t *col) {
  // padding entries point at the row itself (value 0): always a valid, cached address
  const int64_t s = blockIdx.x;
  if (s >= nslices) return;
  const int base = slice_off[s], end = slice_off[s + 1];
  for (int p = base + threadIdx.x; p < end; p += blockDim.x) {
    const int64_t q = s * 32 + ((p - base) & 31);
    const int64_t r = sell_row ? (int64_t)sell_row[q] : q;
    col[p] = (int32_t)(r < No ? r : 0);
  }
}
__global__ void k_sell_pos(const int32_t *rowptr, const int32_t *csr_col, const int32_t *slice_off,
                           const int32_t *sell_pos, int64_t No, int32_t *csr_pos, int32_t *col) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= No) return;
  const int64_t q = sell_pos ? (int64_t)sell_pos[r] : r;
  const int base = slice_off[q >> 5] + (int)(q & 31);
  for (int p = rowptr[r], k = 0; p < rowptr[r + 1]; p++, k++) {
    const int q = base + 32 * k;
    csr_pos[p] = q;
    col[q] = csr_col[p];
  }
}
__global__ void k_iota(int32_t *a, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}
__global__ void k_remap(int32_t *slots, int64_t n, const int32_t *csr_pos) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && slots[i] >= 0) slots[i] = csr_pos[slots[i]];
}

template <int NVC, int NE>
void build_from_cells(Ctx *ctx, DBuf<int32_t> &cellsG, int64_t ncand, const double *gcoords_dev,
                      const GridDesc *grid, const LocalCoords *lc = nullptr) {
  Temp tmp;
  const int64_t vb = ctx->vb, ve = ctx->ve, No = ve - vb;
  ctx->No = No;
  // ---- 1. cells touching an owned vertex --------------------------------------------
  int64_t nc = ncand;
  if (ctx->nranks > 1) {
    DBuf<int32_t> flag, pos;
    flag.alloc(ncand);
    pos.alloc(ncand);
    LAUNCH(ctx, (k_flag_cells<NVC>), ncand, cellsG.p, ncand, vb, ve, flag.p);
    exclusive_scan_i32(ctx, tmp, flag.p, pos.p, ncand);
    nc = ncand ? fetch(ctx, pos.p + ncand - 1) + fetch(ctx, flag.p + ncand - 1) : 0;
    DBuf<int32_t> sel;
    sel.alloc(nc * NVC);
    LAUNCH(ctx, (k_compact_cells<NVC>), ncand, cellsG.p, flag.p, pos.p, ncand, sel.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    cellsG.swap(sel);
  }
  ctx->nc = nc;
  // ---- 2. ghosts -------------------------------------------------------------------
  DBuf<uint32_t> ghost;
  int64_t Ng = 0;
  if (ctx->nranks > 1 && nc > 0) {
    DBuf<uint32_t> keys;
    keys.alloc(nc * NVC);
    LAUNCH(ctx, k_ghost_keys, nc * NVC, cellsG.p, nc * NVC, vb, ve, keys.p);
    sort_keys_u32(ctx, tmp, keys, nc * NVC);
    DBuf<int32_t> head, incl;
    head.alloc(nc * NVC);
    incl.alloc(nc * NVC);
    LAUNCH(ctx, k_head_flags_u32, nc * NVC, keys.p, nc * NVC, head.p);
    inclusive_scan_i32(ctx, tmp, head.p, incl.p, nc * NVC);
    Ng = fetch(ctx, incl.p + nc * NVC - 1);
    ghost.alloc(Ng);
    LAUNCH(ctx, k_scatter_unique_u32, nc * NVC, keys.p, head.p, incl.p, nc * NVC, ghost.p);
  } else {
    ghost.alloc(1);
  }
  ctx->Ng = Ng;
  ctx->Nl = No + Ng;
  const int64_t Nl = ctx->Nl;
  ctx->gid.alloc(Nl);
  LAUNCH(ctx, k_fill_gid, Nl, vb, No, ghost.p, Ng, ctx->gid.p);
  // ---- 3. local cells + coordinates ---------------------------------------------------
  ctx->cells.alloc(nc * NVC);
  LAUNCH(ctx, k_localize, nc * NVC, cellsG.p, nc * NVC, vb, ve, No, ghost.p, (int)Ng, ctx->cells.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cellsG.release();
  ctx->coords.alloc(Nl * 3);
  if (grid)
    LAUNCH(ctx, k_tetgrid_coords, Nl, ctx->gid.p, Nl, *grid, ctx->coords.p);
  else if (lc) {
    DBuf<int> cerr;
    cerr.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(cerr.p, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_gather_coords_local, Nl, *lc, ctx->gid.p, Nl, ctx->coords.p, cerr.p);
    if (fetch(ctx, cerr.p)) NOSH_THROW(NOSH_EMESH, "internal: a vertex of a local cell has no coordinates");
  } else
    LAUNCH(ctx, k_gather_coords, Nl, gcoords_dev, ctx->gid.p, Nl, ctx->coords.p);
  // ---- 4. unique edges, cell->edge, edge->cell incidence -------------------------------
  const int64_t nke = nc * NE;
  DBuf<uint64_t> ekeys;
  DBuf<uint32_t> evals;
  ekeys.alloc(nke);
  evals.alloc(nke);
  LAUNCH(ctx, (k_edge_keys<NVC, NE>), nke, ctx->cells.p, ctx->gid.p, nc, No, ekeys.p, evals.p);
  sort_pairs_u64(ctx, tmp, ekeys, evals, nke);
  int64_t E = 0;
  DBuf<uint32_t> inc_ptr;
  {
    DBuf<int32_t> head, incl;
    head.alloc(nke);
    incl.alloc(nke);
    LAUNCH(ctx, k_head_flags_u64, nke, ekeys.p, nke, head.p);
    inclusive_scan_i32(ctx, tmp, head.p, incl.p, nke);
    E = nke ? fetch(ctx, incl.p + nke - 1) : 0;
    ctx->E = E;
    ctx->edges.alloc(E * 2);
    inc_ptr.alloc(E + 1);
    LAUNCH(ctx, k_edges_from_sorted, nke, ekeys.p, evals.p, head.p, incl.p, nke, vb, ve, No, ghost.p,
           (int)Ng, ctx->edges.p, inc_ptr.p);
    // end pointer: first sentinel position (== number of non-sentinel keys)
    k_count_nonsent<<<1, 1, 0, ctx->stream>>>(ekeys.p, nke, inc_ptr.p + E);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  ekeys.release();
  // ---- 5. edge data -----------------------------------------------------------------------
  ctx->elen.alloc(E);
  ctx->ecov.alloc(E);
  LAUNCH(ctx, k_edge_length, E, ctx->coords.p, ctx->edges.p, E, ctx->elen.p);
  DBuf<int> derr;
  derr.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(derr.p, 0, sizeof(int), ctx->stream));
  {
    DBuf<double> coef;
    coef.alloc(nke);
    if (NE == 6)
      LAUNCH(ctx, k_cell_coeff_tet, nc, ctx->coords.p, ctx->cells.p, ctx->gid.p, nc, coef.p, derr.p);
    else
      LAUNCH(ctx, k_cell_coeff_tri, nc, ctx->coords.p, ctx->cells.p, ctx->gid.p, nc, coef.p, derr.p);
    LAUNCH(ctx, k_edge_covolume, E, coef.p, evals.p, inc_ptr.p, ctx->elen.p, E, ctx->ecov.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  evals.release();
  inc_ptr.release();
  // ---- 6. control volumes -------------------------------------------------------------------
  ctx->cv.alloc(No);
  {
    DBuf<double> contrib;
    contrib.alloc(nc * NVC);
    if (NVC == 4)
      LAUNCH(ctx, k_cell_cv_tet, nc, ctx->coords.p, ctx->cells.p, nc, contrib.p, derr.p);
    else
      LAUNCH(ctx, k_cell_cv_tri, nc, ctx->coords.p, ctx->cells.p, nc, contrib.p, derr.p);
    DBuf<uint32_t> vkeys, vvals;
    vkeys.alloc(nc * NVC);
    vvals.alloc(nc * NVC);
    LAUNCH(ctx, k_vertex_keys, nc * NVC, ctx->cells.p, nc * NVC, No, vkeys.p, vvals.p);
    sort_pairs_u32(ctx, tmp, vkeys, vvals, nc * NVC);
    LAUNCH(ctx, k_vertex_gather, No, vkeys.p, vvals.p, nc * NVC, contrib.p, No, ctx->cv.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  const int herr = fetch(ctx, derr.p);
  if (herr == 1) NOSH_THROW(NOSH_EMESH, "Illegal mesh: tetrahedron too flat (src/mesh_tetra.cpp:361)");
  if (herr == 2) NOSH_THROW(NOSH_EMESH, "Illegal mesh: degenerate cell (src/mesh_tetra.cpp:393)");
  // ---- 7. block graph of the owned rows + assembly slots -------------------------------------
  const int64_t nent = 2 * E + No;
  DBuf<uint64_t> gkeys;
  DBuf<uint32_t> gvals;
  gkeys.alloc(nent);
  gvals.alloc(nent);
  LAUNCH(ctx, k_graph_keys, nent, ctx->edges.p, ctx->gid.p, E, No, gkeys.p, gvals.p);
  sort_pairs_u64(ctx, tmp, gkeys, gvals, nent);
  ctx->rowptr.alloc(No + 1);
  LAUNCH(ctx, k_rowptr, No + 1, gkeys.p, nent, No, ctx->rowptr.p);
  const int64_t nb = fetch(ctx, ctx->rowptr.p + No);
  ctx->nb = nb;
  ctx->csr_col.alloc(nb);
  ctx->edge_of.alloc(nb);
  ctx->slot_ij.alloc(E);
  ctx->slot_ji.alloc(E);
  ctx->diag_slot.alloc(No);
  CUDA_CHECK(cudaMemsetAsync(ctx->slot_ij.p, 0xFF, sizeof(int32_t) * (E ? E : 1), ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ctx->slot_ji.p, 0xFF, sizeof(int32_t) * (E ? E : 1), ctx->stream));
  LAUNCH(ctx, k_graph_fill, nb, gkeys.p, gvals.p, nb, vb, ve, No, ghost.p, (int)Ng, ctx->csr_col.p,
         ctx->edge_of.p, ctx->slot_ij.p, ctx->slot_ji.p, ctx->diag_slot.p);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  gkeys.release();
  gvals.release();
  // ---- 8. storage layout ---------------------------------------------------------------------
  ctx->csr_pos.alloc(nb);
  if (ctx->layout == NOSH_LAYOUT_SELL32) {
    const int64_t ns = cdiv(No, 32);
    ctx->nslices = ns;
    DBuf<int32_t> w32;
    w32.alloc(ns + 1);
    ctx->slice_off.alloc(ns + 1);
    ctx->sell_row.release();
    ctx->sell_pos.release();
    ctx->sell_permuted = false;
    // identity order first; if the padding exceeds 5 % (or sigma sorting is forced) sort the rows of every
    // 512-row window by length.  The decision must not depend on the partition: with several ranks it is taken
    // from the tuning value only (auto = on), one GPU measures its own padding.
    for (int pass = 0; pass < 2; pass++) {
      const bool sorted = pass == 1;
      if (sorted) {
        const int64_t nwin = cdiv(No, CHUNK);
        ctx->sell_row.alloc(nwin * CHUNK);
        ctx->sell_pos.alloc(No);
        if (nwin > 0) {
          k_sell_sort_window<<<(unsigned)nwin, CHUNK, 0, ctx->stream>>>(ctx->rowptr.p, No, ctx->sell_row.p,
                                                                        ctx->sell_pos.p);
          ctx->launches++;
          CUDA_CHECK(cudaGetLastError());
        }
        ctx->sell_permuted = true;
      }
      CUDA_CHECK(cudaMemsetAsync(w32.p, 0, sizeof(int32_t) * (ns + 1), ctx->stream));
      if (sorted)
        LAUNCH(ctx, k_slice_width_perm, ns, ctx->rowptr.p, ctx->sell_row.p, No, ns, w32.p);
      else
        LAUNCH(ctx, k_slice_width, ns, ctx->rowptr.p, No, ns, w32.p);
      exclusive_scan_i32(ctx, tmp, w32.p, ctx->slice_off.p, ns + 1);
      ctx->nstored = fetch(ctx, ctx->slice_off.p + ns);
      if (sorted) break;
      ctx->stats["sell.stored_over_blocks_unsorted"] = nb > 0 ? (double)ctx->nstored / (double)nb : 1.0;
      bool want;
      if (ctx->sell_sigma == 0) {
        want = false;
      } else if (ctx->sell_sigma == 1) {
        want = true;
      } else {
        // auto: the padding of the WHOLE mesh decides (slices never straddle a rank boundary, so the sums over
        // ranks equal the one-GPU numbers and every rank count takes the same decision -- the layout fixes the
        // summation order inside a chunk, which the partition-independent bits rest on)
        int64_t mine[2] = {ctx->nstored, nb}, tot[2] = {ctx->nstored, nb};
        if (ctx->nranks > 1) {
          std::vector<int64_t> all(2 * (size_t)ctx->nranks);
          exchange_allgather(ctx, mine, all.data(), sizeof(mine));
          tot[0] = tot[1] = 0;
          for (int r = 0; r < ctx->nranks; r++) {
            tot[0] += all[2 * r];
            tot[1] += all[2 * r + 1];
          }
        }
        want = tot[0] > tot[1] + tot[1] / 20;
      }
      if (!want) break;
    }
    ctx->stats["sell.stored_over_blocks"] = nb > 0 ? (double)ctx->nstored / (double)nb : 1.0;
    ctx->stats["sell.sigma"] = ctx->sell_permuted ? (double)CHUNK : 0.0;
    ctx->col.alloc(ctx->nstored);
    if (ns > 0) {
      k_sell_init<<<(unsigned)ns, 128, 0, ctx->stream>>>(ctx->nstored, No, ctx->slice_off.p, ns,
                                                         ctx->sell_permuted ? ctx->sell_row.p : nullptr, ctx->col.p);
      ctx->launches++;
      CUDA_CHECK(cudaGetLastError());
    }
    LAUNCH(ctx, k_sell_pos, No, ctx->rowptr.p, ctx->csr_col.p, ctx->slice_off.p,
           ctx->sell_permuted ? ctx->sell_pos.p : nullptr, No, ctx->csr_pos.p, ctx->col.p);
    LAUNCH(ctx, k_remap, E, ctx->slot_ij.p, E, ctx->csr_pos.p);
    LAUNCH(ctx, k_remap, E, ctx->slot_ji.p, E, ctx->csr_pos.p);
    LAUNCH(ctx, k_remap, No, ctx->diag_slot.p, No, ctx->csr_pos.p);
  } else {
    ctx->nslices = 0;
    ctx->sell_row.release();
    ctx->sell_pos.release();
    ctx->sell_permuted = false;
    ctx->nstored = nb;
    LAUNCH(ctx, k_iota, nb, ctx->csr_pos.p, nb);
    ctx->col.alloc(nb);
    CUDA_CHECK(cudaMemcpyAsync(ctx->col.p, ctx->csr_col.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToDevice,
                               ctx->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->has_mesh = true;
}

}  // namespace

// ---- partition of the global vertex range into contiguous, group-aligned pieces -------------
void partition_range(int64_t n_global, int nranks, int rank, int64_t group, int64_t *begin, int64_t *end,
                     int64_t *group_used, int64_t *ngroups_out, int64_t *gbegin, int64_t *gcount) {
  int64_t G = group;
  while (cdiv(n_global, G) > MAX_GROUPS) G *= 2;
  const int64_t ngroups = cdiv(n_global, G);
  const int64_t g0 = ngroups * rank / nranks, g1 = ngroups * (rank + 1) / nranks;
  if (begin) *begin = std::min(n_global, g0 * G);
  if (end) *end = std::min(n_global, g1 * G);
  if (group_used) *group_used = G;
  if (ngroups_out) *ngroups_out = ngroups;
  if (gbegin) *gbegin = g0;
  if (gcount) *gcount = g1 - g0;
}

void setup_partition(Ctx *ctx, int64_t n_global) {
  if (n_global >= (int64_t)2147483647) NOSH_THROW(NOSH_EINVAL, "n_vertices must be < 2^31");
  ctx->n_global = n_global;
  ctx->part_begin.assign(ctx->nranks + 1, 0);
  int64_t G = 0;
  for (int r = 0; r < ctx->nranks; r++) {
    int64_t b, e;
    partition_range(n_global, ctx->nranks, r, ctx->group_vertices, &b, &e, &G, nullptr, nullptr, nullptr);
    ctx->part_begin[r] = b;
    ctx->part_begin[r + 1] = e;
  }
  partition_range(n_global, ctx->nranks, ctx->rank, ctx->group_vertices, &ctx->vb, &ctx->ve, &G,
                  &ctx->n_groups_global, &ctx->group_begin, &ctx->n_groups_local);
  ctx->group_vertices = G;
  ctx->chunks_per_group = (int)(G / CHUNK);
}

void mesh_from_host(Ctx *ctx, int dim, int64_t nv, const double *coords, int64_t ncells,
                    const int32_t *cells) {
  if (dim != 2 && dim != 3) NOSH_THROW(NOSH_EINVAL, "dim must be 2 or 3");
  if (nv <= 0 || ncells <= 0 || !coords || !cells) NOSH_THROW(NOSH_EINVAL, "empty mesh");
  ctx->dim = dim;
  setup_partition(ctx, nv);
  const int nvc = dim + 1;
  DBuf<int32_t> cellsG;
  cellsG.alloc(ncells * nvc);
  CUDA_CHECK(cudaMemcpyAsync(cellsG.p, cells, sizeof(int32_t) * ncells * nvc, cudaMemcpyHostToDevice,
                             ctx->stream));
  // the connectivity comes from the caller: every later kernel indexes with it, so check it first
  // (negative / too large indices would read and write out of bounds on the device)
  {
    DBuf<int> verr;
    DBuf<unsigned long long> vfirst;
    verr.alloc(1);
    vfirst.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(verr.p, 0, sizeof(int), ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(vfirst.p, 0xFF, sizeof(unsigned long long), ctx->stream));
    if (dim == 3)
      LAUNCH(ctx, (k_validate_cells<4>), ncells, cellsG.p, ncells, nv, verr.p, vfirst.p);
    else
      LAUNCH(ctx, (k_validate_cells<3>), ncells, cellsG.p, ncells, nv, verr.p, vfirst.p);
    const int e = fetch(ctx, verr.p);
    if (e) {
      const unsigned long long c = fetch(ctx, vfirst.p);
      if (e == 2)
        NOSH_THROW(NOSH_EMESH, "Illegal mesh: cell %llu references a vertex outside [0, %lld)", c, (long long)nv);
      NOSH_THROW(NOSH_EMESH, "Illegal mesh: cell %llu references the same vertex twice", c);
    }
  }
  DBuf<double> gc;
  gc.alloc(nv * 3);
  CUDA_CHECK(cudaMemcpyAsync(gc.p, coords, sizeof(double) * nv * 3, cudaMemcpyHostToDevice, ctx->stream));
  if (dim == 3)
    build_from_cells<4, 6>(ctx, cellsG, ncells, gc.p, nullptr);
  else
    build_from_cells<3, 3>(ctx, cellsG, ncells, gc.p, nullptr);
}

// Partitioned ingestion (the READ_PART analogue of src/mesh_reader.cpp:32-35): every rank passes only ITS part --
// the cells that touch a vertex of its owned range [nosh_partition_range) (more cells are allowed and dropped),
// the vertices those cells use with their global ids and coordinates, cells indexing into that local list.
void mesh_from_host_local(Ctx *ctx, int dim, int64_t n_global, int64_t nv_local, const int64_t *gids,
                          const double *coords, int64_t ncells, const int32_t *cells) {
  if (dim != 2 && dim != 3) NOSH_THROW(NOSH_EINVAL, "dim must be 2 or 3");
  if (n_global <= 0 || nv_local < 0 || ncells < 0) NOSH_THROW(NOSH_EINVAL, "negative size");
  if ((nv_local > 0 && (!gids || !coords)) || (ncells > 0 && !cells)) NOSH_THROW(NOSH_EINVAL, "NULL array");
  ctx->dim = dim;
  setup_partition(ctx, n_global);
  const int nvc = dim + 1;
  Temp tmp;
  DBuf<int64_t> dg;
  DBuf<uint32_t> keys, vals;
  DBuf<double> dc;
  DBuf<int> derr;
  dg.alloc(nv_local);
  keys.alloc(nv_local);
  vals.alloc(nv_local);
  dc.alloc(nv_local * 3);
  derr.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(derr.p, 0, sizeof(int), ctx->stream));
  if (nv_local) {
    CUDA_CHECK(cudaMemcpyAsync(dg.p, gids, sizeof(int64_t) * nv_local, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(dc.p, coords, sizeof(double) * nv_local * 3, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gids_to_keys, nv_local, dg.p, nv_local, n_global, keys.p, vals.p, derr.p);
    if (fetch(ctx, derr.p)) NOSH_THROW(NOSH_EMESH, "Illegal mesh: a local vertex has a global id outside [0, %lld)", (long long)n_global);
    sort_pairs_u32(ctx, tmp, keys, vals, nv_local);
    LAUNCH(ctx, k_check_unique_sorted, nv_local, keys.p, nv_local, derr.p);
    if (fetch(ctx, derr.p)) NOSH_THROW(NOSH_EMESH, "Illegal mesh: a global vertex id appears twice in the local vertex list");
  }
  DBuf<int32_t> cl, cellsG;
  cl.alloc(ncells * nvc);
  cellsG.alloc(ncells * nvc);
  if (ncells) {
    CUDA_CHECK(cudaMemcpyAsync(cl.p, cells, sizeof(int32_t) * ncells * nvc, cudaMemcpyHostToDevice, ctx->stream));
    DBuf<unsigned long long> vfirst;
    vfirst.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(vfirst.p, 0xFF, sizeof(unsigned long long), ctx->stream));
    if (dim == 3)
      LAUNCH(ctx, (k_validate_cells<4>), ncells, cl.p, ncells, nv_local, derr.p, vfirst.p);
    else
      LAUNCH(ctx, (k_validate_cells<3>), ncells, cl.p, ncells, nv_local, derr.p, vfirst.p);
    const int e = fetch(ctx, derr.p);
    if (e) {
      const unsigned long long c = fetch(ctx, vfirst.p);
      if (e == 2)
        NOSH_THROW(NOSH_EMESH, "Illegal mesh: cell %llu references a vertex outside the local list [0, %lld)", c,
                   (long long)nv_local);
      NOSH_THROW(NOSH_EMESH, "Illegal mesh: cell %llu references the same vertex twice", c);
    }
    LAUNCH(ctx, k_cells_to_global, ncells * nvc, cl.p, dg.p, ncells * nvc, cellsG.p);
  }
  cl.release();
  LocalCoords lc{keys.p, vals.p, nv_local, dc.p};
  if (dim == 3)
    build_from_cells<4, 6>(ctx, cellsG, ncells, nullptr, nullptr, &lc);
  else
    build_from_cells<3, 3>(ctx, cellsG, ncells, nullptr, nullptr, &lc);
}

void mesh_tetgrid(Ctx *ctx, int nx, int ny, int nz, const double lo[3], const double hi[3],
                  double jitter, uint64_t seed) {
  if (nx < 2 || ny < 2 || nz < 2) NOSH_THROW(NOSH_EINVAL, "tetgrid needs at least 2 vertices per axis");
  ctx->dim = 3;
  const int64_t plane = (int64_t)nx * ny;
  setup_partition(ctx, plane * nz);
  GridDesc g;
  g.n[0] = nx; g.n[1] = ny; g.n[2] = nz;
  for (int d = 0; d < 3; d++) {
    g.lo[d] = lo[d];
    g.h[d] = (hi[d] - lo[d]) / (g.n[d] - 1);
    g.jh[d] = jitter * g.h[d];
  }
  g.seed = seed;
  // hex layers that can touch an owned vertex
  int64_t kz_lo = 0, kz_hi = nz - 1;  // hex layers [kz_lo, kz_hi)
  if (ctx->ve > ctx->vb) {
    kz_lo = std::max<int64_t>(0, ctx->vb / plane - 1);
    kz_hi = std::min<int64_t>(nz - 1, (ctx->ve - 1) / plane + 1);
  } else {
    kz_hi = kz_lo;
  }
  const int64_t hex_per_layer = (int64_t)(nx - 1) * (ny - 1);
  const int64_t nhex = hex_per_layer * (kz_hi - kz_lo);
  DBuf<int32_t> cellsG;
  cellsG.alloc(nhex * 6 * 4);
  if (nhex > 0)
    LAUNCH(ctx, k_tetgrid_cells, nhex * 6, kz_lo * hex_per_layer, nhex, nx, ny, (int4 *)cellsG.p);
  build_from_cells<4, 6>(ctx, cellsG, nhex * 6, nullptr, &g);
}

}  // namespace nosh
