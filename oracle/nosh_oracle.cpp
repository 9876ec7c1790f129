// =============================================================================
// nosh_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain C++ restatement (C ABI, no third-party dependencies) of the
// Newton-Krylov hot path of nschloe/nosh, following the reference source line
// by line and keeping the reference's *data layout*: a real 2N x 2N CSR matrix
// with int column indices (what Tpetra::CrsMatrix<double,int,int> stores),
// interleaved (re,im) vectors, the SpMV followed by a separate diagonal
// epilogue loop.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product path
// (nosh_b200/csrc) never does; it fails loudly when its CUDA library is absent.
//
// PARITY PINNING: the arithmetic restated here is pinned against the
// reference's own known-answer numbers for the `rectanglesmall` and
// `cubesmall` fixtures (test/mesh.cpp:40-49,88-97, test/keo.cpp:118-141,155-167,
// test/compute_f.cpp:74-109, test/jac.cpp:113-148) -- see
// tests/test_oracle_golden.py.  The reference itself cannot be built here
// (Trilinos / MOAB / Eigen / MPI absent).  The Krylov (MINRES/CG) and Newton
// parts live in Trilinos Belos / NOX, which are NOT in the reference tree:
// they are restated from the published algorithms (Paige-Saunders MINRES as
// organised in Belos::MinresIter) and are "parity unpinned" -- no reference
// test pins their iteration counts.
//
// Reference files followed (all under /root/reference/src):
//   mesh.cpp:629-691 (relations), :895-1027 (triangle helpers)
//   mesh_tetra.cpp:41-415, mesh_tri.cpp:44-210
//   vector_field_explicit_values.cpp:13-90, vector_field_constant_curl.cpp:74-227
//   scalar_field_constant.cpp:43-75
//   parameter_matrix_keo.cpp:74-231, parameter_matrix_dkeo_dp.cpp:60-155
//   jacobian_operator.cpp:38-199, model_evaluator_nls.cpp:527-695
//   keo_regularized.cpp:181-264
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

struct V3 {
  double x, y, z;
};
inline V3 operator+(const V3 &a, const V3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3 &a, const V3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, const V3 &a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3 &a, const V3 &b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
inline V3 load3(const double *c, int64_t i) { return {c[3 * i], c[3 * i + 1], c[3 * i + 2]}; }

// Dense solve by LU with full pivoting, the algorithm behind Eigen's
// A.fullPivLu().solve(rhs) used at mesh_tetra.cpp:155 / mesh_tri.cpp:147.
// Pivot = entry of largest magnitude in the trailing block (column-major scan,
// first strictly-greater wins, as Eigen's maxCoeff visitor does).
template <int n>
bool full_piv_lu_solve(double A[n][n], const double *rhs, double *x) {
  int rowp[n], colp[n];
  for (int i = 0; i < n; i++) {
    rowp[i] = i;
    colp[i] = i;
  }
  double b[n];
  for (int i = 0; i < n; i++) b[i] = rhs[i];
  for (int k = 0; k < n; k++) {
    int pr = k, pc = k;
    double big = -1.0;
    for (int j = k; j < n; j++)
      for (int i = k; i < n; i++) {
        double v = std::fabs(A[i][j]);
        if (v > big) {
          big = v;
          pr = i;
          pc = j;
        }
      }
    if (big == 0.0) return false;
    if (pr != k) {
      for (int j = 0; j < n; j++) std::swap(A[k][j], A[pr][j]);
      std::swap(b[k], b[pr]);
      std::swap(rowp[k], rowp[pr]);
    }
    if (pc != k) {
      for (int i = 0; i < n; i++) std::swap(A[i][k], A[i][pc]);
      std::swap(colp[k], colp[pc]);
    }
    for (int i = k + 1; i < n; i++) {
      double l = A[i][k] / A[k][k];
      A[i][k] = l;
      for (int j = k + 1; j < n; j++) A[i][j] -= l * A[k][j];
      b[i] -= l * b[k];
    }
  }
  double yv[n];
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int j = i + 1; j < n; j++) s -= A[i][j] * yv[j];
    yv[i] = s / A[i][i];
  }
  for (int i = 0; i < n; i++) x[colp[i]] = yv[i];
  return true;
}

// mesh.cpp:895-927
V3 triangle_circumcenter(const V3 &n0, const V3 &n1, const V3 &n2, bool *ok) {
  V3 a = n0 - n1, b = n1 - n2, c = n2 - n0;
  V3 ab = cross(a, b);
  const double omega = 2.0 * dot(ab, ab);
  if (std::fabs(omega) < 1.0e-10) *ok = false;
  const double alpha = -dot(b, b) * dot(a, c) / omega;
  const double beta = -dot(c, c) * dot(b, a) / omega;
  const double gamma = -dot(a, a) * dot(c, b) / omega;
  return alpha * n0 + beta * n1 + gamma * n2;
}

// mesh_tetra.cpp:375-415
V3 tetra_circumcenter(const V3 v[4], bool *ok) {
  V3 r0 = v[1] - v[0], r1 = v[2] - v[0], r2 = v[3] - v[0];
  double omega = 2.0 * dot(r0, cross(r1, r2));
  if (std::fabs(omega) < 1.0e-10) *ok = false;
  const double alpha = dot(r0, r0) / omega;
  const double beta = dot(r1, r1) / omega;
  const double gamma = dot(r2, r2) / omega;
  return v[0] + alpha * cross(r1, r2) + beta * cross(r2, r0) + gamma * cross(r0, r1);
}

// mesh_tetra.cpp:273-331
double covolume3d(const V3 &cc, const V3 &x0, const V3 &x1, const V3 &other0, const V3 &other1,
                  bool *ok) {
  double covolume = 0.0;
  V3 mp = 0.5 * (x0 + x1);
  V3 ccf0 = triangle_circumcenter(x0, x1, other0, ok);
  V3 ccf1 = triangle_circumcenter(x0, x1, other1, ok);
  V3 gauge = cross(other0 - mp, other1 - mp);
  double h0 = norm(mp - ccf0);
  double a0 = 0.5 * h0 * norm(ccf0 - cc);
  V3 n0 = cross(ccf0 - mp, cc - mp);
  covolume += std::copysign(a0, dot(n0, gauge));
  double h1 = norm(mp - ccf1);
  double a1 = 0.5 * h1 * norm(ccf1 - cc);
  V3 n1 = cross(cc - mp, ccf1 - mp);
  covolume += std::copysign(a1, dot(n1, gauge));
  return covolume;
}

// mesh.cpp:1003-1027
double covolume2d(const V3 &cc, const V3 &x0, const V3 &x1, const V3 &other0) {
  V3 mp = 0.5 * (x0 + x1);
  double coedge_length = norm(mp - cc);
  V3 cell_normal = cross(other0 - x0, mp - x0);
  V3 cc_normal = cross(cc - x0, mp - x0);
  return std::copysign(coedge_length, dot(cc_normal, cell_normal));
}

// symOrtho of Belos::MinresIter (Givens rotation, Paige-Saunders' SymOrtho).
// [Belos is not in the reference tree -- restated from the published routine.]
void sym_ortho(double a, double b, double *c, double *s, double *r) {
  const double absA = std::fabs(a), absB = std::fabs(b);
  if (absB == 0.0) {
    *s = 0.0;
    *r = absA;
    *c = (absA == 0.0) ? 1.0 : (a >= 0.0 ? 1.0 : -1.0);
  } else if (absA == 0.0) {
    *c = 0.0;
    *s = (b >= 0.0 ? 1.0 : -1.0);
    *r = absB;
  } else if (absB >= absA) {
    double tau = a / b;
    *s = (b >= 0.0 ? 1.0 : -1.0) / std::sqrt(1.0 + tau * tau);
    *c = *s * tau;
    *r = b / *s;
  } else {
    double tau = b / a;
    *c = (a >= 0.0 ? 1.0 : -1.0) / std::sqrt(1.0 + tau * tau);
    *s = *c * tau;
    *r = a / *c;
  }
}

const int TET_PAIR[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
const int TRI_PAIR[3][2] = {{0, 1}, {0, 2}, {1, 2}};

}  // namespace

// -----------------------------------------------------------------------------
// a1: entity relations.  mesh.cpp:629-691 -- unique edges with v0 < v1 (MOAB
// Ranges are handle-sorted), per cell its edges (a Range, hence sorted by edge
// handle => here: ascending edge id, edges numbered in (v0,v1) lexicographic
// order).
// Pass edges == NULL to query the edge count.
// cells: nc x (dim+1) int32 vertex ids (0-based).  Returns #edges or <0.
// -----------------------------------------------------------------------------
ORC_API int64_t orc_build_edges(int dim, int64_t nc, const int32_t *cells, int32_t *edges /*E x 2*/,
                                int32_t *cell_edges /*nc x ne*/) {
  const int nvc = dim + 1;
  const int ne = (dim == 3) ? 6 : 3;
  const int(*pair)[2] = (dim == 3) ? TET_PAIR : TRI_PAIR;
  std::vector<uint64_t> keys((size_t)nc * ne);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nc; k++)
    for (int e = 0; e < ne; e++) {
      uint32_t a = (uint32_t)cells[k * nvc + pair[e][0]];
      uint32_t b = (uint32_t)cells[k * nvc + pair[e][1]];
      if (a > b) std::swap(a, b);
      keys[(size_t)k * ne + e] = ((uint64_t)a << 32) | b;
    }
  std::vector<uint64_t> uniq(keys);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  const int64_t E = (int64_t)uniq.size();
  if (!edges) return E;
  for (int64_t i = 0; i < E; i++) {
    edges[2 * i] = (int32_t)(uniq[i] >> 32);
    edges[2 * i + 1] = (int32_t)(uniq[i] & 0xffffffffu);
  }
  if (cell_edges) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nc; k++) {
      int32_t ids[6];
      for (int e = 0; e < ne; e++) {
        uint64_t key = keys[(size_t)k * ne + e];
        ids[e] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), key) - uniq.begin());
      }
      std::sort(ids, ids + ne);  // Range order = ascending handle
      for (int e = 0; e < ne; e++) cell_edges[k * ne + e] = ids[e];
    }
  }
  return E;
}

// -----------------------------------------------------------------------------
// a2: edge data.  mesh_tetra.cpp:41-160 (3D), mesh_tri.cpp:44-151 (2D).
// length_e = |x_v0 - x_v1|; covolume_e = sum_cells coeff_i * length_e where
// coeff solves  sum_i coeff_i <u,e_i><e_i,v> = |cell| <u,v>  on the cell's edges.
// Returns 0, or -1 for a too-flat tetrahedron (the reference throws).
// -----------------------------------------------------------------------------
ORC_API int orc_edge_data(int dim, int64_t nv, const double *coords, int64_t nc,
                          const int32_t *cell_edges, int64_t E, const int32_t *edges,
                          double *length, double *covolume) {
  (void)nv;
  std::vector<V3> ec(E);
  for (int64_t k = 0; k < E; k++) {
    ec[k] = load3(coords, edges[2 * k]) - load3(coords, edges[2 * k + 1]);  // :65-67
    length[k] = norm(ec[k]);                                                  // :69
    covolume[k] = 0.0;  // value-initialised std::vector<edge_data> (:54)
  }
  int err = 0;
  // The reference loops cells serially (:73); keep the order so every edge
  // receives its cell contributions in ascending cell index.
  for (int64_t k = 0; k < nc; k++) {
    if (dim == 3) {
      V3 e[6];
      for (int i = 0; i < 6; i++) e[i] = ec[cell_edges[6 * k + i]];
      // get_tetrahedron_volume_ :351-373 with the retry of :122-128
      double vol = 0.0;
      {
        double alpha = dot(e[0], cross(e[1], e[2]));
        if (std::fabs(alpha) / norm(e[0]) / norm(e[1]) / norm(e[2]) < 1.0e-5) {
          alpha = dot(e[0], cross(e[1], e[3]));
          if (std::fabs(alpha) / norm(e[0]) / norm(e[1]) / norm(e[3]) < 1.0e-5) err = -1;
        }
        vol = std::fabs(alpha) / 6.0;
      }
      double A[6][6], rhs[6], x[6];
      for (int i = 0; i < 6; i++) {  // :140-149
        double alpha = dot(e[i], e[i]);
        rhs[i] = vol * alpha;
        A[i][i] = alpha * alpha;
        for (int j = i + 1; j < 6; j++) {
          A[i][j] = dot(e[i], e[j]) * dot(e[j], e[i]);
          A[j][i] = A[i][j];
        }
      }
      if (!full_piv_lu_solve<6>(A, rhs, x)) err = -1;
      for (int i = 0; i < 6; i++) {  // :95-99
        const int32_t ei = cell_edges[6 * k + i];
        covolume[ei] += x[i] * length[ei];
      }
    } else {
      V3 e[3];
      for (int i = 0; i < 3; i++) e[i] = ec[cell_edges[3 * k + i]];
      const double vol = 0.5 * norm(cross(e[0], e[1]));  // mesh_tri.cpp:121
      double A[3][3], rhs[3], x[3];
      for (int i = 0; i < 3; i++) {
        double alpha = dot(e[i], e[i]);
        rhs[i] = vol * alpha;
        A[i][i] = alpha * alpha;
        for (int j = i + 1; j < 3; j++) {
          A[i][j] = dot(e[i], e[j]) * dot(e[j], e[i]);
          A[j][i] = A[i][j];
        }
      }
      if (!full_piv_lu_solve<3>(A, rhs, x)) err = -1;
      for (int i = 0; i < 3; i++) {
        const int32_t ei = cell_edges[3 * k + i];
        covolume[ei] += x[i] * length[ei];
      }
    }
  }
  return err;
}

// -----------------------------------------------------------------------------
// a3: control volumes.  mesh_tetra.cpp:162-271 (3D), mesh_tri.cpp:185-210 with
// mesh.cpp:929-1027 (2D).  Serial (one "rank"): the Export/ADD of :185-189 is
// the identity.
// -----------------------------------------------------------------------------
ORC_API int orc_control_volumes(int dim, int64_t nv, const double *coords, int64_t nc,
                                const int32_t *cells, double *cv) {
  for (int64_t i = 0; i < nv; i++) cv[i] = 0.0;
  bool ok = true;
  for (int64_t k = 0; k < nc; k++) {
    if (dim == 3) {
      V3 x[4];
      for (int i = 0; i < 4; i++) x[i] = load3(coords, cells[4 * k + i]);
      const V3 cc = tetra_circumcenter(x, &ok);
      for (int e0 = 0; e0 < 4; e0++)
        for (int e1 = e0 + 1; e1 < 4; e1++) {
          int other[2], n = 0;
          for (int i = 0; i < 4; i++)
            if (i != e0 && i != e1) other[n++] = i;  // std::set order (:229-231)
          const double edge_length = norm(x[e1] - x[e0]);
          const double covol = covolume3d(cc, x[e0], x[e1], x[other[0]], x[other[1]], &ok);
          const double pyramid = 0.5 * edge_length * covol / 3;  // :263
          cv[cells[4 * k + e0]] += pyramid;
          cv[cells[4 * k + e1]] += pyramid;
        }
    } else {
      V3 x[3];
      for (int i = 0; i < 3; i++) x[i] = load3(coords, cells[3 * k + i]);
      const V3 cc = triangle_circumcenter(x[0], x[1], x[2], &ok);
      double split[3] = {0.0, 0.0, 0.0};
      for (int e0 = 0; e0 < 3; e0++)
        for (int e1 = e0 + 1; e1 < 3; e1++) {
          const int other = 3 - e0 - e1;
          const double edge_length = norm(x[e1] - x[e0]);
          const double covol = covolume2d(cc, x[e0], x[e1], x[other]);
          const double pyramid = 0.5 * edge_length * covol / 2;  // mesh.cpp:971
          split[e0] += pyramid;
          split[e1] += pyramid;
        }
      for (int i = 0; i < 3; i++) cv[cells[3 * k + i]] += split[i];
    }
  }
  return ok ? 0 : -1;
}

// -----------------------------------------------------------------------------
// a5: vector_field::explicit_values edge-projection cache
// (vector_field_explicit_values.cpp:31-48): cache_e = 0.5 (A_v0 + A_v1).(x_v0 - x_v1)
// a_e = mu * cache_e (:73-76); d a_e / d mu = cache_e, other names 0 (:79-90).
// -----------------------------------------------------------------------------
ORC_API void orc_edge_cache_explicit(const double *coords, const double *A, int64_t E,
                                     const int32_t *edges, double *cache) {
  for (int64_t k = 0; k < E; k++) {
    const int32_t i0 = edges[2 * k], i1 = edges[2 * k + 1];
    V3 av = 0.5 * (load3(A, i0) + load3(A, i1));
    V3 ecoord = load3(coords, i0) - load3(coords, i1);
    cache[k] = dot(av, ecoord);
  }
}

// -----------------------------------------------------------------------------
// a6: vector_field::constantCurl (vector_field_constant_curl.cpp:74-227).
// The reference's initializeEdgeCache_ is a stub that throws (:208); its dead
// formula has the opposite sign of its own comment.  Restated so that
// constantCurl(B) == explicit_values(A = 0.5 B x X):  edgeCache_e = 0.5 x_v1 x x_v0.
// cache3: E x 3.
// -----------------------------------------------------------------------------
ORC_API void orc_edge_cache_constcurl(const double *coords, int64_t E, const int32_t *edges,
                                      double *cache3) {
  for (int64_t k = 0; k < E; k++) {
    V3 c = 0.5 * cross(load3(coords, edges[2 * k + 1]), load3(coords, edges[2 * k]));
    cache3[3 * k] = c.x;
    cache3[3 * k + 1] = c.y;
    cache3[3 * k + 2] = c.z;
  }
}

// rotate_ (:144-170) and dRotateDTheta_ (:174-197).  u may be NULL (then theta
// must be 0 and no rotation is applied).  Outputs rb = R_theta(b), drb = dR/dtheta b.
ORC_API void orc_constcurl_rotate(const double *b, const double *u, double theta, double *rb,
                                  double *drb) {
  V3 B = {b[0], b[1], b[2]};
  V3 v = B, dv = B;
  if (u) {
    V3 U = {u[0], u[1], u[2]};
    double s, c;
    sincos(theta, &s, &c);
    if (s != 0.0) v = c * B + s * cross(U, B) + ((1.0 - c) * dot(U, B)) * U;
    dv = (-s) * B + c * cross(U, B) + ((1.0 + s) * dot(U, B)) * U;
  }
  rb[0] = v.x;
  rb[1] = v.y;
  rb[2] = v.z;
  drb[0] = dv.x;
  drb[1] = dv.y;
  drb[2] = dv.z;
}

// a_e = mu * rb . cache3_e (:103); d/dmu = rb . cache3_e (:131);
// d/dtheta = mu * drb . cache3_e (:133).  Any output may be NULL.
ORC_API void orc_constcurl_projection(const double *rb, const double *drb, double mu, int64_t E,
                                      const double *cache3, double *a, double *da_dmu,
                                      double *da_dtheta) {
  for (int64_t k = 0; k < E; k++) {
    const double p = rb[0] * cache3[3 * k] + rb[1] * cache3[3 * k + 1] + rb[2] * cache3[3 * k + 2];
    if (a) a[k] = mu * p;
    if (da_dmu) da_dmu[k] = p;
    if (da_dtheta)
      da_dtheta[k] =
          mu * (drb[0] * cache3[3 * k] + drb[1] * cache3[3 * k + 1] + drb[2] * cache3[3 * k + 2]);
  }
}

// -----------------------------------------------------------------------------
// a8: keo::build_alpha_cache_ (parameter_matrix_keo.cpp:186-231)
// alpha_e = (covolume_e / length_e) * 0.5 (t_v0 + t_v1)
// -----------------------------------------------------------------------------
ORC_API void orc_alpha_cache(int64_t E, const int32_t *edges, const double *length,
                             const double *covolume, const double *thickness, double *alpha) {
  for (int64_t k = 0; k < E; k++) {
    const double a = covolume[k] / length[k];
    alpha[k] = a * 0.5 * (thickness[edges[2 * k]] + thickness[edges[2 * k + 1]]);
  }
}

// -----------------------------------------------------------------------------
// a4: mesh::build_complex_graph (mesh.cpp:787-881).  Real 2N x 2N CRS pattern:
// for every edge rows idx[0..3] x cols idx[0..3], idx = {2v0,2v0+1,2v1,2v1+1};
// fillComplete sorts/merges the rows.  rowptr: 2N+1 int64, cols: int32.
// Pass cols == NULL to get the nnz count only.
// -----------------------------------------------------------------------------
ORC_API int64_t orc_build_complex_graph(int64_t nv, int64_t E, const int32_t *edges,
                                        int64_t *rowptr, int32_t *cols) {
  std::vector<int32_t> deg(nv, 0);
  for (int64_t k = 0; k < E; k++) {
    deg[edges[2 * k]]++;
    deg[edges[2 * k + 1]]++;
  }
  // a vertex without edges has an empty row in the reference; keep that.
  std::vector<int64_t> vptr(nv + 1, 0);
  for (int64_t i = 0; i < nv; i++) vptr[i + 1] = vptr[i] + (deg[i] ? deg[i] + 1 : 0);
  const int64_t nnz = 4 * vptr[nv];
  if (!cols) return nnz;
  std::vector<int32_t> nb(vptr[nv]);
  std::vector<int64_t> fill(vptr.begin(), vptr.end() - 1);
  for (int64_t i = 0; i < nv; i++)
    if (deg[i]) nb[fill[i]++] = (int32_t)i;
  for (int64_t k = 0; k < E; k++) {
    const int32_t a = edges[2 * k], b = edges[2 * k + 1];
    nb[fill[a]++] = b;
    nb[fill[b]++] = a;
  }
  rowptr[0] = 0;
  for (int64_t i = 0; i < nv; i++) {
    std::sort(nb.begin() + vptr[i], nb.begin() + vptr[i + 1]);
    const int64_t w = 2 * (vptr[i + 1] - vptr[i]);
    rowptr[2 * i + 1] = rowptr[2 * i] + w;
    rowptr[2 * i + 2] = rowptr[2 * i + 1] + w;
    for (int r = 0; r < 2; r++) {
      int32_t *c = cols + rowptr[2 * i + r];
      for (int64_t j = vptr[i]; j < vptr[i + 1]; j++) {
        *c++ = 2 * nb[j];
        *c++ = 2 * nb[j] + 1;
      }
    }
  }
  return nnz;
}

namespace {
// Tpetra sumIntoGlobalValues: locate each column in the (sorted) row, add.
inline int sum_into(const int64_t *rowptr, const int32_t *cols, double *vals, int32_t row,
                    const int32_t idx[4], const double v[4], bool atomic) {
  const int32_t *b = cols + rowptr[row], *e = cols + rowptr[row + 1];
  int n = 0;
  for (int j = 0; j < 4; j++) {
    const int32_t *p = std::lower_bound(b, e, idx[j]);
    if (p != e && *p == idx[j]) {
      double *dst = vals + (p - cols);
      if (atomic) {
#pragma omp atomic
        *dst += v[j];
      } else {
        *dst += v[j];
      }
      n++;
    }
  }
  return n;
}
}  // namespace

// -----------------------------------------------------------------------------
// a9: keo::refill_ (parameter_matrix_keo.cpp:74-184).  a[k] = get_edge_projection(k).
// mode 0 = KEO; mode 1 = DkeoDP::refill_ (parameter_matrix_dkeo_dp.cpp:60-155,
// a10) with da[k] = get_d_edge_projection_dp(k, param_name).
// nthreads > 1 parallelises the edge loop with atomic adds (baseline timing).
// -----------------------------------------------------------------------------
ORC_API int orc_keo_fill(int mode, int64_t nv, int64_t E, const int32_t *edges,
                         const int64_t *rowptr, const int32_t *cols, const double *alpha,
                         const double *a, const double *da, double *vals, int nthreads) {
  const int64_t nnz = rowptr[2 * nv];
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t i = 0; i < nnz; i++) vals[i] = 0.0;  // setAllToScalar(0.0) (:90)
  int bad = 0;
  const bool atomic = nthreads > 1;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : bad)
  for (int64_t k = 0; k < E; k++) {
    double s, c, v[3];
    sincos(a[k], &s, &c);  // :143
    if (mode == 0) {
      v[0] = -c * alpha[k];  // :144-146
      v[1] = -s * alpha[k];
      v[2] = alpha[k];
    } else {
      v[0] = da[k] * s;  // dkeo :110-120
      v[1] = -da[k] * c;
      v[2] = 0.0;
      v[0] *= alpha[k];
      v[1] *= alpha[k];
      v[2] *= alpha[k];
    }
    const double vals4[4][4] = {{v[2], 0.0, v[0], v[1]},  // :148-153
                                {0.0, v[2], -v[1], v[0]},
                                {v[0], -v[1], v[2], 0.0},
                                {v[1], v[0], 0.0, v[2]}};
    const int32_t idx[4] = {2 * edges[2 * k], 2 * edges[2 * k] + 1, 2 * edges[2 * k + 1],
                            2 * edges[2 * k + 1] + 1};
    for (int i = 0; i < 4; i++)
      if (sum_into(rowptr, cols, vals, idx[i], idx, vals4[i], atomic) != 4) bad++;
  }
  return bad ? -1 : 0;
}

// -----------------------------------------------------------------------------
// a16: keo_regularized::rebuild (keo_regularized.cpp:181-264): after the KEO
// fill, add [[al+ga, be],[be, al-ga]] to every diagonal 2x2 block, only if g>0.
// -----------------------------------------------------------------------------
ORC_API int orc_keoreg_add_diag(int64_t nv, const int64_t *rowptr, const int32_t *cols,
                                double *vals, double g, const double *cvol, const double *thick,
                                const double *x) {
  if (!(g > 0.0)) return 0;  // :200
  int bad = 0;
  for (int64_t k = 0; k < nv; k++) {
    const double al =
        g * cvol[k] * thick[k] * 2.0 * (x[2 * k] * x[2 * k] + x[2 * k + 1] * x[2 * k + 1]);
    const double be = g * cvol[k] * thick[k] * (2.0 * x[2 * k] * x[2 * k + 1]);
    const double ga = g * cvol[k] * thick[k] * (x[2 * k] * x[2 * k] - x[2 * k + 1] * x[2 * k + 1]);
    for (int r = 0; r < 2; r++) {
      const int32_t row = (int32_t)(2 * k + r);
      const int32_t *b = cols + rowptr[row], *e = cols + rowptr[row + 1];
      const double v[2] = {r == 0 ? al + ga : be, r == 0 ? be : al - ga};
      for (int j = 0; j < 2; j++) {
        const int32_t col = (int32_t)(2 * k + j);
        const int32_t *p = std::lower_bound(b, e, col);
        if (p != e && *p == col)
          vals[p - cols] += v[j];
        else
          bad++;
      }
    }
  }
  return bad ? -1 : 0;
}

// -----------------------------------------------------------------------------
// a19: Tpetra::CrsMatrix::apply, local part: y = A x on a real CSR matrix.
// Rows are statically partitioned over `nthreads` (one MPI rank per core with
// Tpetra's serial node).
// -----------------------------------------------------------------------------
// g_row_reverse: sum every row right to left instead of left to right -- a rounding-level perturbation of the
// operator apply (what a different column order in Tpetra's local graph, or the GPU's fused-multiply-add complex
// arithmetic, amounts to), used to measure how sensitive the Krylov iteration counts are to it
// (tests/golden/make_counts_golden.py).
static int g_row_reverse = 0;
ORC_API void orc_set_row_reverse(int on) { g_row_reverse = on != 0; }
ORC_API void orc_csr_apply(int64_t nrows, const int64_t *rowptr, const int32_t *cols,
                           const double *vals, const double *x, double *y, int nthreads) {
  const int rev = g_row_reverse;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t i = 0; i < nrows; i++) {
    double s = 0.0;
    if (rev)
      for (int64_t j = rowptr[i + 1] - 1; j >= rowptr[i]; j--) s += vals[j] * x[cols[j]];
    else
      for (int64_t j = rowptr[i]; j < rowptr[i + 1]; j++) s += vals[j] * x[cols[j]];
    y[i] = s;
  }
}

// -----------------------------------------------------------------------------
// a12: jacobian_operator::rebuild_diags_ (jacobian_operator.cpp:143-199)
// -----------------------------------------------------------------------------
ORC_API void orc_jac_diags(int64_t nv, double g, const double *c, const double *t, const double *s,
                           const double *x, double *d0, double *d1b, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t k = 0; k < nv; k++) {
    const double alpha =
        c[k] * t[k] * (s[k] + g * 2.0 * (x[2 * k] * x[2 * k] + x[2 * k + 1] * x[2 * k + 1]));
    const double realX2 = g * c[k] * t[k] * (x[2 * k] * x[2 * k] - x[2 * k + 1] * x[2 * k + 1]);
    d0[2 * k] = alpha + realX2;
    d0[2 * k + 1] = alpha - realX2;
    d1b[k] = g * c[k] * t[k] * (2.0 * x[2 * k] * x[2 * k + 1]);
  }
}

// -----------------------------------------------------------------------------
// a11: jacobian_operator::apply (jacobian_operator.cpp:38-108): Y = K X, then
// the per-vertex 2x2 diagonal epilogue (:95-100), column by column.
// X, Y column-major with leading dimensions ldx, ldy.
// -----------------------------------------------------------------------------
ORC_API void orc_jac_apply(int64_t nv, const int64_t *rowptr, const int32_t *cols,
                           const double *vals, const double *d0, const double *d1b, int nvec,
                           const double *X, int64_t ldx, double *Y, int64_t ldy, int nthreads) {
  for (int v = 0; v < nvec; v++) {
    const double *x = X + (size_t)v * ldx;
    double *y = Y + (size_t)v * ldy;
    orc_csr_apply(2 * nv, rowptr, cols, vals, x, y, nthreads);  // :65
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t k = 0; k < nv; k++) {
      y[2 * k] += d0[2 * k] * x[2 * k] + d1b[k] * x[2 * k + 1];
      y[2 * k + 1] += d1b[k] * x[2 * k] + d0[2 * k + 1] * x[2 * k + 1];
    }
  }
}

// -----------------------------------------------------------------------------
// a13: nls::compute_f_ (model_evaluator_nls.cpp:527-628); `vals` = KEO values.
// -----------------------------------------------------------------------------
ORC_API void orc_compute_f(int64_t nv, const int64_t *rowptr, const int32_t *cols,
                           const double *vals, double g, const double *c, const double *t,
                           const double *s, const double *x, double *f, int nthreads) {
  orc_csr_apply(2 * nv, rowptr, cols, vals, x, f, nthreads);  // :537
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t k = 0; k < nv; k++) {
    const double alpha =
        c[k] * t[k] * (s[k] + g * (x[2 * k] * x[2 * k] + x[2 * k + 1] * x[2 * k + 1]));  // :618
    f[2 * k] += alpha * x[2 * k];
    f[2 * k + 1] += alpha * x[2 * k + 1];
  }
}

// -----------------------------------------------------------------------------
// a14: nls::computeDFDP_ (model_evaluator_nls.cpp:630-695); `dvals` = dKEO/dp
// values.  is_g != 0: the "g" branch (:665-674); else dvdp = dV/dp vector (:676-691).
// -----------------------------------------------------------------------------
ORC_API void orc_compute_dfdp(int64_t nv, const int64_t *rowptr, const int32_t *cols,
                              const double *dvals, int is_g, const double *c, const double *t,
                              const double *dvdp, const double *x, double *f) {
  orc_csr_apply(2 * nv, rowptr, cols, dvals, x, f, 1);  // :641
  for (int64_t k = 0; k < nv; k++) {
    double alpha;
    if (is_g)
      alpha = c[k] * t[k] * (x[2 * k] * x[2 * k] + x[2 * k + 1] * x[2 * k + 1]);
    else
      alpha = c[k] * t[k] * dvdp[k];
    f[2 * k] += alpha * x[2 * k];
    f[2 * k + 1] += alpha * x[2 * k + 1];
  }
}

// -----------------------------------------------------------------------------
// BLAS-1 as Tpetra::MultiVector::{dot,norm2,update,scale}: serial left-to-right
// sums per thread chunk, chunks combined in thread order (what an MPI
// all-reduce over ranks does).
// -----------------------------------------------------------------------------
namespace {
// g_dot_parts > 0: the sum is split into that many contiguous parts ("ranks") whatever the number of
// threads executing it, so iteration counts can be generated for the reference's rank counts (1, 2, 7:
// test/CMakeLists.txt:18-23) independently of the machine the oracle runs on.  0: parts = threads.
int g_dot_parts = 0;
double pdot(int64_t n, const double *a, const double *b, int nt) {
  const int parts = g_dot_parts > 0 ? g_dot_parts : (nt > 1 ? nt : 1);
  if (parts <= 1) {
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
  }
  std::vector<double> part(parts, 0.0);
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
  for (int t = 0; t < parts; t++) {
    int64_t lo = n * t / parts, hi = n * (t + 1) / parts;
    double s = 0.0;
    for (int64_t i = lo; i < hi; i++) s += a[i] * b[i];
    part[t] = s;
  }
  double s = 0.0;
  for (int t = 0; t < parts; t++) s += part[t];
  return s;
}

struct JacOp {
  int64_t nv;
  const int64_t *rowptr;
  const int32_t *cols;
  const double *vals;
  const double *d0, *d1b;  // NULL => plain KEO apply
  int nt;
  void apply(const double *x, double *y) const {
    if (d0)
      orc_jac_apply(nv, rowptr, cols, vals, d0, d1b, 1, x, 2 * nv, y, 2 * nv, nt);
    else
      orc_csr_apply(2 * nv, rowptr, cols, vals, x, y, nt);
  }
};

// Preconditioned MINRES organised as Belos::MinresIter::iterate()
// [Belos not in the reference tree: restated from Paige & Saunders 1975 /
// Choi's SymOrtho as used by Belos; unpinned].  No preconditioner (the live
// reference path sets "Preconditioner Type" = "None",
// model_evaluator_nls.cpp:291).  x0 = 0.  Convergence: implicit residual
// phibar / ||r0|| <= tol, checked before each iteration; at most maxit its.
int minres(const JacOp &A, const double *b, double *x, double tol, int maxit, double *relres,
           double *hist /* maxit+1 or NULL */) {
  const int64_t n = 2 * A.nv;
  const int nt = A.nt;
  std::vector<double> Y(b, b + n), V(n), R1(b, b + n), R2(b, b + n), W(n, 0.0), W1(n, 0.0),
      W2(n, 0.0);
  for (int64_t i = 0; i < n; i++) x[i] = 0.0;
  double beta1 = pdot(n, R1.data(), Y.data(), nt);
  if (hist) hist[0] = 1.0;
  if (beta1 <= 0.0) {
    *relres = 0.0;
    return 0;
  }
  beta1 = std::sqrt(beta1);
  double oldBeta = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, oldeps, phibar = beta1, cs = -1.0,
         sn = 0.0, alpha, delta, gbar, gamma, phi;
  int iter = 0;
  while (iter < maxit && !(phibar / beta1 <= tol)) {
    iter++;
    const double ib = 1.0 / beta;
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
    for (int64_t i = 0; i < n; i++) V[i] = Y[i] * ib;
    A.apply(V.data(), Y.data());
    if (iter > 1) {
      const double f = beta / oldBeta;
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
      for (int64_t i = 0; i < n; i++) Y[i] -= f * R1[i];
    }
    alpha = pdot(n, V.data(), Y.data(), nt);
    {
      const double f = alpha / beta;
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
      for (int64_t i = 0; i < n; i++) {
        Y[i] -= f * R2[i];
        R1[i] = R2[i];
        R2[i] = Y[i];
      }
    }
    oldBeta = beta;
    beta = pdot(n, R2.data(), Y.data(), nt);
    if (beta < 0.0) break;
    beta = std::sqrt(beta);
    oldeps = epsln;
    delta = cs * dbar + sn * alpha;
    gbar = sn * dbar - cs * alpha;
    epsln = sn * beta;
    dbar = -cs * beta;
    sym_ortho(gbar, beta, &cs, &sn, &gamma);
    phi = cs * phibar;
    phibar = sn * phibar;
    if (gamma == 0.0) break;
    const double ig = 1.0 / gamma;
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
    for (int64_t i = 0; i < n; i++) {
      const double w1 = W2[i], w2 = W[i];
      W1[i] = w1;
      W2[i] = w2;
      const double w = (V[i] - oldeps * w1 - delta * w2) * ig;
      W[i] = w;
      x[i] += phi * w;
    }
    if (hist) hist[iter] = phibar / beta1;
  }
  *relres = phibar / beta1;
  return iter;
}

// Unpreconditioned CG organised as Belos::PseudoBlockCGIter (one RHS), the
// live default of model_evaluator_nls.cpp:282.  x0 = 0; ||r||/||r0|| <= tol.
int cg(const JacOp &A, const double *b, double *x, double tol, int maxit, double *relres) {
  const int64_t n = 2 * A.nv;
  const int nt = A.nt;
  std::vector<double> R(b, b + n), P(b, b + n), AP(n);
  for (int64_t i = 0; i < n; i++) x[i] = 0.0;
  double rho = pdot(n, R.data(), R.data(), nt);
  const double r0 = std::sqrt(rho);
  if (r0 == 0.0) {
    *relres = 0.0;
    return 0;
  }
  int iter = 0;
  while (iter < maxit && !(std::sqrt(rho) / r0 <= tol)) {
    iter++;
    A.apply(P.data(), AP.data());
    const double pAp = pdot(n, P.data(), AP.data(), nt);
    const double al = rho / pAp;
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
    for (int64_t i = 0; i < n; i++) {
      x[i] += al * P[i];
      R[i] -= al * AP[i];
    }
    const double rho_new = pdot(n, R.data(), R.data(), nt);
    const double be = rho_new / rho;
    rho = rho_new;
#pragma omp parallel for schedule(static) num_threads(nt > 0 ? nt : 1)
    for (int64_t i = 0; i < n; i++) P[i] = R[i] + be * P[i];
  }
  *relres = std::sqrt(rho) / r0;
  return iter;
}
}  // namespace

// a18: Krylov solve of  A x = b,  A = K + diag terms (d0 != NULL) or K.
// solver 0 = MINRES, 1 = CG.  Returns the iteration count.
ORC_API int orc_krylov(int solver, int64_t nv, const int64_t *rowptr, const int32_t *cols,
                       const double *vals, const double *d0, const double *d1b, const double *b,
                       double *x, double tol, int maxit, double *relres, double *hist,
                       int nthreads) {
  JacOp A{nv, rowptr, cols, vals, d0, d1b, nthreads > 0 ? nthreads : 1};
  if (solver == 0) return minres(A, b, x, tol, maxit, relres, hist);
  return cg(A, b, x, tol, maxit, relres);
}

// -----------------------------------------------------------------------------
// Newton as wired by Piro/NOX for nosh-cont (examples/conf.xml:76-191):
// "Line Search Based" / "Full Step" (step 1), status test ||F||_2 < nl_tol
// (NormF, unscaled) OR nl_maxit iterations; each step: F, rebuild J (KEO refill
// + diags -- with unchanged mu the refill reproduces the same values, so `vals`
// is reused), solve J d = -F by MINRES(x0 = 0, lin_tol, lin_maxit), x += d.
// [NOX is not in the reference tree -- unpinned.]
// lin_iters: nl_maxit ints (MINRES iterations per Newton step);
// fnorms: nl_maxit+1 doubles.  Returns the number of Newton steps taken.
// -----------------------------------------------------------------------------
ORC_API int orc_newton(int64_t nv, const int64_t *rowptr, const int32_t *cols, const double *vals,
                       double g, const double *c, const double *t, const double *s, double *x,
                       double nl_tol, int nl_maxit, double lin_tol, int lin_maxit, int *lin_iters,
                       double *fnorms, int nthreads) {
  const int64_t n = 2 * nv;
  const int nt = nthreads > 0 ? nthreads : 1;
  std::vector<double> F(n), d0(n), d1b(nv), rhs(n), d(n);
  int k = 0;
  orc_compute_f(nv, rowptr, cols, vals, g, c, t, s, x, F.data(), nt);
  double fn = std::sqrt(pdot(n, F.data(), F.data(), nt));
  fnorms[0] = fn;
  while (k < nl_maxit && !(fn < nl_tol)) {
    orc_jac_diags(nv, g, c, t, s, x, d0.data(), d1b.data(), nt);
    for (int64_t i = 0; i < n; i++) rhs[i] = -F[i];
    double relres;
    JacOp A{nv, rowptr, cols, vals, d0.data(), d1b.data(), nt};
    lin_iters[k] = minres(A, rhs.data(), d.data(), lin_tol, lin_maxit, &relres, nullptr);
    for (int64_t i = 0; i < n; i++) x[i] += d[i];
    orc_compute_f(nv, rowptr, cols, vals, g, c, t, s, x, F.data(), nt);
    fn = std::sqrt(pdot(n, F.data(), F.data(), nt));
    k++;
    fnorms[k] = fn;
  }
  return k;
}

ORC_API void orc_set_dot_parts(int parts) { g_dot_parts = parts > 0 ? parts : 0; }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
