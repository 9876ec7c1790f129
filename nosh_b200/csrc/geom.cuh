// geom.cuh -- per-cell geometry on the device: FVM edge coefficients and circumcentric
// control-volume contributions.  What it computes is defined by the reference
// (src/mesh_tetra.cpp:105-160,206-415, src/mesh_tri.cpp:105-151, src/mesh.cpp:895-1027);
// how it is computed is a register-resident per-thread formulation.
#pragma once
#include "common.cuh"

namespace nosh {

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double norm3(V3 a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ V3 ldv3(const double *c, int i) {
  return {c[3 * i], c[3 * i + 1], c[3 * i + 2]};
}

// Solve the n x n system A x = rhs by LU with full pivoting (largest |a_ij| of the
// trailing block, column-major scan, first maximum wins) -- the pivoting rule of the
// solver the reference calls at src/mesh_tetra.cpp:155.  Returns false if singular.
template <int n>
__device__ bool lu_full_pivot_solve(double (&A)[n][n], double (&b)[n], double (&x)[n]) {
  int colp[n];
#pragma unroll
  for (int i = 0; i < n; i++) colp[i] = i;
  for (int k = 0; k < n; k++) {
    int pr = k, pc = k;
    double big = -1.0;
    for (int j = k; j < n; j++)
      for (int i = k; i < n; i++) {
        const double v = fabs(A[i][j]);
        if (v > big) {
          big = v;
          pr = i;
          pc = j;
        }
      }
    if (big == 0.0) return false;
    if (pr != k) {
      for (int j = 0; j < n; j++) {
        const double t = A[k][j];
        A[k][j] = A[pr][j];
        A[pr][j] = t;
      }
      const double t = b[k];
      b[k] = b[pr];
      b[pr] = t;
    }
    if (pc != k) {
      for (int i = 0; i < n; i++) {
        const double t = A[i][k];
        A[i][k] = A[i][pc];
        A[i][pc] = t;
      }
      const int t = colp[k];
      colp[k] = colp[pc];
      colp[pc] = t;
    }
    for (int i = k + 1; i < n; i++) {
      const double l = A[i][k] / A[k][k];
      for (int j = k + 1; j < n; j++) A[i][j] -= l * A[k][j];
      b[i] -= l * b[k];
    }
  }
  double y[n];
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int j = i + 1; j < n; j++) s -= A[i][j] * y[j];
    y[i] = s / A[i][i];
  }
  for (int i = 0; i < n; i++) x[colp[i]] = y[i];
  return true;
}

// coefficients c_i with  sum_i c_i <u,e_i><e_i,v> = vol <u,v>   (src/mesh_tetra.cpp:113-159)
template <int n>
__device__ bool edge_coefficients(const V3 (&e)[n], double vol, double (&coef)[n]) {
  double A[n][n], rhs[n];
#pragma unroll
  for (int i = 0; i < n; i++) {
    const double a = dot3(e[i], e[i]);
    rhs[i] = vol * a;
    A[i][i] = a * a;
#pragma unroll
    for (int j = i + 1; j < n; j++) {
      const double d = dot3(e[i], e[j]);
      A[i][j] = d * d;
      A[j][i] = A[i][j];
    }
  }
  return lu_full_pivot_solve<n>(A, rhs, coef);
}

// src/mesh.cpp:895-927
__device__ __forceinline__ V3 tri_circumcenter(V3 n0, V3 n1, V3 n2, bool &ok) {
  const V3 a = n0 - n1, b = n1 - n2, c = n2 - n0;
  const V3 ab = cross3(a, b);
  const double omega = 2.0 * dot3(ab, ab);
  if (fabs(omega) < 1.0e-10) ok = false;
  const double al = -dot3(b, b) * dot3(a, c) / omega;
  const double be = -dot3(c, c) * dot3(b, a) / omega;
  const double ga = -dot3(a, a) * dot3(c, b) / omega;
  return al * n0 + be * n1 + ga * n2;
}

// src/mesh_tetra.cpp:375-415
__device__ __forceinline__ V3 tet_circumcenter(const V3 (&v)[4], bool &ok) {
  const V3 r0 = v[1] - v[0], r1 = v[2] - v[0], r2 = v[3] - v[0];
  const double omega = 2.0 * dot3(r0, cross3(r1, r2));
  if (fabs(omega) < 1.0e-10) ok = false;
  const double al = dot3(r0, r0) / omega, be = dot3(r1, r1) / omega, ga = dot3(r2, r2) / omega;
  return v[0] + al * cross3(r1, r2) + be * cross3(r2, r0) + ga * cross3(r0, r1);
}

// signed area of the dual facet of edge (x0,x1) inside one tetrahedron
// (src/mesh_tetra.cpp:273-331)
__device__ __forceinline__ double covolume3d(V3 cc, V3 x0, V3 x1, V3 o0, V3 o1, bool &ok) {
  const V3 mp = 0.5 * (x0 + x1);
  const V3 f0 = tri_circumcenter(x0, x1, o0, ok);
  const V3 f1 = tri_circumcenter(x0, x1, o1, ok);
  const V3 gauge = cross3(o0 - mp, o1 - mp);
  const double a0 = 0.5 * norm3(mp - f0) * norm3(f0 - cc);
  const double a1 = 0.5 * norm3(mp - f1) * norm3(f1 - cc);
  double cov = 0.0;
  cov += copysign(a0, dot3(cross3(f0 - mp, cc - mp), gauge));
  cov += copysign(a1, dot3(cross3(cc - mp, f1 - mp), gauge));
  return cov;
}

// src/mesh.cpp:1003-1027
__device__ __forceinline__ double covolume2d(V3 cc, V3 x0, V3 x1, V3 o0) {
  const V3 mp = 0.5 * (x0 + x1);
  const double len = norm3(mp - cc);
  const V3 cell_n = cross3(o0 - x0, mp - x0);
  const V3 cc_n = cross3(cc - x0, mp - x0);
  return copysign(len, dot3(cc_n, cell_n));
}

}  // namespace nosh
