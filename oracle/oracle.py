"""ctypes front-end of the CPU oracle (oracle/nosh_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

``OracleProblem`` strings the restated reference functions together in the order
the reference calls them: mesh relations -> edge data ->
control volumes -> edge-projection cache -> alpha cache -> complex graph ->
KEO fill -> F / J / dF/dp / Krylov / Newton.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnosh_oracle.so")
_lib = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build_library(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "nosh_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnosh_oracle.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build_library()
        L = C.CDLL(_SO)
        vp = C.c_void_p
        L.orc_build_edges.restype = C.c_int64
        L.orc_build_edges.argtypes = [C.c_int, C.c_int64, _i32p, vp, vp]
        L.orc_edge_data.restype = C.c_int
        L.orc_edge_data.argtypes = [C.c_int, C.c_int64, _f64p, C.c_int64, _i32p, C.c_int64, _i32p,
                                    _f64p, _f64p]
        L.orc_control_volumes.restype = C.c_int
        L.orc_control_volumes.argtypes = [C.c_int, C.c_int64, _f64p, C.c_int64, _i32p, _f64p]
        L.orc_edge_cache_explicit.restype = None
        L.orc_edge_cache_explicit.argtypes = [_f64p, _f64p, C.c_int64, _i32p, _f64p]
        L.orc_edge_cache_constcurl.restype = None
        L.orc_edge_cache_constcurl.argtypes = [_f64p, C.c_int64, _i32p, _f64p]
        L.orc_constcurl_rotate.restype = None
        L.orc_constcurl_rotate.argtypes = [_f64p, vp, C.c_double, _f64p, _f64p]
        L.orc_constcurl_projection.restype = None
        L.orc_constcurl_projection.argtypes = [_f64p, _f64p, C.c_double, C.c_int64, _f64p, vp, vp, vp]
        L.orc_alpha_cache.restype = None
        L.orc_alpha_cache.argtypes = [C.c_int64, _i32p, _f64p, _f64p, _f64p, _f64p]
        L.orc_build_complex_graph.restype = C.c_int64
        L.orc_build_complex_graph.argtypes = [C.c_int64, C.c_int64, _i32p, vp, vp]
        L.orc_keo_fill.restype = C.c_int
        L.orc_keo_fill.argtypes = [C.c_int, C.c_int64, C.c_int64, _i32p, _i64p, _i32p, _f64p, _f64p,
                                   vp, _f64p, C.c_int]
        L.orc_keoreg_add_diag.restype = C.c_int
        L.orc_keoreg_add_diag.argtypes = [C.c_int64, _i64p, _i32p, _f64p, C.c_double, _f64p, _f64p,
                                          _f64p]
        L.orc_csr_apply.restype = None
        L.orc_csr_apply.argtypes = [C.c_int64, _i64p, _i32p, _f64p, _f64p, _f64p, C.c_int]
        L.orc_jac_diags.restype = None
        L.orc_jac_diags.argtypes = [C.c_int64, C.c_double, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p,
                                    C.c_int]
        L.orc_jac_apply.restype = None
        L.orc_jac_apply.argtypes = [C.c_int64, _i64p, _i32p, _f64p, _f64p, _f64p, C.c_int, _f64p,
                                    C.c_int64, _f64p, C.c_int64, C.c_int]
        L.orc_compute_f.restype = None
        L.orc_compute_f.argtypes = [C.c_int64, _i64p, _i32p, _f64p, C.c_double, _f64p, _f64p, _f64p,
                                    _f64p, _f64p, C.c_int]
        L.orc_compute_dfdp.restype = None
        L.orc_compute_dfdp.argtypes = [C.c_int64, _i64p, _i32p, _f64p, C.c_int, _f64p, _f64p, vp,
                                       _f64p, _f64p]
        L.orc_krylov.restype = C.c_int
        L.orc_krylov.argtypes = [C.c_int, C.c_int64, _i64p, _i32p, _f64p, vp, vp, _f64p, _f64p,
                                 C.c_double, C.c_int, C.POINTER(C.c_double), vp, C.c_int]
        L.orc_newton.restype = C.c_int
        L.orc_newton.argtypes = [C.c_int64, _i64p, _i32p, _f64p, C.c_double, _f64p, _f64p, _f64p,
                                 _f64p, C.c_double, C.c_int, C.c_double, C.c_int, _i32p, _f64p,
                                 C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_dot_parts.restype = None
        L.orc_set_dot_parts.argtypes = [C.c_int]
        L.orc_set_row_reverse.restype = None
        L.orc_set_row_reverse.argtypes = [C.c_int]
        _lib = L
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_dot_parts(parts):
    """Split every dot product of the Krylov solvers into `parts` contiguous partial sums (the image of
    `parts` MPI ranks), independently of the number of threads; 0 = one part per thread (default)."""
    lib().orc_set_dot_parts(int(parts))


def set_row_reverse(on):
    """Sum every row of the sparse matrix-vector products right to left: a rounding-level perturbation of the
    operator apply, to measure the sensitivity of iteration counts to it."""
    lib().orc_set_row_reverse(int(bool(on)))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleProblem:
    """The reference's objects for one mesh + field set, in the reference's layout.

    coords: (N,3) float64; cells: (C,3|4) int32 (0-based); thickness, V: (N,) or scalars.
    mvp: ("explicit", A (N,3))  -> vector_field::explicit_values (a5)
         ("constcurl", b(3), u(3) or None) -> vector_field::constantCurl (a6)
    """

    def __init__(self, coords, cells, mvp, V=-1.0, thickness=1.0, nthreads=1):
        L = lib()
        self.nt = int(nthreads)
        self.coords = np.ascontiguousarray(coords, np.float64)
        self.cells = np.ascontiguousarray(cells, np.int32)
        self.N = self.coords.shape[0]
        self.dim = self.cells.shape[1] - 1
        nc = self.cells.shape[0]
        ne = 6 if self.dim == 3 else 3
        # a1
        E = L.orc_build_edges(self.dim, nc, self.cells, None, None)
        self.E = int(E)
        self.edges = np.empty((E, 2), np.int32)
        self.cell_edges = np.empty((nc, ne), np.int32)
        L.orc_build_edges(self.dim, nc, self.cells, _ptr(self.edges), _ptr(self.cell_edges))
        # a2
        self.length = np.empty(E)
        self.covolume = np.empty(E)
        rc = L.orc_edge_data(self.dim, self.N, self.coords, nc, self.cell_edges, E, self.edges,
                             self.length, self.covolume)
        if rc:
            raise RuntimeError("Illegal mesh: tetrahedron too flat (mesh_tetra.cpp:361)")
        # a3
        self.cv = np.empty(self.N)
        rc = L.orc_control_volumes(self.dim, self.N, self.coords, nc, self.cells, self.cv)
        if rc:
            raise RuntimeError("degenerate cell (mesh_tetra.cpp:393 / mesh.cpp:913)")
        # a7 scalar fields
        self.thickness = np.ascontiguousarray(np.broadcast_to(np.float64(thickness), (self.N,)))
        self.V = np.ascontiguousarray(np.broadcast_to(np.float64(V), (self.N,)))
        # a5 / a6
        self.mvp_kind = mvp[0]
        if mvp[0] == "explicit":
            A = np.ascontiguousarray(mvp[1], np.float64)
            self.cache = np.empty(E)
            L.orc_edge_cache_explicit(self.coords, A, E, self.edges, self.cache)
        elif mvp[0] == "constcurl":
            self.b = np.ascontiguousarray(mvp[1], np.float64)
            self.u = None if mvp[2] is None else np.ascontiguousarray(mvp[2], np.float64)
            if self.b @ self.b != 1.0:
                raise ValueError("Curl vector not normalized")  # constant_curl.cpp:35-38
            if self.u is not None and self.u @ self.u != 1.0:
                raise ValueError("Rotation vector not normalized")  # :40-44
            self.cache3 = np.empty((E, 3))
            L.orc_edge_cache_constcurl(self.coords, E, self.edges, self.cache3)
        else:
            raise ValueError(mvp[0])
        # a8
        self.alpha = np.empty(E)
        L.orc_alpha_cache(E, self.edges, self.length, self.covolume, self.thickness, self.alpha)
        # a4
        nnz = L.orc_build_complex_graph(self.N, E, self.edges, None, None)
        self.rowptr = np.empty(2 * self.N + 1, np.int64)
        self.cols = np.empty(nnz, np.int32)
        L.orc_build_complex_graph(self.N, E, self.edges, _ptr(self.rowptr), _ptr(self.cols))
        self.vals = np.zeros(nnz)
        self.dvals = None
        self.d0 = None
        self.d1b = None

    # -- fields ---------------------------------------------------------------
    def edge_projection(self, mu, theta=0.0, dname=None):
        """get_edge_projection (and get_d_edge_projection_dp(., dname)) for all edges."""
        L = lib()
        if self.mvp_kind == "explicit":
            a = mu * self.cache
            if dname is None:
                return a, None
            da = self.cache.copy() if dname == "mu" else np.zeros(self.E)
            return a, da
        rb = np.empty(3)
        drb = np.empty(3)
        L.orc_constcurl_rotate(self.b, _ptr(self.u), float(theta), rb, drb)
        a = np.empty(self.E)
        dmu = np.empty(self.E)
        dth = np.empty(self.E)
        L.orc_constcurl_projection(rb, drb, float(mu), self.E, self.cache3, _ptr(a), _ptr(dmu),
                                   _ptr(dth))
        if dname is None:
            return a, None
        if dname == "mu":
            return a, dmu
        if dname == "theta":
            return a, dth
        raise ValueError('Illegal parameter "%s".' % dname)  # constant_curl.cpp:135-139

    # -- a9 / a10 ---------------------------------------------------------------
    def keo_fill(self, mu, theta=0.0, nthreads=None):
        a, _ = self.edge_projection(mu, theta)
        rc = lib().orc_keo_fill(0, self.N, self.E, self.edges, self.rowptr, self.cols, self.alpha,
                                np.ascontiguousarray(a), None, self.vals,
                                self.nt if nthreads is None else nthreads)
        assert rc == 0
        return self.vals

    def dkeo_fill(self, mu, theta=0.0, dname="mu"):
        a, da = self.edge_projection(mu, theta, dname)
        if self.dvals is None:
            self.dvals = np.zeros_like(self.vals)
        rc = lib().orc_keo_fill(1, self.N, self.E, self.edges, self.rowptr, self.cols, self.alpha,
                                np.ascontiguousarray(a), _ptr(np.ascontiguousarray(da)), self.dvals,
                                self.nt)
        assert rc == 0
        return self.dvals

    def keoreg_fill(self, mu, g, x, theta=0.0):
        vals = self.keo_fill(mu, theta).copy()
        rc = lib().orc_keoreg_add_diag(self.N, self.rowptr, self.cols, vals, float(g), self.cv,
                                       self.thickness, np.ascontiguousarray(x))
        assert rc == 0
        return vals

    # -- operators ---------------------------------------------------------------
    def csr_apply(self, vals, x):
        y = np.empty(2 * self.N)
        lib().orc_csr_apply(2 * self.N, self.rowptr, self.cols, vals, np.ascontiguousarray(x), y,
                            self.nt)
        return y

    def keo_apply(self, x):
        return self.csr_apply(self.vals, x)

    def jac_rebuild(self, g, x, V=None):
        self.d0 = np.empty(2 * self.N)
        self.d1b = np.empty(self.N)
        lib().orc_jac_diags(self.N, float(g), self.cv, self.thickness,
                            self.V if V is None else V, np.ascontiguousarray(x), self.d0, self.d1b,
                            self.nt)

    def jac_apply(self, X):
        X = np.ascontiguousarray(X, np.float64)
        nvec = 1 if X.ndim == 1 else X.shape[0]  # rows of a C array = columns of a col-major MV
        Y = np.empty_like(X)
        lib().orc_jac_apply(self.N, self.rowptr, self.cols, self.vals, self.d0, self.d1b, nvec,
                            X.reshape(-1), 2 * self.N, Y.reshape(-1), 2 * self.N, self.nt)
        return Y

    def compute_f(self, g, x, V=None):
        f = np.empty(2 * self.N)
        lib().orc_compute_f(self.N, self.rowptr, self.cols, self.vals, float(g), self.cv,
                            self.thickness, self.V if V is None else V, np.ascontiguousarray(x), f,
                            self.nt)
        return f

    def compute_dfdp(self, x, is_g, dvdp=None):
        f = np.empty(2 * self.N)
        lib().orc_compute_dfdp(self.N, self.rowptr, self.cols, self.dvals, int(is_g), self.cv,
                               self.thickness, _ptr(dvdp), np.ascontiguousarray(x), f)
        return f

    def krylov(self, b, tol, maxit, solver="minres", jacobian=True, history=False):
        x = np.empty(2 * self.N)
        rr = C.c_double(0.0)
        hist = np.full(maxit + 1, np.nan) if history else None
        it = lib().orc_krylov(0 if solver == "minres" else 1, self.N, self.rowptr, self.cols,
                              self.vals, _ptr(self.d0) if jacobian else None,
                              _ptr(self.d1b) if jacobian else None, np.ascontiguousarray(b), x,
                              float(tol), int(maxit), C.byref(rr), _ptr(hist), self.nt)
        if history:
            return x, int(it), rr.value, hist[:it + 1]
        return x, int(it), rr.value

    def newton(self, g, x0, nl_tol=1e-8, nl_maxit=20, lin_tol=1e-10, lin_maxit=1000):
        x = np.array(x0, np.float64)
        lin = np.zeros(nl_maxit, np.int32)
        fn = np.full(nl_maxit + 1, np.nan)
        k = lib().orc_newton(self.N, self.rowptr, self.cols, self.vals, float(g), self.cv,
                             self.thickness, self.V, x, nl_tol, nl_maxit, lin_tol, lin_maxit, lin,
                             fn, self.nt)
        return x, int(k), lin[:k].copy(), fn[:k + 1].copy()

    # -- model-evaluator scalars as the reference's commented code states them
    # (model_evaluator_nls.cpp:699-770; the live code returns 0.0)
    def inner_product(self, phi, psi):
        return float(np.sum(self.cv * (phi[0::2] * psi[0::2] + phi[1::2] * psi[1::2])) / self.cv.sum())

    def gibbs_energy(self, psi):
        a = psi[0::2] ** 2 + psi[1::2] ** 2
        return float(-np.sum(self.cv * a * a) / self.cv.sum())

    def continuation(self, g, pname, p0, dp, nsteps, psi0, theta=0.0, nl_tol=1e-8, nl_maxit=20,
                     lin_tol=1e-10, lin_maxit=1000):
        """Natural continuation in mu (pname == "mu") with tangent predictor -- the restatement
        the device driver nosh_continuation is compared against.  [LOCA is not in the reference
        tree: unpinned.]"""
        assert pname == "mu"
        x = np.array(psi0, np.float64)
        recs = []
        mu = p0
        for k in range(nsteps + 1):
            pred_its = 0
            if k > 0:
                self.keo_fill(mu, theta)
                self.jac_rebuild(g, x)
                self.dkeo_fill(mu, theta, "mu")
                dF = self.compute_dfdp(x, False, np.zeros(self.N))
                t, pred_its, _ = self.krylov(-dF, lin_tol, lin_maxit)
                x = x + dp * t
            mu = p0 + k * dp
            self.keo_fill(mu, theta)
            x, steps, lin, fn = self.newton(g, x, nl_tol, nl_maxit, lin_tol, lin_maxit)
            recs.append(dict(step=k, param=mu, newton_steps=steps, linear_iterations=int(lin.sum()),
                             predictor_linear_iterations=pred_its, fnorm=float(fn[-1]),
                             gibbs_energy=self.gibbs_energy(x), norm=np.sqrt(self.inner_product(x, x))))
            if not fn[-1] < nl_tol:
                break
        return x, recs

    # -- complex block view (for entry-wise parity with the device block-CSR) ----
    def complex_blocks(self, vals):
        """Return (rowptr_v (N+1), colv, K complex) of the complex N x N matrix the real
        2N x 2N matrix represents: K_ij = vals[(2i,2j)] + i * vals[(2i+1,2j)]."""
        w = (self.rowptr[1::2] - self.rowptr[0:-1:2]) // 2
        rp = np.zeros(self.N + 1, np.int64)
        np.cumsum(w, out=rp[1:])
        cols_top = np.concatenate([self.cols[self.rowptr[2 * i]:self.rowptr[2 * i + 1]:2]
                                   for i in range(self.N)]) // 2 if self.N < 200000 else None
        # vectorised extraction
        starts_top = self.rowptr[0:-1:2]
        starts_bot = self.rowptr[1::2]
        idx_top = np.repeat(starts_top - 2 * rp[:-1], w) + 2 * np.arange(rp[-1])
        idx_bot = np.repeat(starts_bot - 2 * rp[:-1], w) + 2 * np.arange(rp[-1])
        colv = self.cols[idx_top] // 2
        if cols_top is not None:
            assert np.array_equal(colv, cols_top)
        K = vals[idx_top] + 1j * vals[idx_bot]
        return rp, colv.astype(np.int32), K
