"""Python front-end over the C ABI, used by the tests, bench.py and as a usage example.

Vectors may be numpy arrays (host: staged H2D/D2H inside every call) or anything with a
CUDA ``data_ptr()`` (torch tensors: used in place).  Parameters are plain dicts, the image
of the reference's std::map<std::string,double>.

Error mapping (reference behaviour in parentheses):
  NOSH_EKEY   -> KeyError      (std::map::at -> std::out_of_range)
  NOSH_EINVAL -> ValueError    (TEUCHOS_TEST_FOR_EXCEPT_MSG -> std::logic_error)
  NOSH_EMESH  -> RuntimeError  ("Illegal mesh: tetrahedron too flat")
  others      -> RuntimeError
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (AMG_REUSE_FULL, AMG_REUSE_NONE, ARC_HIT_BOUND, ARC_SCALING, LAYOUT_CSR, LAYOUT_SELL32, MAT_DKEO, MAT_KEO,  # noqa: F401
                   NO_TRANS, OP_JACOBIAN, OP_KEO, OP_KEOREG, PREC_KEOREG_AMG, PREC_NONE, AmgInfo,
                   ArclengthOptions, ArclengthStep, ContinuationStep, KrylovResult, MeshInfo, NewtonResult)


class NoshError(RuntimeError):
    pass


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    if isinstance(a, int):
        return C.c_void_p(a)
    raise TypeError(type(a))


def _params(p):
    names = (C.c_char_p * len(p))(*[k.encode() for k in p])
    vals = np.array([float(v) for v in p.values()], np.float64)
    return len(p), names, vals


def partition_range(n_global, nranks, rank, group_vertices=65536):
    """Host-only: (begin, end, group_used) of the vertex range `rank` owns."""
    b, e, g = C.c_int64(), C.c_int64(), C.c_int64()
    rc = _lib.lib().nosh_partition_range(int(n_global), int(nranks), int(rank), int(group_vertices),
                                         C.byref(b), C.byref(e), C.byref(g))
    if rc != 0:
        raise ValueError("nosh_partition_range: bad arguments")
    return b.value, e.value, g.value


def _io_check(rc):
    if rc == 0:
        return
    msg = (_lib.lib().nosh_meshfile_last_error() or b"").decode()
    if rc == _lib.NOSH_EKEY:
        raise KeyError(msg)
    if rc == _lib.NOSH_EINVAL:
        raise ValueError(msg)
    raise NoshError("status %d: %s" % (rc, msg))


def read_mesh(path):
    """nosh::read + the vertex tags (host only): returns (coords (N,3), cells (C,dim+1) int32, fields dict).
    Legacy VTK unstructured grids (ASCII or BINARY), Exodus II files in the netCDF classic container and gmsh MSH
    files (ASCII 2.x / 4.1); complex states come back as (N,2) arrays."""
    L = _lib.lib()
    h = C.c_void_p()
    _io_check(L.nosh_meshfile_read(str(path).encode(), C.byref(h)))
    try:
        dim, nv, nc, nf = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32()
        _io_check(L.nosh_meshfile_info(h, C.byref(dim), C.byref(nv), C.byref(nc), C.byref(nf)))
        coords = np.empty((nv.value, 3))
        cells = np.empty((nc.value, dim.value + 1), np.int32)
        _io_check(L.nosh_meshfile_get(h, _ptr(coords), _ptr(cells)))
        fields = {}
        for i in range(nf.value):
            name, ncomp = C.c_char_p(), C.c_int32()
            _io_check(L.nosh_meshfile_field_name(h, i, C.byref(name), C.byref(ncomp)))
            v = np.empty((nv.value, ncomp.value))
            _io_check(L.nosh_meshfile_get_field(h, name.value, None, _ptr(v)))
            fields[name.value.decode()] = v[:, 0].copy() if ncomp.value == 1 else v
        return coords, cells, fields
    finally:
        L.nosh_meshfile_free(h)


def write_mesh(path, coords, cells, fields=None, binary=False):
    """mesh::write (the outNNNN dumps): legacy VTK with the given vertex tags.  A state vector in the
    interleaved (re,im) layout is passed as psi.reshape(-1, 2)."""
    coords = np.ascontiguousarray(coords, np.float64)
    cells = np.ascontiguousarray(cells, np.int32)
    fields = fields or {}
    names = (C.c_char_p * max(1, len(fields)))(*[k.encode() for k in fields])
    arrs = [np.ascontiguousarray(np.asarray(v, np.float64).reshape(coords.shape[0], -1)) for v in fields.values()]
    ncomps = np.array([a.shape[1] for a in arrs] or [0], np.int32)
    ptrs = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
    _io_check(_lib.lib().nosh_meshfile_write(str(path).encode(), cells.shape[1] - 1, coords.shape[0], _ptr(coords),
                                             cells.shape[0], _ptr(cells), len(arrs), names, _ptr(ncomps), ptrs,
                                             1 if binary else 0))


def morton_order(coords):
    """perm with perm[i] = old id of the vertex that gets new id i (spatially local numbering)."""
    coords = np.ascontiguousarray(coords, np.float64)
    perm = np.empty(coords.shape[0], np.int64)
    _io_check(_lib.lib().nosh_morton_order(coords.shape[0], _ptr(coords), _ptr(perm)))
    return perm


def renumber(coords, cells, perm, fields=None):
    """apply a vertex permutation (new <- old = perm) to a mesh and its vertex fields"""
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    out_fields = {k: np.asarray(v)[perm] for k, v in (fields or {}).items()}
    return coords[perm], inv[cells].astype(np.int32), out_fields


class Context:
    """One nosh_ctx: one process, one GPU."""

    def __init__(self, device=0, stream=None, layout=None, group_vertices=None):
        self.L = _lib.lib()
        h = C.c_void_p()
        rc = self.L.nosh_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise NoshError("nosh_ctx_create failed with status %d (no usable CUDA device? there is "
                            "no CPU fallback)" % rc)
        self.h = h
        if layout is not None:
            self._ck(self.L.nosh_ctx_set_layout(self.h, int(layout)))
        if group_vertices is not None:
            self._ck(self.L.nosh_ctx_set_group_vertices(self.h, int(group_vertices)))

    def close(self):
        if getattr(self, "h", None):
            self.L.nosh_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc == 0:
            return
        msg = self.L.nosh_last_error(self.h).decode()
        if rc == _lib.NOSH_EKEY:
            raise KeyError(msg)
        if rc == _lib.NOSH_EINVAL:
            raise ValueError(msg)
        raise NoshError("status %d: %s" % (rc, msg))

    # ---- comm -------------------------------------------------------------------------
    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        rc = _lib.lib().nosh_comm_unique_id(buf)
        if rc != 0:
            raise NoshError("nosh_comm_unique_id failed (%d)" % rc)
        return bytes(buf)

    def comm_init(self, uid, rank, nranks):
        b = (C.c_char * 128).from_buffer_copy(uid)
        self._ck(self.L.nosh_ctx_comm_init(self.h, b, int(rank), int(nranks)))

    def comm_init_host(self, rank, nranks, allgather):
        """Use the caller's communicator for set-up (no NCCL communicator inside the library):
        allgather(bytes) -> list of every rank's bytes, in rank order.  All data-path exchange then goes
        over CUDA-IPC peer memory."""
        def _cb(user, send, recv, nbytes):
            try:
                parts = allgather(C.string_at(send, nbytes))
                if len(parts) != nranks or any(len(q) != nbytes for q in parts):
                    return 2
                C.memmove(recv, b"".join(parts), nbytes * nranks)
                return 0
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._ag_cb = _lib.ALLGATHER_FN(_cb)      # keep the trampoline alive as long as the ctx
        self._ck(self.L.nosh_ctx_comm_init_host(self.h, int(rank), int(nranks), self._ag_cb, None))

    def comm_init_torch(self, group=None):
        """comm_init_host over torch.distributed (any backend that can all_gather_object)."""
        import torch.distributed as dist

        def ag(b):
            out = [None] * dist.get_world_size(group)
            dist.all_gather_object(out, b, group=group)
            return out
        self.comm_init_host(dist.get_rank(group), dist.get_world_size(group), ag)

    def stat(self, key):
        v = C.c_double()
        self._ck(self.L.nosh_ctx_get_stat(self.h, key.encode(), C.byref(v)))
        return v.value

    def prefetch(self, host_vector):
        """start the H2D copy of a host state vector now; the next call given the same array waits for it"""
        self._ck(self.L.nosh_prefetch(self.h, _ptr(host_vector)))

    def set_async_output(self, enabled):
        """results in host memory are copied back on the copy stream; valid after synchronize()"""
        self._ck(self.L.nosh_ctx_set_async_output(self.h, int(bool(enabled))))

    def stats(self, prefix=""):
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.L.nosh_ctx_list_stats(self.h, buf, len(buf)))
        out = {}
        for ln in buf.value.decode().splitlines():
            k, _, v = ln.rpartition("=")
            if k.startswith(prefix):
                out[k] = float(v)
        return out

    def synchronize(self):
        self._ck(self.L.nosh_ctx_synchronize(self.h))

    # ---- mesh -------------------------------------------------------------------------
    def mesh_set(self, coords, cells):
        coords = np.ascontiguousarray(coords, np.float64)
        cells = np.ascontiguousarray(cells, np.int32)
        self._ck(self.L.nosh_mesh_set(self.h, cells.shape[1] - 1, coords.shape[0], _ptr(coords),
                                      cells.shape[0], _ptr(cells)))
        return self.info()

    def mesh_set_local(self, n_global, vertex_gids, coords, cells):
        """Partitioned ingestion: this rank's vertices (global ids + coordinates) and cells (indices into them)."""
        gids = np.ascontiguousarray(vertex_gids, np.int64)
        coords = np.ascontiguousarray(coords, np.float64)
        cells = np.ascontiguousarray(cells, np.int32)
        self._ck(self.L.nosh_mesh_set_local(self.h, cells.shape[1] - 1, int(n_global), gids.shape[0], _ptr(gids),
                                            _ptr(coords), cells.shape[0], _ptr(cells)))
        return self.info()

    @staticmethod
    def local_part(coords, cells, begin, end):
        """Host helper: the part of a global mesh a rank owning the vertex range [begin, end) has to pass to
        mesh_set_local -- the cells touching an owned vertex, their vertices, local connectivity."""
        cells = np.asarray(cells)
        keep = ((cells >= begin) & (cells < end)).any(axis=1)
        sub = cells[keep]
        gids, inv = np.unique(sub, return_inverse=True)
        return gids.astype(np.int64), np.asarray(coords)[gids], inv.reshape(sub.shape).astype(np.int32)

    def mesh_tetgrid(self, nx, ny=None, nz=None, lo=(-5.0, -5.0, -5.0), hi=(5.0, 5.0, 5.0),
                     jitter=0.2, seed=1234):
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        lo = np.array(lo, np.float64)
        hi = np.array(hi, np.float64)
        self._ck(self.L.nosh_mesh_tetgrid(self.h, nx, ny, nz, _ptr(lo), _ptr(hi), float(jitter),
                                          int(seed)))
        return self.info()

    def info(self):
        mi = MeshInfo()
        self._ck(self.L.nosh_mesh_info(self.h, C.byref(mi)))
        self._info = mi
        return mi

    @property
    def n_owned(self):
        return int(self._info.n_owned)

    def local_gids(self):
        mi = self._info
        g = np.empty(mi.n_owned + mi.n_ghost, np.int64)
        self._ck(self.L.nosh_mesh_local_gids(self.h, _ptr(g)))
        return g

    def coords(self):
        mi = self._info
        c = np.empty((mi.n_owned + mi.n_ghost, 3))
        self._ck(self.L.nosh_mesh_get_coords(self.h, _ptr(c)))
        return c

    def cells(self):
        mi = self._info
        c = np.empty((mi.n_cells, mi.dim + 1), np.int32)
        self._ck(self.L.nosh_mesh_get_cells(self.h, _ptr(c)))
        return c

    def edges(self):
        mi = self._info
        e = np.empty((mi.n_edges, 2), np.int32)
        ln = np.empty(mi.n_edges)
        cov = np.empty(mi.n_edges)
        self._ck(self.L.nosh_mesh_get_edges(self.h, _ptr(e), _ptr(ln), _ptr(cov)))
        return e, ln, cov

    def control_volumes(self):
        cv = np.empty(self._info.n_owned)
        self._ck(self.L.nosh_mesh_get_control_volumes(self.h, _ptr(cv)))
        return cv

    # ---- fields -----------------------------------------------------------------------
    def set_thickness(self, values=None, c=1.0):
        v = None if values is None else np.ascontiguousarray(values, np.float64)
        self._ck(self.L.nosh_set_thickness(self.h, _ptr(v), float(c)))

    def set_potential_constant(self, c, param1_name=None):
        self._ck(self.L.nosh_set_potential_constant(self.h, float(c),
                                                    param1_name.encode() if param1_name else None))

    def set_potential_values(self, values):
        v = np.ascontiguousarray(values, np.float64)
        self._ck(self.L.nosh_set_potential_values(self.h, _ptr(v)))

    def set_mvp_explicit(self, A):
        A = np.ascontiguousarray(A, np.float64)
        self._ck(self.L.nosh_set_mvp_explicit(self.h, _ptr(A)))

    def set_mvp_explicit_curl(self, B):
        B = np.array(B, np.float64)
        self._ck(self.L.nosh_set_mvp_explicit_curl(self.h, _ptr(B)))

    def set_mvp_constcurl(self, b, u=None):
        b = np.array(b, np.float64)
        u = None if u is None else np.array(u, np.float64)
        self._ck(self.L.nosh_set_mvp_constcurl(self.h, _ptr(b), _ptr(u)))

    def alpha_cache(self):
        a = np.empty(self._info.n_edges)
        self._ck(self.L.nosh_get_alpha_cache(self.h, _ptr(a)))
        return a

    def edge_projection(self, params, dname=None):
        n, names, vals = _params(params)
        a = np.empty(self._info.n_edges)
        da = np.empty(self._info.n_edges) if dname else None
        self._ck(self.L.nosh_get_edge_projection(self.h, n, names, _ptr(vals),
                                                 dname.encode() if dname else None, _ptr(a), _ptr(da)))
        return a, da

    # ---- operators --------------------------------------------------------------------
    def keo_fill(self, params):
        n, names, vals = _params(params)
        self._ck(self.L.nosh_keo_fill(self.h, n, names, _ptr(vals)))

    def dkeo_fill(self, params, dname):
        n, names, vals = _params(params)
        self._ck(self.L.nosh_dkeo_fill(self.h, n, names, _ptr(vals), dname.encode()))

    def _out_like(self, x):
        if isinstance(x, np.ndarray):
            return np.empty_like(x)
        import torch
        return torch.empty_like(x)

    @staticmethod
    def _shape(X, n2):
        if isinstance(X, np.ndarray):
            nvec = 1 if X.ndim == 1 else X.shape[0]
        else:
            nvec = 1 if X.dim() == 1 else X.shape[0]
        return nvec, n2

    def matrix_apply(self, which, X, Y=None, mode=NO_TRANS, alpha=1.0, beta=0.0):
        """Rows of a 2-D X are the columns of the (column-major) multi-vector."""
        Y = self._out_like(X) if Y is None else Y
        nvec, ld = self._shape(X, 2 * self.n_owned)
        self._ck(self.L.nosh_matrix_apply(self.h, which, _ptr(X), ld, _ptr(Y), ld, nvec, mode,
                                          float(alpha), float(beta)))
        return Y

    def keo_apply(self, X, Y=None, **kw):
        return self.matrix_apply(MAT_KEO, X, Y, **kw)

    def dkeo_apply(self, X, Y=None, **kw):
        return self.matrix_apply(MAT_DKEO, X, Y, **kw)

    def block_csr(self, which=MAT_KEO, values=True):
        mi = self._info
        rp = np.empty(mi.n_owned + 1, np.int64)
        cols = np.empty(mi.n_blocks, np.int32)
        vals = np.empty(mi.n_blocks, np.complex128) if values else None
        self._ck(self.L.nosh_get_block_csr(self.h, which, _ptr(rp), _ptr(cols), _ptr(vals)))
        return rp, cols, vals

    def jac_rebuild(self, params, psi):
        n, names, vals = _params(params)
        self._ck(self.L.nosh_jac_rebuild(self.h, n, names, _ptr(vals), _ptr(psi)))

    def jac_apply(self, X, Y=None, mode=NO_TRANS, alpha=1.0, beta=0.0):
        Y = self._out_like(X) if Y is None else Y
        nvec, ld = self._shape(X, 2 * self.n_owned)
        self._ck(self.L.nosh_jac_apply(self.h, _ptr(X), ld, _ptr(Y), ld, nvec, mode, float(alpha),
                                       float(beta)))
        return Y

    def jac_diags(self):
        d0 = np.empty(2 * self.n_owned)
        d1 = np.empty(self.n_owned)
        self._ck(self.L.nosh_jac_get_diags(self.h, _ptr(d0), _ptr(d1)))
        return d0, d1

    def compute_f(self, params, psi, f=None):
        f = self._out_like(psi) if f is None else f
        n, names, vals = _params(params)
        self._ck(self.L.nosh_compute_f(self.h, n, names, _ptr(vals), _ptr(psi), _ptr(f)))
        return f

    def compute_dfdp(self, params, pname, psi, out=None):
        out = self._out_like(psi) if out is None else out
        n, names, vals = _params(params)
        self._ck(self.L.nosh_compute_dfdp(self.h, n, names, _ptr(vals), pname.encode(), _ptr(psi),
                                          _ptr(out)))
        return out

    def keoreg_rebuild(self, params, psi):
        n, names, vals = _params(params)
        self._ck(self.L.nosh_keoreg_rebuild(self.h, n, names, _ptr(vals), _ptr(psi)))

    def keoreg_matrix_apply(self, X, Y=None):
        Y = self._out_like(X) if Y is None else Y
        nvec, ld = self._shape(X, 2 * self.n_owned)
        self._ck(self.L.nosh_keoreg_matrix_apply(self.h, _ptr(X), ld, _ptr(Y), ld, nvec))
        return Y

    def keoreg_diags(self):
        d0 = np.empty(2 * self.n_owned)
        d1 = np.empty(self.n_owned)
        self._ck(self.L.nosh_keoreg_get_diags(self.h, _ptr(d0), _ptr(d1)))
        return d0, d1

    def keoreg_apply(self, X, Y=None, mode=NO_TRANS, alpha=1.0, beta=0.0):
        """keo_regularized::apply: one AMG V-cycle on the regularised KEO."""
        Y = self._out_like(X) if Y is None else Y
        nvec, ld = self._shape(X, 2 * self.n_owned)
        self._ck(self.L.nosh_keoreg_apply(self.h, _ptr(X), ld, _ptr(Y), ld, nvec, mode, alpha, beta))
        return Y

    # ---- AMG hierarchy ------------------------------------------------------------------
    def amg_set_options(self, degree=0, coarse_degree=0, coarse_max=0, max_levels=0, reuse=-1):
        self._ck(self.L.nosh_amg_set_options(self.h, int(degree), int(coarse_degree), int(coarse_max), int(max_levels),
                                             int(reuse)))

    def amg_setup(self):
        self._ck(self.L.nosh_amg_setup(self.h))

    def amg_info(self):
        info = AmgInfo()
        self._ck(self.L.nosh_amg_info(self.h, C.byref(info)))
        return info

    def amg_aggregates(self, level):
        n = self.amg_info().nodes[level]
        agg = np.empty(n, np.int32)
        self._ck(self.L.nosh_amg_get_aggregates(self.h, int(level), _ptr(agg)))
        return agg

    def _amg_csr(self, fn, level, nrows, nnz):
        rp = np.empty(nrows + 1, np.int64)
        cols = np.empty(nnz, np.int32)
        vals = np.empty((nnz, 2, 2))
        self._ck(fn(self.h, int(level), _ptr(rp), _ptr(cols), _ptr(vals)))
        return rp, cols, vals

    def amg_matrix(self, level):
        """block CSR (rowptr, cols, vals[nnz,2,2]) of the level matrix, level >= 1"""
        info = self.amg_info()
        return self._amg_csr(self.L.nosh_amg_get_matrix, level, info.nodes[level], info.blocks[level])

    def amg_prolongator(self, level):
        """block CSR of the prolongator from level+1 to level"""
        info = self.amg_info()
        return self._amg_csr(self.L.nosh_amg_get_prolongator, level, info.nodes[level],
                             info.p_blocks[level])

    def set_linear_solver(self, solver, gmres_restart=0):
        """Krylov solver of newton / continuation / continuation_arclength (SOLVER_MINRES, _CG, _GMRES)."""
        self._ck(self.L.nosh_ctx_set_linear_solver(self.h, int(solver), int(gmres_restart)))

    def set_preconditioner(self, prec):
        self._ck(self.L.nosh_ctx_set_preconditioner(self.h, int(prec)))

    # ---- reductions / solvers -----------------------------------------------------------
    def dot(self, x, y):
        r = C.c_double()
        self._ck(self.L.nosh_dot(self.h, _ptr(x), _ptr(y), C.byref(r)))
        return r.value

    def norm2(self, x):
        r = C.c_double()
        self._ck(self.L.nosh_norm2(self.h, _ptr(x), C.byref(r)))
        return r.value

    def _krylov(self, fn, fn_prec, op, prec, b, x, tol, maxit, history):
        x = self._out_like(b) if x is None else x
        res = KrylovResult()
        hist = np.full(maxit + 1, np.nan) if history else None
        if prec == PREC_NONE:
            self._ck(fn(self.h, op, _ptr(b), _ptr(x), float(tol), int(maxit), C.byref(res), _ptr(hist)))
        else:
            self._ck(fn_prec(self.h, op, int(prec), _ptr(b), _ptr(x), float(tol), int(maxit),
                             C.byref(res), _ptr(hist)))
        if history:
            return x, res, hist[:res.iterations + 1]
        return x, res

    def minres(self, b, x=None, op=OP_JACOBIAN, tol=1e-10, maxit=1000, history=False, prec=PREC_NONE):
        return self._krylov(self.L.nosh_minres, self.L.nosh_minres_prec, op, prec, b, x, tol, maxit,
                            history)

    def cg(self, b, x=None, op=OP_JACOBIAN, tol=1e-10, maxit=1000, history=False, prec=PREC_NONE):
        return self._krylov(self.L.nosh_cg, self.L.nosh_cg_prec, op, prec, b, x, tol, maxit, history)

    def gmres(self, b, x=None, op=OP_JACOBIAN, tol=1e-10, maxit=1000, restart=300, history=False,
              prec=PREC_NONE):
        x = self._out_like(b) if x is None else x
        res = KrylovResult()
        hist = np.full(maxit + 1, np.nan) if history else None
        self._ck(self.L.nosh_gmres(self.h, op, int(prec), _ptr(b), _ptr(x), float(tol), int(maxit), int(restart),
                                   C.byref(res), _ptr(hist)))
        if history:
            return x, res, hist[:res.iterations + 1]
        return x, res

    def newton(self, params, psi, nl_tol=1e-8, nl_maxit=20, lin_tol=1e-10, lin_maxit=1000):
        """psi is updated in place.  Returns (result, lin_iters, fnorms)."""
        n, names, vals = _params(params)
        res = NewtonResult()
        lin = np.zeros(max(nl_maxit, 1), np.int32)
        fn = np.full(nl_maxit + 1, np.nan)
        self._ck(self.L.nosh_newton(self.h, n, names, _ptr(vals), _ptr(psi), float(nl_tol),
                                    int(nl_maxit), float(lin_tol), int(lin_maxit), C.byref(res),
                                    _ptr(lin), _ptr(fn)))
        return res, lin[:res.steps].copy(), fn[:res.steps + 1].copy()

    def inner_product(self, phi, psi):
        r = C.c_double()
        self._ck(self.L.nosh_inner_product(self.h, _ptr(phi), _ptr(psi), C.byref(r)))
        return r.value

    def gibbs_energy(self, psi):
        r = C.c_double()
        self._ck(self.L.nosh_gibbs_energy(self.h, _ptr(psi), C.byref(r)))
        return r.value

    def continuation(self, params, pname, dp, nsteps, psi, nl_tol=1e-8, nl_maxit=20, lin_tol=1e-10,
                     lin_maxit=1000):
        """Natural continuation in `pname`; psi is updated in place.  Returns the step records."""
        n, names, vals = _params(params)
        steps = (ContinuationStep * (nsteps + 1))()
        self._ck(self.L.nosh_continuation(self.h, n, names, _ptr(vals), pname.encode(), float(dp),
                                          int(nsteps), _ptr(psi), float(nl_tol), int(nl_maxit),
                                          float(lin_tol), int(lin_maxit), steps))
        return [s for s in steps if s.step >= 0]

    def continuation_arclength(self, params, pname, psi, initial_step_size=1e-3, min_step_size=1e-7,
                               max_step_size=1e-2, aggressiveness=2.0, max_steps=10, nl_tol=1e-8, nl_maxit=20,
                               lin_tol=1e-10, lin_maxit=1000, min_value=-100.0, max_value=100.0, scaling=False,
                               hit_bound=False, goal_contribution=0.0, max_contribution=0.0, min_scale=0.0,
                               initial_scale=0.0):
        """Pseudo-arclength continuation (defaults: the LOCA settings of examples/conf.xml:35-75); psi is
        updated in place.  Returns the step records.  scaling / hit_bound: LOCA's "Enable Arc Length Scaling" and
        "Hit Continuation Bound" (both on in LOCA by default; see include/nosh_b200.h)."""
        n, names, vals = _params(params)
        flags = (ARC_SCALING if scaling else 0) | (ARC_HIT_BOUND if hit_bound else 0)
        opt = ArclengthOptions(float(initial_step_size), float(min_step_size), float(max_step_size),
                               float(aggressiveness), int(max_steps), int(nl_maxit), float(nl_tol),
                               float(lin_tol), int(lin_maxit), flags, float(min_value), float(max_value),
                               float(goal_contribution), float(max_contribution), float(min_scale),
                               float(initial_scale))
        steps = (ArclengthStep * (max_steps + 2))()
        nrec = C.c_int32(0)
        self._ck(self.L.nosh_continuation_arclength(self.h, n, names, _ptr(vals), pname.encode(), C.byref(opt),
                                                    _ptr(psi), steps, C.byref(nrec)))
        return [steps[i] for i in range(nrec.value)]

    def set_step_observer(self, fn):
        """fn(step, param, gibbs_energy, norm, psi) is called after every accepted continuation step with this rank's
        owned part of the solution (a numpy copy); returning True stops the run.  The image of the reference's
        observer (CSV row, src/observer.cpp:134-159) and continuation_data_saver (outNNNN dumps).  None removes it."""
        if fn is None:
            self._obs_cb = _lib.STEP_OBSERVER_FN(0)
            self._ck(self.L.nosh_ctx_set_step_observer(self.h, self._obs_cb, None))
            return

        def _cb(user, step, param, energy, norm, psi, n):
            try:
                arr = np.ctypeslib.as_array(psi, shape=(int(n),)).copy() if n > 0 else np.zeros(0)
                return 1 if fn(int(step), float(param), float(energy), float(norm), arr) else 0
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._obs_cb = _lib.STEP_OBSERVER_FN(_cb)
        self._ck(self.L.nosh_ctx_set_step_observer(self.h, self._obs_cb, None))

    @staticmethod
    def write_continuation_csv(path, steps, pname):
        """The CSV the reference's observer writes (src/observer.cpp:134-159, src/csv_writer.cpp)."""
        with open(path, "w") as f:
            f.write("(0) step,(1) %s,(2) Gibbs energy,(2) ||x||_2 scaled\n" % pname)
            for s in steps:
                f.write("%d,%.15e,%.15e,%.15e\n" % (s.step, s.param, s.gibbs_energy, s.norm))

    # ---- generic FVM matrix / operator (row f4) --------------------------------------------
    def boundary_vertices(self):
        f = np.empty(self.n_owned, np.int32)
        self._ck(self.L.nosh_mesh_boundary_vertices(self.h, _ptr(f)))
        return f

    def fvm_matrix_fill(self, edge_coeff=None, edge_lhs=None, edge_rhs=None, vertex_lhs=None, vertex_rhs=None,
                        dirichlet_mask=None, dirichlet_values=None):
        """fvm_matrix::fill; returns the right-hand side."""
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)  # noqa: E731
        ec, el, er, vl, vr, dv = map(f64, (edge_coeff, edge_lhs, edge_rhs, vertex_lhs, vertex_rhs, dirichlet_values))
        dm = None if dirichlet_mask is None else np.ascontiguousarray(dirichlet_mask, np.int32)
        rhs = np.empty(self.n_owned)
        self._ck(self.L.nosh_fvm_matrix_fill(self.h, _ptr(ec), _ptr(el), _ptr(er), _ptr(vl), _ptr(vr), _ptr(dm), _ptr(dv),
                                             _ptr(rhs)))
        return rhs

    def fvm_matrix_apply(self, x, y=None):
        y = self._out_like(x) if y is None else y
        self._ck(self.L.nosh_fvm_matrix_apply(self.h, _ptr(x), _ptr(y)))
        return y

    def fvm_csr(self):
        mi = self._info
        rp = np.empty(mi.n_owned + 1, np.int64)
        cols = np.empty(mi.n_blocks, np.int32)
        vals = np.empty(mi.n_blocks)
        self._ck(self.L.nosh_fvm_get_csr(self.h, _ptr(rp), _ptr(cols), _ptr(vals)))
        return rp, cols, vals

    def fvm_operator_apply(self, x, y=None, with_matrix=True, vertex_core=0, alpha=0.0, u0=None, dirichlet_mask=None,
                           dirichlet_kind=0, dirichlet_values=None):
        y = self._out_like(x) if y is None else y
        dm = None if dirichlet_mask is None else np.ascontiguousarray(dirichlet_mask, np.int32)
        dv = None if dirichlet_values is None else np.ascontiguousarray(dirichlet_values, np.float64)
        self._ck(self.L.nosh_fvm_operator_apply(self.h, int(bool(with_matrix)), int(vertex_core), float(alpha), _ptr(u0),
                                                _ptr(dm), int(dirichlet_kind), _ptr(dv), _ptr(x), _ptr(y)))
        return y

    def fvm_cg(self, b, x=None, tol=1e-10, maxit=1000):
        x = self._out_like(b) if x is None else x
        res = KrylovResult()
        self._ck(self.L.nosh_fvm_cg(self.h, _ptr(b), _ptr(x), float(tol), int(maxit), C.byref(res)))
        return x, res

    # ---- measurement ------------------------------------------------------------------
    def scratch_vector(self, slot):
        p = C.c_void_p()
        self._ck(self.L.nosh_scratch_vector(self.h, int(slot), C.byref(p)))
        return p.value

    def set_tuning(self, key, value):
        self._ck(self.L.nosh_ctx_set_tuning(self.h, key.encode(), int(value)))

    def launch_count(self):
        return int(self.L.nosh_launch_count(self.h))

    def timer_start(self):
        self._ck(self.L.nosh_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.nosh_timer_stop(self.h, C.byref(ms)))
        return ms.value
