"""CPU restatement of the reference's generic finite-volume assembly for REAL scalar problems.  TEST
INFRASTRUCTURE ONLY (only tests/ may import it).

  fill(...)      nosh::fvm_matrix::fill           src/fvm_matrix.hpp:44-70: setAllToScalar(0) / rhs = 0, then
                 add_edge_contributions_ (:75-150), add_vertex_contributions_ (:152-180),
                 add_domain_boundary_contributions_ (:182-211, same shape as the vertex loop) and apply_dbcs_
                 (:213-250: the ROW becomes the unit row, rhs = bc value; the column stays)
  operator_apply nosh::fvm_operator::apply        src/fvm_operator.hpp:45-94: edge cores, vertex cores, Dirichlet last
  boundary_vertices  mesh::compute_boundary_skin_ / compute_boundary_vertices_   src/mesh.cpp:76-131 (MOAB Skinner:
                 the faces that belong to one cell only; here by counting sorted faces)

The cores are the ones nfc generates for examples/poisson/poisson.py and examples/bratu/bratu.py:
  integrate(-n_dot_grad(u), dS)   edge core  covolume * -(u1 - u0)/edge_length for vertex 0 and its mirror for
                                  vertex 1 (nfc/nfc/discretize_edge_integral.py:62,108-114)
                                  => lhs = alpha [[1,-1],[-1,1]], alpha = covolume / edge_length
  integrate(f(x), dV)             vertex core  control_volume * f(x_k) (nfc/nfc/integral_vertex.py:164); affine parts
                                  go to the right-hand side with the opposite sign (:114)
The reference itself cannot be run (Trilinos / MOAB absent): PARITY of this file is pinned by the identities the
tests check (row sums of the Laplacian vanish, symmetric positive semi-definite, sum of control volumes, the
discrete solution of a problem with a known linear solution) rather than by reference golden numbers -- the
reference's tests hold none for these examples.
"""
import numpy as np
import scipy.sparse as sp


def boundary_vertices(cells, n_vertices):
    """flags[k] = 1 if vertex k lies on a face (3D) / edge (2D) that belongs to exactly one cell."""
    cells = np.asarray(cells, np.int64)
    nvc = cells.shape[1]
    faces = np.concatenate([np.delete(cells, i, axis=1) for i in range(nvc)], axis=0)
    faces.sort(axis=1)
    uniq, cnt = np.unique(faces, axis=0, return_counts=True)
    flags = np.zeros(n_vertices, np.int32)
    flags[np.unique(uniq[cnt == 1])] = 1
    return flags


def fill(P, edge_coeff=None, edge_lhs=None, edge_rhs=None, vertex_lhs=None, vertex_rhs=None, dirichlet_mask=None,
         dirichlet_values=None):
    """P: OracleProblem (edges, covolume, edge lengths).  Returns (A csr, rhs)."""
    N, E = P.N, P.E
    i, j = P.edges[:, 0].astype(np.int64), P.edges[:, 1].astype(np.int64)
    if edge_lhs is None:
        a = P.covolume / P.length * (1.0 if edge_coeff is None else np.asarray(edge_coeff))
        lhs = np.stack([a, -a, -a, a], 1)
    else:
        lhs = np.asarray(edge_lhs, np.float64).reshape(E, 4)
    rows = np.concatenate([i, i, j, j])
    cols = np.concatenate([i, j, i, j])
    A = sp.coo_matrix((np.concatenate([lhs[:, 0], lhs[:, 1], lhs[:, 2], lhs[:, 3]]), (rows, cols)), shape=(N, N)).tocsr()
    rhs = np.zeros(N)
    if edge_rhs is not None:
        er = np.asarray(edge_rhs, np.float64).reshape(E, 2)
        np.add.at(rhs, i, er[:, 0])
        np.add.at(rhs, j, er[:, 1])
    if vertex_lhs is not None:
        A = A + sp.diags(np.asarray(vertex_lhs, np.float64))
    if vertex_rhs is not None:
        rhs = rhs + np.asarray(vertex_rhs, np.float64)
    A = A.tocsr()
    A.sort_indices()
    if dirichlet_mask is not None:
        # getGlobalRowCopy / fill(vals, 0) / vals[diag] = 1 / replaceGlobalValues: the pattern stays (:226-243)
        m = np.asarray(dirichlet_mask) != 0
        row = np.repeat(np.arange(N), np.diff(A.indptr))
        A.data[m[row]] = 0.0
        A.data[m[row] & (A.indices == row)] = 1.0
        rhs = np.where(m, np.asarray(dirichlet_values, np.float64), rhs)
    return A, rhs


def operator_apply(A, cv, x, vertex_core=0, alpha=0.0, u0=None, dirichlet_mask=None, dirichlet_kind=0,
                   dirichlet_values=None):
    y = A @ x if A is not None else np.zeros_like(x)
    if vertex_core == 1:
        y = y - alpha * cv * np.exp(x)
    elif vertex_core == 2:
        y = y - alpha * cv * np.exp(u0) * x
    if dirichlet_mask is not None and dirichlet_kind:
        m = np.asarray(dirichlet_mask) != 0
        if dirichlet_kind == 1:
            y = np.where(m, x, y)
        elif dirichlet_kind == 2:
            y = np.where(m, 0.0, y)
        else:
            y = np.where(m, x - np.asarray(dirichlet_values), y)
    return y


def cg(A, b, tol, maxit, x0=None):
    """oracle/nosh_oracle.cpp:cg (Belos PseudoBlockCG organisation) with an initial guess.  With Dirichlet rows
    the matrix is not symmetric (rows eliminated, columns kept); started from the Dirichlet lift (x0 = g on the
    Dirichlet vertices, 0 elsewhere) the residual and every search direction vanish on those rows, and CG runs on
    the symmetric interior block.  (examples/poisson/poisson.cpp:28 starts from 0 and relies on the MueLu
    preconditioner; plain CG does not converge from there -- measured: relative residual 0.2 after 2000 steps.)"""
    x = np.zeros_like(b) if x0 is None else np.array(x0, np.float64)
    r = b - A @ x
    p = r.copy()
    rho = r @ r
    r0 = np.sqrt(rho)
    it = 0
    if r0 == 0.0:
        return x, 0, 0.0
    while it < maxit and not np.sqrt(rho) / r0 <= tol:
        it += 1
        ap = A @ p
        al = rho / (p @ ap)
        x += al * p
        r -= al * ap
        rho_new = r @ r
        p = r + (rho_new / rho) * p
        rho = rho_new
    return x, it, np.sqrt(rho) / r0
