"""Host (numpy) mesh fixtures and synthetic generators.  TEST INFRASTRUCTURE ONLY.

* ``rectanglesmall`` / ``cubesmall``: analytic reconstructions of the reference's
  two small fixtures (test/data/rectanglesmall.e.md5, cubesmall.e.md5 are only md5
  stubs; geometry inferred from test/mesh.cpp:40-49,88-97 and test/io.cpp:51-59,95-103,
  fields from examples/state-equippers/plain-gl:22-39) -- SURVEY.md section 8c.
* ``tetgrid`` / ``trigrid``: the synthetic inputs of SURVEY.md section 8d.  ``tetgrid``
  restates, in numpy, exactly what the device generator ``nosh_mesh_tetgrid`` does
  (same splitmix64 jitter keyed on the global vertex id, same rounding), so a parity
  test can compare the two bit for bit.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def jitter_unit(seed, gid, comp):
    """U(-1,1) keyed on (seed, global vertex id, component); partition independent."""
    with np.errstate(over="ignore"):
        k = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (
            np.asarray(gid, np.uint64) * np.uint64(3) + np.uint64(comp + 1))
    z = splitmix64(k)
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return 2.0 * u - 1.0


# The 6 Kuhn tetrahedra of the unit cube around the diagonal (0,0,0)-(1,1,1): one per
# permutation of the axes; corner bit masks (x=1, y=2, z=4).
KUHN = np.array([
    [0, 1, 3, 7],  # x, y, z
    [0, 1, 5, 7],  # x, z, y
    [0, 2, 3, 7],  # y, x, z
    [0, 2, 6, 7],  # y, z, x
    [0, 4, 5, 7],  # z, x, y
    [0, 4, 6, 7],  # z, y, x
], np.int64)


def tetgrid(nx, ny=None, nz=None, lo=(-5.0, -5.0, -5.0), hi=(5.0, 5.0, 5.0), jitter=0.2,
            seed=1234):
    """Structured nx*ny*nz vertex grid (x fastest), every hex cell split into 6 Kuhn tets;
    interior vertices displaced by jitter*h*U(-1,1) per component."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = (nx, ny, nz)
    gid = np.arange(nx * ny * nz, dtype=np.int64)
    ijk = (gid % nx, (gid // nx) % ny, gid // (nx * ny))
    interior = np.ones(gid.shape, bool)
    for d in range(3):
        interior &= (ijk[d] > 0) & (ijk[d] < n[d] - 1)
    coords = np.empty((gid.size, 3))
    for d in range(3):
        h = (hi[d] - lo[d]) / (n[d] - 1)
        base = lo[d] + ijk[d].astype(np.float64) * h
        disp = (jitter * h) * jitter_unit(seed, gid, d)
        coords[:, d] = np.where(interior, base + disp, base)
    # cells: hex cell c = (i,j,k), i<nx-1 ..., x fastest; 6 tets each, in KUHN order
    c = np.arange((nx - 1) * (ny - 1) * (nz - 1), dtype=np.int64)
    ci, cj, ck = c % (nx - 1), (c // (nx - 1)) % (ny - 1), c // ((nx - 1) * (ny - 1))
    v0 = ci + nx * (cj + ny * ck)
    off = np.array([(m & 1) + nx * (((m >> 1) & 1) + ny * ((m >> 2) & 1)) for m in range(8)])
    cells = (v0[:, None, None] + off[KUHN][None, :, :]).reshape(-1, 4).astype(np.int32)
    return coords, cells


# The 5-tetrahedron split of a cube (the reference's cubesmall fixture, test/mesh.cpp:88-97): four corner
# tets + the inner tet on the corners of odd local parity; cubes of odd (i+j+k) use the mirrored split so that
# the face diagonals of neighbouring cubes agree.  Corner bit masks as in KUHN.
FIVE_EVEN = np.array([[0, 1, 2, 4], [3, 1, 2, 7], [5, 1, 4, 7], [6, 2, 4, 7], [1, 2, 4, 7]], np.int64)
FIVE_ODD = np.array([[1, 0, 3, 5], [2, 0, 3, 6], [4, 0, 5, 6], [7, 3, 5, 6], [0, 3, 5, 6]], np.int64)


def tetgrid5(nx, ny=None, nz=None, lo=(-5.0, -5.0, -5.0), hi=(5.0, 5.0, 5.0), jitter=0.1, seed=1234):
    """Same vertices as ``tetgrid``, every hex cell split into 5 tets with alternating orientation.  Vertices
    of odd (i+j+k) own all 12 face diagonals around them (18 neighbours), the others have the 6 axis
    neighbours only: block rows of 19 and 7 entries alternate -- the stress case for a sliced-ELL layout
    (unstructured meshes have strongly varying valence; the Kuhn grid has 15 everywhere)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    coords, _ = tetgrid(nx, ny, nz, lo, hi, jitter, seed)
    c = np.arange((nx - 1) * (ny - 1) * (nz - 1), dtype=np.int64)
    ci, cj, ck = c % (nx - 1), (c // (nx - 1)) % (ny - 1), c // ((nx - 1) * (ny - 1))
    v0 = ci + nx * (cj + ny * ck)
    off = np.array([(m & 1) + nx * (((m >> 1) & 1) + ny * ((m >> 2) & 1)) for m in range(8)])
    odd = ((ci + cj + ck) & 1).astype(bool)
    loc = np.where(odd[:, None, None], off[FIVE_ODD][None, :, :], off[FIVE_EVEN][None, :, :])
    cells = (v0[:, None, None] + loc).reshape(-1, 4).astype(np.int32)
    return coords, cells


def trigrid(nx, ny, lo=(-5.0, -0.5), hi=(5.0, 0.5)):
    """Structured nx*ny triangle grid in the z=0 plane, every quad split along the same
    diagonal (config 1's scalable sibling of rectanglesmall)."""
    gid = np.arange(nx * ny)
    i, j = gid % nx, gid // nx
    coords = np.zeros((gid.size, 3))
    coords[:, 0] = lo[0] + i * ((hi[0] - lo[0]) / (nx - 1))
    coords[:, 1] = lo[1] + j * ((hi[1] - lo[1]) / (ny - 1))
    ci, cj = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="ij")
    v0 = (ci + nx * cj).T.ravel()
    t0 = np.stack([v0, v0 + 1, v0 + nx + 1], 1)
    t1 = np.stack([v0, v0 + nx + 1, v0 + nx], 1)
    cells = np.stack([t0, t1], 1).reshape(-1, 3).astype(np.int32)
    return coords, cells


def rectanglesmall():
    """4 vertices (+-5, +-0.5, 0), 2 triangles sharing the diagonal v0-v1."""
    coords = np.array([[5.0, 0.5, 0.0], [-5.0, -0.5, 0.0], [5.0, -0.5, 0.0], [-5.0, 0.5, 0.0]])
    cells = np.array([[0, 1, 2], [0, 3, 1]], np.int32)
    return coords, cells


def cubesmall():
    """8 vertices (+-.5, +-.5, +-5), 5-tet split (4 corner tets + 1 inner tet)."""
    coords = np.array([[sx * 0.5, sy * 0.5, sz * 5.0]
                       for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], np.float64)
    # vertex id = bx + 2*by + 4*bz; even-parity corners {0,3,5,6} form the inner tet
    cells = np.array([[0, 3, 5, 6],
                      [1, 0, 3, 5],
                      [2, 0, 3, 6],
                      [4, 0, 5, 6],
                      [7, 3, 5, 6]], np.int32)
    return coords, cells


def plain_gl_fields(coords, B=(0.0, 0.0, 1.0)):
    """examples/state-equippers/plain-gl:22-39: psi = 1+0i, V = -1, A = 0.5 B x X."""
    N = coords.shape[0]
    psi = np.zeros(2 * N)
    psi[0::2] = 1.0
    A = 0.5 * np.cross(np.asarray(B, np.float64)[None, :], coords)
    return psi, A


def random_state(N, seed=42):
    """Apply-timing state of SURVEY.md 8d: x_k = rho (cos xi, sin xi), xi~U(0,2pi), rho~U(.5,1)."""
    gid = np.arange(N, dtype=np.int64)
    xi = (jitter_unit(seed, gid, 0) + 1.0) * np.pi
    rho = 0.75 + 0.25 * jitter_unit(seed, gid, 1)
    x = np.empty(2 * N)
    x[0::2] = rho * np.cos(xi)
    x[1::2] = rho * np.sin(xi)
    return x
