"""Mesh file I/O and vertex ordering ("next" row f3): host-only entry points of the C ABI, no GPU needed.
Reference behaviour mirrored: nosh::read + vertex tags psi / A / V (src/mesh_reader.cpp:19-162,
src/mesh.cpp:249-446) and mesh::write for the outNNNN dumps, on legacy VTK files (MOAB/HDF5 are not
available offline)."""
import numpy as np
import pytest

import nosh_b200
from oracle import meshgen


def tagged_mesh(n=5):
    coords, cells = meshgen.tetgrid(n)
    psi, A = meshgen.plain_gl_fields(coords)       # examples/state-equippers/plain-gl:22-39
    psi = meshgen.random_state(coords.shape[0], 3)
    V = -np.ones(coords.shape[0])
    return coords, cells, {"psi": psi.reshape(-1, 2), "A": A, "V": V}


@pytest.mark.parametrize("binary", [False, True])
def test_vtk_round_trip(tmp_path, binary):
    coords, cells, fields = tagged_mesh()
    path = tmp_path / ("out0001_%d.vtk" % binary)
    nosh_b200.write_mesh(path, coords, cells, fields, binary=binary)
    c2, t2, f2 = nosh_b200.read_mesh(path)
    assert np.array_equal(c2, coords)              # 17 significant digits / raw doubles: bit-exact
    assert np.array_equal(t2, cells)
    assert sorted(f2) == ["A", "V", "psi"]
    assert f2["psi"].shape == (coords.shape[0], 2) and np.array_equal(f2["psi"], fields["psi"])
    assert f2["A"].shape == (coords.shape[0], 3) and np.array_equal(f2["A"], fields["A"])
    assert f2["V"].shape == (coords.shape[0],) and np.array_equal(f2["V"], fields["V"])
    # the known answers of test/io.cpp are norms of exactly these tags
    assert np.abs(f2["psi"]).sum() > 0 and np.abs(f2["A"]).max(axis=0).shape == (3,)


def test_triangle_mesh_and_foreign_cells(tmp_path):
    """a hand-written file in the style meshio produces: float points, mixed cell types, cell data"""
    path = tmp_path / "rect.vtk"
    path.write_text("""# vtk DataFile Version 4.2
written by hand
ASCII
DATASET UNSTRUCTURED_GRID
POINTS 4 float
5 0.5 0  -5 -0.5 0  5 -0.5 0  -5 0.5 0
CELLS 4 13
3 0 1 2
3 0 3 1
2 0 1
1 2
CELL_TYPES 4
5
5
3
1
POINT_DATA 4
SCALARS psi double 2
LOOKUP_TABLE default
1 0 1 0 1 0 1 0
VECTORS A double
0.25 2.5 0 -0.25 -2.5 0 0.25 2.5 0 -0.25 -2.5 0
CELL_DATA 4
SCALARS region int 1
LOOKUP_TABLE default
1 1 2 3
""")
    coords, cells, fields = nosh_b200.read_mesh(path)
    ref_c, ref_t = meshgen.rectanglesmall()
    assert np.array_equal(coords, ref_c) and np.array_equal(cells, ref_t)
    # test/io.cpp:51-59 (rectanglesmall): ||psi||_1 = 4, ||psi||_inf = 1, ||A||_inf = (0.25, 2.5, 0)
    psi = fields["psi"]
    assert np.hypot(psi[:, 0], psi[:, 1]).sum() == 4.0 and np.hypot(psi[:, 0], psi[:, 1]).max() == 1.0
    assert np.array_equal(np.abs(fields["A"]).max(axis=0), [0.25, 2.5, 0.0])
    assert "region" not in fields


def test_errors(tmp_path):
    with pytest.raises(nosh_b200.NoshError, match="meshio-convert"):
        nosh_b200.read_mesh(tmp_path / "pacman.h5m")
    with pytest.raises(ValueError):
        nosh_b200.read_mesh(tmp_path / "missing.vtk")
    bad = tmp_path / "bad.vtk"
    bad.write_text("not a vtk file\n")
    with pytest.raises(ValueError):
        nosh_b200.read_mesh(bad)
    trunc = tmp_path / "trunc.vtk"
    trunc.write_text("# vtk DataFile Version 3.0\nx\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 3 double\n0 0 0 1 0\n")
    with pytest.raises(ValueError):
        nosh_b200.read_mesh(trunc)


def test_morton_order_makes_contiguous_ranges_compact():
    rng = np.random.default_rng(0)
    coords, cells = meshgen.tetgrid(12)
    shuffle = rng.permutation(coords.shape[0])            # an "unstructured" numbering
    c1, t1, _ = nosh_b200.renumber(coords, cells, shuffle)
    perm = nosh_b200.morton_order(c1)
    assert np.array_equal(np.sort(perm), np.arange(perm.size))
    c2, t2, f2 = nosh_b200.renumber(c1, t1, perm, {"id": np.arange(perm.size)})
    assert np.array_equal(f2["id"], perm)
    # same mesh: cell volumes are a permutation-invariant fingerprint
    def vol(c, t):
        e = c[t[:, 1:]] - c[t[:, :1]]
        return np.sort(np.abs(np.linalg.det(e)))
    assert np.allclose(vol(c2, t2), vol(coords, cells), rtol=0, atol=1e-12)

    def cut_vertices(t, nparts=4):
        n = t.max() + 1
        part = np.minimum(np.arange(n) * nparts // n, nparts - 1)
        pc = part[t]
        cut = (pc != pc[:, :1]).any(axis=1)
        return np.unique(t[cut]).size
    # contiguous ranges of the Morton numbering cut far fewer cells than ranges of the shuffled numbering
    assert cut_vertices(t2) < 0.5 * cut_vertices(t1)


def test_vtk5_offsets_connectivity_layout(tmp_path):
    """files written by VTK >= 9 / recent meshio use 'CELLS n+1 nconn' + OFFSETS / CONNECTIVITY arrays"""
    path = tmp_path / "rect5.vtk"
    path.write_text("""# vtk DataFile Version 5.1
vtk output
ASCII
DATASET UNSTRUCTURED_GRID
POINTS 4 double
5 0.5 0 -5 -0.5 0 5 -0.5 0 -5 0.5 0
CELLS 3 6
OFFSETS vtktypeint64
0 3 6
CONNECTIVITY vtktypeint64
0 1 2 0 3 1
CELL_TYPES 2
5
5
POINT_DATA 4
FIELD FieldData 2
V 1 4 double
-1 -1 -1 -1
psi 2 4 double
1 0 1 0 1 0 1 0
""")
    coords, cells, fields = nosh_b200.read_mesh(path)
    ref_c, ref_t = meshgen.rectanglesmall()
    assert np.array_equal(coords, ref_c) and np.array_equal(cells, ref_t)
    assert np.array_equal(fields["V"], -np.ones(4)) and fields["psi"].shape == (4, 2)
    assert np.array_equal(fields["psi"][:, 0], np.ones(4))
