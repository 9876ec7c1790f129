"""Exodus II files in the netCDF classic container (nosh_b200/csrc/exodus.inc; host only, no GPU): the format of the
reference's own test meshes (test/data/*.e.md5), which its build converts with MOAB.  Neither MOAB nor the netCDF
library is available here, so the reader parses the container itself; the files of these tests are written with an
INDEPENDENT implementation of the container, scipy.io.netcdf_file (CDF-1 and 64-bit-offset CDF-2), laid out the way
meshio / SEACAS write Exodus; a CDF-5 header is assembled by hand from the format description."""
import struct

import numpy as np
import pytest
from scipy.io import netcdf_file

import nosh_b200
from oracle import meshgen


def write_exodus(path, coords, blocks, nodal=None, version=2, large_model=True, steps=2, packed_vars=False):
    """blocks: list of (elem_type, (n, k) 0-based connectivity); nodal: dict name -> (N,) array (last time step)."""
    nodal = nodal or {}
    N = coords.shape[0]
    nd = 3 if np.abs(coords[:, 2]).max() > 0 else 2
    f = netcdf_file(str(path), "w", version=version)
    f.title = b"written by tests/test_exodus.py"
    f.api_version = np.float32(5.1)
    f.floating_point_word_size = np.int32(8)
    f.createDimension("time_step", None)                # scipy wants the record dimension first
    f.createDimension("len_string", 33)
    f.createDimension("len_line", 81)
    f.createDimension("four", 4)
    f.createDimension("num_dim", nd)
    f.createDimension("num_nodes", N)
    f.createDimension("num_elem", sum(c.shape[0] for _, c in blocks))
    f.createDimension("num_el_blk", len(blocks))
    tw = f.createVariable("time_whole", "d", ("time_step",))
    st = f.createVariable("eb_status", "i", ("num_el_blk",))
    st[:] = 1
    if large_model:
        for d, nm in enumerate(("coordx", "coordy", "coordz")[:nd]):
            v = f.createVariable(nm, "d", ("num_nodes",))
            v[:] = coords[:, d]
    else:
        v = f.createVariable("coord", "d", ("num_dim", "num_nodes"))
        v[:] = coords[:, :nd].T
    for b, (ety, conn) in enumerate(blocks, 1):
        f.createDimension("num_el_in_blk%d" % b, conn.shape[0])
        f.createDimension("num_nod_per_el%d" % b, conn.shape[1])
        v = f.createVariable("connect%d" % b, "i", ("num_el_in_blk%d" % b, "num_nod_per_el%d" % b))
        v.elem_type = ety.encode()
        v[:] = conn + 1                                  # Exodus numbers nodes from 1
    if nodal:
        names = list(nodal)
        f.createDimension("num_nod_var", len(names))
        nm = f.createVariable("name_nod_var", "c", ("num_nod_var", "len_string"))
        for k, s in enumerate(names):
            nm[k] = np.frombuffer(s.encode().ljust(33, b"\0"), "S1")
        if packed_vars:
            vv = f.createVariable("vals_nod_var", "d", ("time_step", "num_nod_var", "num_nodes"))
            for t in range(steps):
                for k, s in enumerate(names):
                    vv[t, k, :] = nodal[s] if t == steps - 1 else 0.0 * nodal[s] + 7.0
        else:
            vs = [f.createVariable("vals_nod_var%d" % (k + 1), "d", ("time_step", "num_nodes")) for k in range(len(names))]
            for t in range(steps):
                for k, s in enumerate(names):
                    vs[k][t, :] = nodal[s] if t == steps - 1 else 0.0 * nodal[s] + 7.0   # earlier steps: junk
    for t in range(steps):
        tw[t] = float(t)
    f.close()


def tagged_mesh(n=4):
    coords, cells = meshgen.tetgrid(n)
    psi, A = meshgen.plain_gl_fields(coords)
    psi = meshgen.random_state(coords.shape[0], 3).reshape(-1, 2)
    nodal = {"psi_R": psi[:, 0], "psi_Z": psi[:, 1], "A_X": A[:, 0], "A_Y": A[:, 1], "A_Z": A[:, 2],
             "V": -np.ones(coords.shape[0]), "thickness": 1.0 + 0.1 * coords[:, 0]}
    return coords, cells, psi, A, nodal


@pytest.mark.parametrize("version,large,packed", [(1, True, False), (2, True, False), (2, False, True)])
def test_tetrahedral_mesh_with_the_reference_tags(tmp_path, version, large, packed):
    coords, cells, psi, A, nodal = tagged_mesh()
    path = tmp_path / "state.e"
    write_exodus(path, coords, [("TETRA", cells)], nodal, version=version, large_model=large, packed_vars=packed)
    c2, t2, f2 = nosh_b200.read_mesh(path)
    assert np.array_equal(c2, coords) and np.array_equal(t2, cells)
    assert sorted(f2) == ["A", "V", "psi", "thickness"]
    assert f2["psi"].shape == psi.shape and np.array_equal(f2["psi"], psi)          # (re, im) joined, last time step
    assert f2["A"].shape == A.shape and np.array_equal(f2["A"], A)
    assert np.array_equal(f2["V"], nodal["V"]) and np.array_equal(f2["thickness"], nodal["thickness"])


def test_same_mesh_through_vtk_and_exodus(tmp_path):
    """the two readers agree, so everything downstream (nosh::read, get_complex_vector, ...) is format independent"""
    coords, cells, psi, A, nodal = tagged_mesh(5)
    write_exodus(tmp_path / "m.exo", coords, [("TETRA4", cells[:100]), ("TETRA4", cells[100:])], nodal)
    nosh_b200.write_mesh(tmp_path / "m.vtk", coords, cells, {"psi": psi, "A": A, "V": nodal["V"]}, binary=True)
    ce, te, fe = nosh_b200.read_mesh(tmp_path / "m.exo")
    cv, tv, fv = nosh_b200.read_mesh(tmp_path / "m.vtk")
    assert np.array_equal(ce, cv) and np.array_equal(te, tv)                         # blocks concatenated in order
    for k in ("psi", "A", "V"):
        assert np.array_equal(fe[k], fv[k])


def test_triangles_mixed_blocks_and_no_variables(tmp_path):
    coords, cells = meshgen.rectanglesmall()
    quads = np.array([[0, 1, 2, 3]])
    path = tmp_path / "rect.g"
    write_exodus(path, coords, [("QUAD4", quads), ("TRI3", cells)], None, version=1)
    c2, t2, f2 = nosh_b200.read_mesh(path)
    assert t2.shape == cells.shape and np.array_equal(t2, cells) and f2 == {}          # the quad block is ignored
    assert np.array_equal(c2[:, :2], coords[:, :2]) and np.all(c2[:, 2] == 0.0)
    # tetrahedra win over triangles when both are present (the VTK reader's rule)
    coords3, tets = meshgen.tetgrid(2)
    write_exodus(tmp_path / "both.e", coords3, [("TRI3", tets[:, :3]), ("TETRA", tets)], None)
    _, t3, _ = nosh_b200.read_mesh(tmp_path / "both.e")
    assert t3.shape[1] == 4 and np.array_equal(t3, tets)


def nc_name(s, cnt):
    b = s.encode()
    return cnt(len(b)) + b + b"\0" * ((4 - len(b) % 4) % 4)


def test_cdf5_header(tmp_path):
    """64-bit data variant (magic CDF\\x05): counts, dimension ids and vsize are 8 bytes.  Assembled by hand from the
    format description: 3 nodes, one triangle, no record variables."""
    q = lambda v: struct.pack(">q", v)
    i4 = lambda v: struct.pack(">i", v)
    dims = [("num_dim", 2), ("num_nodes", 3), ("num_el_blk", 1), ("num_el_in_blk1", 1), ("num_nod_per_el1", 3)]
    hdr = b"CDF\x05" + q(0)
    hdr += i4(0x0A) + q(len(dims)) + b"".join(nc_name(n, q) + q(l) for n, l in dims)
    hdr += i4(0) + q(0)                                                  # no global attributes
    et = b"TRI3"
    var_specs = [("coordx", [1], b"", 6, 24), ("coordy", [1], b"", 6, 24),
                 ("connect1", [3, 4], i4(0x0C) + q(1) + nc_name("elem_type", q) + i4(2) + q(len(et)) + et, 4, 12)]
    # two passes: the header length fixes the data offsets
    def build(begins):
        out = i4(0x0B) + q(len(var_specs))
        for (name, dimids, atts, ty, vsize), beg in zip(var_specs, begins):
            out += nc_name(name, q) + q(len(dimids)) + b"".join(q(d) for d in dimids)
            out += (atts if atts else i4(0) + q(0)) + i4(ty) + q(vsize) + q(beg)
        return out
    hlen = len(hdr) + len(build([0, 0, 0]))
    begins = [hlen, hlen + 24, hlen + 48]
    data = struct.pack(">3d", 0.0, 2.0, 0.0) + struct.pack(">3d", 0.0, 0.0, 1.5) + struct.pack(">3i", 1, 2, 3)
    path = tmp_path / "tri5.e"
    path.write_bytes(hdr + build(begins) + data)
    c, t, f = nosh_b200.read_mesh(path)
    assert np.array_equal(c, [[0, 0, 0], [2, 0, 0], [0, 1.5, 0]]) and np.array_equal(t, [[0, 1, 2]]) and f == {}


def test_errors(tmp_path):
    coords, cells, psi, A, nodal = tagged_mesh(3)
    # netCDF-4 (= HDF5) container: recognised by its signature, unsupported, with the way out in the message
    h5 = tmp_path / "pacman.e"
    h5.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(nosh_b200.NoshError, match="nccopy"):
        nosh_b200.read_mesh(h5)
    junk = tmp_path / "junk.e"
    junk.write_bytes(b"this is not a mesh")
    with pytest.raises(ValueError, match="netCDF"):
        nosh_b200.read_mesh(junk)
    with pytest.raises(ValueError):
        nosh_b200.read_mesh(tmp_path / "missing.e")
    # a connectivity entry outside 1..num_nodes
    bad = cells.copy()
    bad[0, 0] = coords.shape[0]
    write_exodus(tmp_path / "bad.e", coords, [("TETRA", bad)], None)
    with pytest.raises(ValueError, match="connectivity"):
        nosh_b200.read_mesh(tmp_path / "bad.e")
    # a truncated file: the header promises more data than there is
    good = tmp_path / "good.e"
    write_exodus(good, coords, [("TETRA", cells)], nodal)
    raw = good.read_bytes()
    (tmp_path / "cut.e").write_bytes(raw[:len(raw) // 2])
    with pytest.raises(ValueError):
        nosh_b200.read_mesh(tmp_path / "cut.e")
    # only cells the path cannot use
    write_exodus(tmp_path / "hex.e", coords, [("HEX8", np.arange(8)[None, :])], None)
    with pytest.raises(nosh_b200.NoshError, match="no triangles"):
        nosh_b200.read_mesh(tmp_path / "hex.e")
    # a netCDF file that is not Exodus
    f = netcdf_file(str(tmp_path / "other.e"), "w")
    f.createDimension("x", 3)
    f.close()
    with pytest.raises(ValueError, match="Exodus"):
        nosh_b200.read_mesh(tmp_path / "other.e")
