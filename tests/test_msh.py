"""gmsh MSH files, ASCII 2.2 and 4.1 (nosh_b200/csrc/msh.inc; host only): the reference's meshes are gmsh output
(examples/meshes/*.geo, test/data/*.geo) that is converted before nosh::read sees it; this reader takes the .msh
directly.  gmsh is not installed here, so the files are written by the two small writers below, straight from the
format description in the gmsh reference manual, in the shape gmsh itself emits: point / line / surface elements
before the volume ones, non-contiguous node tags, nodes grouped by the entity they were meshed on."""
import numpy as np
import pytest

import nosh_b200
from oracle import meshgen


def boundary_faces(cells):
    """faces that belong to one tetrahedron only (what gmsh lists as the surface mesh)"""
    f = np.concatenate([cells[:, [1, 2, 3]], cells[:, [0, 2, 3]], cells[:, [0, 1, 3]], cells[:, [0, 1, 2]]])
    key = np.sort(f, axis=1)
    _, idx, cnt = np.unique(key, axis=0, return_index=True, return_counts=True)
    return f[idx[cnt == 1]]


def write_msh22(path, coords, cells, tags, node_data=None, extra_node=True):
    tets = cells.shape[1] == 4
    surf = boundary_faces(cells) if tets else np.zeros((0, 3), int)
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        f.write('$PhysicalNames\n1\n%d 7 "domain"\n$EndPhysicalNames\n' % (3 if tets else 2))
        f.write("$Nodes\n%d\n" % (len(coords) + (1 if extra_node else 0)))
        if extra_node:                                     # a node no cell uses (e.g. the centre of a circle arc)
            f.write("%d 99.0 99.0 99.0\n" % (tags.max() + 5))
        for t, x in zip(tags, coords):
            f.write("%d %.17g %.17g %.17g\n" % (t, *x))
        f.write("$EndNodes\n$Elements\n%d\n" % (1 + len(surf) + len(cells)))
        e = 1
        f.write("%d 15 2 0 1 %d\n" % (e, tags[0]))         # a point element
        for s in surf:
            e += 1
            f.write("%d 2 2 5 5 %d %d %d\n" % (e, *tags[s]))
        for c in cells:
            e += 1
            f.write(("%d %d 2 7 1 " % (e, 4 if tets else 2)) + " ".join(str(tags[v]) for v in c) + "\n")
        f.write("$EndElements\n")
        for name, v in (node_data or {}).items():
            v = np.asarray(v).reshape(len(coords), -1)
            f.write('$NodeData\n1\n"%s"\n1\n0.0\n3\n0\n%d\n%d\n' % (name, v.shape[1], len(coords)))
            for t, row in zip(tags, v):
                f.write("%d " % t + " ".join("%.17g" % x for x in row) + "\n")
            f.write("$EndNodeData\n")


def write_msh41(path, coords, cells, tags, node_data=None):
    tets = cells.shape[1] == 4
    surf = boundary_faces(cells) if tets else np.zeros((0, 3), int)
    on_surf = np.zeros(len(coords), bool)
    on_surf[np.unique(surf)] = True
    groups = [(2, 1, np.flatnonzero(on_surf)), (3 if tets else 2, 1 if tets else 2, np.flatnonzero(~on_surf))]
    groups = [g for g in groups if len(g[2])]
    with open(path, "w") as f:
        f.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
        f.write("$Entities\n0 0 1 1\n1 0 0 0 1 1 1 0 0\n1 0 0 0 1 1 1 0 1 1\n$EndEntities\n")
        f.write("$Nodes\n%d %d %d %d\n" % (len(groups), len(coords), tags.min(), tags.max()))
        for edim, etag, idx in groups:
            param = 1 if edim == 2 else 0                  # parametric coordinates on the surface: u v
            f.write("%d %d %d %d\n" % (edim, etag, param, len(idx)))
            for i in idx:
                f.write("%d\n" % tags[i])
            for i in idx:
                f.write("%.17g %.17g %.17g" % tuple(coords[i]) + (" 0.25 0.75\n" if param else "\n"))
        f.write("$EndNodes\n")
        nblk = (1 if len(surf) else 0) + 1
        f.write("$Elements\n%d %d 1 %d\n" % (nblk, len(surf) + len(cells), len(surf) + len(cells)))
        e = 0
        if len(surf):
            f.write("2 1 2 %d\n" % len(surf))
            for s in surf:
                e += 1
                f.write("%d %d %d %d\n" % (e, *tags[s]))
        f.write("%d 1 %d %d\n" % (3 if tets else 2, 4 if tets else 2, len(cells)))
        for c in cells:
            e += 1
            f.write("%d " % e + " ".join(str(tags[v]) for v in c) + "\n")
        f.write("$EndElements\n")
        for name, v in (node_data or {}).items():
            v = np.asarray(v).reshape(len(coords), -1)
            f.write('$NodeData\n1\n"%s"\n1\n0.0\n4\n0\n%d\n%d\n0\n' % (name, v.shape[1], len(coords)))
            for t, row in zip(tags, v):
                f.write("%d " % t + " ".join("%.17g" % x for x in row) + "\n")
            f.write("$EndNodeData\n")


def tagged(n=4):
    coords, cells = meshgen.tetgrid(n)
    tags = 3 * np.arange(len(coords)) + 11                 # not contiguous, not starting at 1
    psi = meshgen.random_state(len(coords), 2).reshape(-1, 2)
    _, A = meshgen.plain_gl_fields(coords)
    data = {"psi_R": psi[:, 0], "psi_Z": psi[:, 1], "A": A, "V": -np.ones(len(coords))}
    return coords, cells, tags, psi, A, data


def test_msh22_tetrahedra_with_node_data(tmp_path):
    coords, cells, tags, psi, A, data = tagged()
    write_msh22(tmp_path / "m.msh", coords, cells, tags, data)
    c, t, f = nosh_b200.read_mesh(tmp_path / "m.msh")
    assert np.array_equal(c, coords) and np.array_equal(t, cells)         # surface triangles and the point are skipped,
    assert sorted(f) == ["A", "V", "psi"]                                  # the unused node is dropped
    assert np.array_equal(f["psi"], psi) and np.array_equal(f["A"], A) and np.array_equal(f["V"], data["V"])


def test_msh41_blocks_and_parametric_nodes(tmp_path):
    coords, cells, tags, psi, A, data = tagged(5)
    write_msh41(tmp_path / "m.msh", coords, cells, tags, data)
    c, t, f = nosh_b200.read_mesh(tmp_path / "m.msh")
    # 4.1 lists the nodes entity by entity (surface first): same mesh up to that renumbering
    assert c.shape == coords.shape and t.shape == cells.shape
    key = lambda x: np.lexsort(x.T[::-1])
    po, pn = key(coords), key(c)
    assert np.array_equal(coords[po], c[pn])
    new_of_old = np.empty(len(coords), int)
    new_of_old[po] = pn
    assert np.array_equal(new_of_old[cells], t)
    assert np.array_equal(f["psi"][new_of_old], psi) and np.array_equal(f["A"][new_of_old], A)


def test_both_versions_give_the_vtk_readers_mesh(tmp_path):
    coords, cells, tags, psi, A, data = tagged(3)
    write_msh22(tmp_path / "a.msh", coords, cells, tags, data, extra_node=False)
    nosh_b200.write_mesh(tmp_path / "a.vtk", coords, cells, {"psi": psi, "A": A, "V": data["V"]}, binary=True)
    cm, tm, fm = nosh_b200.read_mesh(tmp_path / "a.msh")
    cv, tv, fv = nosh_b200.read_mesh(tmp_path / "a.vtk")
    assert np.array_equal(cm, cv) and np.array_equal(tm, tv)
    assert all(np.array_equal(fm[k], fv[k]) for k in ("psi", "A", "V"))


def test_triangle_mesh(tmp_path):
    coords, cells = meshgen.rectanglesmall()
    tags = np.arange(len(coords)) + 1
    write_msh22(tmp_path / "r22.msh", coords, cells, tags, None)
    write_msh41(tmp_path / "r41.msh", coords, cells, tags, None)
    for name in ("r22.msh", "r41.msh"):
        c, t, f = nosh_b200.read_mesh(tmp_path / name)
        assert t.shape == cells.shape and f == {}
        # same triangles as point sets (4.1 may renumber)
        tri = lambda cc, tt: sorted(sorted(map(tuple, cc[row])) for row in tt)
        assert tri(c, t) == tri(coords, cells)


def test_errors(tmp_path):
    coords, cells, tags, psi, A, data = tagged(3)
    (tmp_path / "bin.msh").write_text("$MeshFormat\n4.1 1 8\n")
    with pytest.raises(nosh_b200.NoshError, match="ASCII"):
        nosh_b200.read_mesh(tmp_path / "bin.msh")
    (tmp_path / "v40.msh").write_text("$MeshFormat\n4.0 0 8\n$EndMeshFormat\n")
    with pytest.raises(nosh_b200.NoshError, match="4.1"):
        nosh_b200.read_mesh(tmp_path / "v40.msh")
    (tmp_path / "junk.msh").write_text("hello\n")
    with pytest.raises(ValueError, match="gmsh"):
        nosh_b200.read_mesh(tmp_path / "junk.msh")
    # an element that names a node tag the file does not define
    write_msh22(tmp_path / "ok.msh", coords, cells, tags, None)
    txt = (tmp_path / "ok.msh").read_text().replace("$EndElements", "").rstrip("\n")
    last = txt.rsplit("\n", 1)[1].split()
    last[-1] = "999999"
    (tmp_path / "badtag.msh").write_text(txt.rsplit("\n", 1)[0] + "\n" + " ".join(last) + "\n$EndElements\n")
    with pytest.raises(ValueError, match="connectivity"):
        nosh_b200.read_mesh(tmp_path / "badtag.msh")
    # truncated in the middle of the node list
    full = (tmp_path / "ok.msh").read_text()
    (tmp_path / "cut.msh").write_text(full[:full.index("$EndNodes") - 40])
    with pytest.raises(ValueError):
        nosh_b200.read_mesh(tmp_path / "cut.msh")
    # only lines and points
    (tmp_path / "lines.msh").write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n2\n1 0 0 0\n2 1 0 0\n$EndNodes\n"
                                        "$Elements\n1\n1 1 2 0 1 1 2\n$EndElements\n")
    with pytest.raises(nosh_b200.NoshError, match="no triangles"):
        nosh_b200.read_mesh(tmp_path / "lines.msh")
