"""Robustness of the three mesh-file readers (legacy VTK, Exodus II / netCDF classic, gmsh MSH): a corrupted or
truncated file must end in an error code -- never in a crash, a hang or an unbounded allocation inside the caller's
process (the readers are C code behind the C ABI).  Runs in a child process so that a crash is seen as one."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent('''
    import os, sys, tempfile
    import numpy as np
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, "tests"))
    import nosh_b200
    import test_exodus as te
    import test_msh as tm
    d = tempfile.mkdtemp()
    coords, cells, psi, A, nodal = te.tagged_mesh(3)
    te.write_exodus(os.path.join(d, "a.e"), coords, [("TETRA", cells)], nodal, version=2)
    te.write_exodus(os.path.join(d, "b.e"), coords, [("TETRA", cells)], nodal, version=1, large_model=False, packed_vars=True)
    c2, t2, tags, psi2, A2, data = tm.tagged(3)
    tm.write_msh22(os.path.join(d, "a.msh"), c2, t2, tags, data)
    tm.write_msh41(os.path.join(d, "b.msh"), c2, t2, tags, data)
    nosh_b200.write_mesh(os.path.join(d, "a.vtk"), coords, cells, {"psi": psi, "A": A}, binary=True)
    nosh_b200.write_mesh(os.path.join(d, "b.vtk"), coords, cells, {"psi": psi, "A": A}, binary=False)
    names = ["a.e", "b.e", "a.msh", "b.msh", "a.vtk", "b.vtk"]
    rng = np.random.default_rng(7)
    parsed = rejected = 0
    for it in range(3000):
        name = names[it %% 6]
        raw = bytearray(open(os.path.join(d, name), "rb").read())
        mode = it // 6 %% 4
        if mode == 0:                                   # a few flipped bytes anywhere
            for _ in range(rng.integers(1, 8)):
                raw[rng.integers(0, len(raw))] = rng.integers(0, 256)
        elif mode == 1:                                 # truncated
            raw = raw[:rng.integers(0, len(raw))]
        elif mode == 2:                                 # header region shredded
            for _ in range(rng.integers(1, 20)):
                raw[rng.integers(0, min(len(raw), 400))] = rng.integers(0, 256)
        else:                                           # one huge 32-bit word / long digit run
            pos = rng.integers(0, max(1, len(raw) - 4))
            raw[pos:pos + 4] = bytes([255, 255, 255, int(rng.integers(0, 256))])
        q = os.path.join(d, "fuzz" + os.path.splitext(name)[1])
        open(q, "wb").write(bytes(raw))
        try:
            c, t, f = nosh_b200.read_mesh(q)
            assert t.min() >= 0 and t.max() < c.shape[0]   # whatever was accepted is at least index-safe
            parsed += 1
        except (ValueError, nosh_b200.NoshError, MemoryError):
            rejected += 1
    print("FUZZ OK parsed=%%d rejected=%%d" %% (parsed, rejected))
''') % (ROOT, ROOT)


def test_corrupted_files_never_crash_the_process():
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    assert "FUZZ OK" in r.stdout
    rejected = int(r.stdout.split("rejected=")[1].split()[0])
    assert rejected > 1000                                 # most corruptions are noticed, none kills the process
