"""CPU: the drop-in boundary.  The C-ABI library loads, exports every symbol the header
declares (no compute calls: there is no GPU here), the product never touches the oracle,
and ctx creation fails loudly without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from nosh_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 45
    L = ctypes.CDLL(_lib.SO_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and the ctypes table binds exactly the declared set
    bound = set(_lib.lib()._nosh_signatures)
    assert bound == set(names), bound ^ set(names)


def test_only_nosh_symbols_are_exported():
    from nosh_b200 import _lib
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH], text=True)
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("nosh_") for s in exported), exported


def test_header_cites_reference_for_each_group():
    txt = open(os.path.join(ROOT, "include", "nosh_b200.h")).read()
    for cite in ["src/jacobian_operator.cpp:38-199", "src/parameter_matrix_keo.cpp:74-184",
                 "src/model_evaluator_nls.cpp:527-695", "src/keo_regularized.cpp:181-264",
                 "src/mesh_tetra.cpp", "src/vector_field_explicit_values.cpp:13-90"]:
        assert cite in txt, cite
    assert "torch" not in txt.lower().replace("no c++/torch types", "")


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "nosh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|libnosh_oracle|#include\s+\"[^\"]*oracle", src, re.M):
                    bad.append(f)
    assert not bad, bad


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_without_gpu():
    from nosh_b200 import Context, NoshError
    with pytest.raises(NoshError):
        Context()


def test_oracle_library_builds_and_loads():
    import oracle
    oracle.build_library()
    assert oracle.num_threads() >= 1


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/nosh_b200.h must compile as C99 (cgo / JNI / ctypes consumers)"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "nosh_b200.h"\n'
                   "int main(void) { nosh_arclength_options o; nosh_amg_info_t i; nosh_mesh_info_t m;\n"
                   "  (void)o; (void)i; (void)m; return (int)sizeof(nosh_krylov_result) == 0; }\n")
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
