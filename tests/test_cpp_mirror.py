"""The C++ mirror of the reference's classes (nosh_b200/hostcpp/nosh.hpp): the reference's own
Catch tests re-typed against it (hostcpp/test_reference_style.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "nosh_b200", "hostcpp", "test_reference_style")


def _build():
    subprocess.check_call(["make", "-C", os.path.dirname(EXE)], stdout=subprocess.DEVNULL)


def test_mirror_compiles_and_fails_loudly_without_gpu():
    _build()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu test")
    except ImportError:
        pass
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stdout


def test_mirror_has_the_reference_signatures():
    src = open(os.path.join(ROOT, "nosh_b200", "hostcpp", "nosh.hpp")).read()
    for needle in ["class jacobian_operator : public Tpetra::Operator<double, int, int>",
                   "void rebuild(const std::map<std::string, double> &params, const Tpetra::Vector<double, int, int> &current_x)",
                   "class keo : public matrix_base", "class DkeoDP : public matrix_base",
                   "class keo_regularized : public Tpetra::Operator<double, int, int>",
                   "const std::string &deriv_parameter", "void evalModel(const InArgs &in, const OutArgs &out) const",
                   "create_W_op() const", "create_W_prec() const", "get_p_names(int l) const"]:
        assert needle in src, needle


@pytest.mark.gpu
def test_reference_style_tests_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM TESTS PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
