"""oracle/continuation.py on the CPU: the invariants of the two LOCA stepper defaults it restates (arc-length scaling,
hit-continuation-bound).  LOCA itself is absent (parity unpinned); these are properties of the algorithm as documented
in the oracle's header, which the device driver is then compared with in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle
from oracle import continuation as oc


@pytest.fixture(scope="module")
def problem():
    coords, cells = oracle.meshgen.tetgrid(7)
    psi, A = oracle.meshgen.plain_gl_fields(coords)
    return oracle.OracleProblem(coords, cells, ("explicit", A), V=-1.0, thickness=1.0), psi


def test_plain_run_is_unchanged_by_the_new_options(problem):
    P, psi = problem
    _, a = oc.arclength(P, 1.0, 0.0, psi, 0.05, 1e-7, 0.1, 2.0, 3)
    _, b = oc.arclength(P, 1.0, 0.0, psi, 0.05, 1e-7, 0.1, 2.0, 3, scaling=False, hit_bound=False)
    assert [r["param"] for r in a] == [r["param"] for r in b]
    assert all(r["scale"] == 1.0 for r in a)
    # the unit tangent: <xdot,xdot>/len + pdot^2 = 1 => |pdot| <= 1
    assert all(abs(r["dparam_ds"]) <= 1.0 for r in a)


def test_scaling_brings_the_parameter_share_to_the_goal(problem):
    P, psi = problem
    _, recs = oc.arclength(P, 1.0, 0.0, psi, 0.05, 1e-7, 0.05, 2.0, 4, scaling=True)
    r1 = recs[1]
    assert r1["scale"] < 1.0
    # at psi = 1, mu = 0 the state hardly reacts to mu: the unscaled share would be ~1 > 0.8; rescaled it is 0.5 at
    # the first tangent and stays below 0.8 afterwards
    for r in recs[1:]:
        assert abs(r["scale"] * r["dparam_ds"]) <= 0.8 + 1e-12
    # step sizes of the options are parameter increments: the first step moves mu by ds0
    assert r1["param"] == pytest.approx(0.05, rel=0.01)
    # every accepted step satisfies its constraint: the parameter moved in the tangent's direction
    assert all(recs[k + 1]["param"] > recs[k]["param"] for k in range(len(recs) - 1))


@pytest.mark.parametrize("ds0,lo,hi", [(0.05, -np.inf, 0.12), (-0.05, -0.08, np.inf)])
def test_hit_bound_ends_on_the_bound(problem, ds0, lo, hi):
    P, psi = problem
    bound = hi if ds0 > 0 else lo
    x, recs = oc.arclength(P, 1.0, 0.0, psi, ds0, 1e-7, 0.05, 2.0, 20, p_min=lo, p_max=hi, hit_bound=True)
    assert len(recs) < 20
    assert recs[-1]["param"] == bound                      # exactly
    assert abs(recs[-2]["param"] - bound) < 0.01           # the shortened arc-length step landed close to it
    assert all(lo <= r["param"] <= hi for r in recs)       # never outside
    # the final natural step solved F(x, bound) = 0
    P.keo_fill(bound, 0.0)
    assert np.linalg.norm(P.compute_f(1.0, x)) < 1e-8
    # without the option the run overshoots and stops outside
    _, plain = oc.arclength(P, 1.0, 0.0, psi, ds0, 1e-7, 0.05, 2.0, 20, p_min=lo, p_max=hi)
    assert not (lo <= plain[-1]["param"] <= hi)
