"""CPU restatement of the pseudo-arclength continuation driver.  TEST INFRASTRUCTURE ONLY.

The reference's nosh-cont hands the model evaluator to LOCA with "Continuation Method" = "Arc Length",
"Predictor" = "Tangent" and an adaptive step size (examples/conf.xml:35-75;
executables/nosh-cont/nosh-cont.cpp:224-344).  LOCA is third-party code outside the reference tree and no
reference test pins a continuation run => PARITY UNPINNED; restated here is the bordering algorithm it
implements [from the LOCA documentation, unverified against its source]:

  constraint   g(x, p) = <xdot, x - x0>/len + pdot (p - p0) - ds        (scaled dot product: Euclidean /
                                                                        vector length, parameter scale 1)
  corrector    J a = -F,  J b = -dF/dp,  dp = -(g + <xdot,a>/len) / (pdot + <xdot,b>/len),
               x += a + dp b,  p += dp,   until sqrt(||F||^2 + g^2) < nl_tol
  tangent      J t = -dF/dp,  (xdot, pdot) = +-(t, 1) / sqrt(<t,t>/len + 1), sign keeping the direction
  step size    ds *= 1 + aggressiveness ((nl_maxit - its)/(nl_maxit - 1))^2 after an accepted step, halved
               after a failed one

The device driver nosh_continuation_arclength (nosh_b200/csrc/krylov.cu) is compared with this file.
"""
import numpy as np


def arclength(P, g, p0, psi0, ds0, ds_min, ds_max, aggressiveness, max_steps, theta=0.0, nl_tol=1e-8,
              nl_maxit=20, lin_tol=1e-10, lin_maxit=1000, p_min=-np.inf, p_max=np.inf):
    """Continuation in mu on the OracleProblem P.  Returns (x, records)."""
    N = P.N
    length = 2.0 * N
    zeros = np.zeros(N)
    recs = []

    def record(k, x, mu, nsteps, lin, pred, fn, ds, pdot):
        recs.append(dict(step=k, param=mu, newton_steps=nsteps, linear_iterations=lin,
                         predictor_linear_iterations=pred, fnorm=fn, step_size=ds, dparam_ds=pdot,
                         gibbs_energy=P.gibbs_energy(x), norm=np.sqrt(P.inner_product(x, x))))

    def dfdp(x, mu):
        P.dkeo_fill(mu, theta, "mu")
        return P.compute_dfdp(x, False, zeros)

    def tangent(x, mu, xdot_old, pdot_old, sign0):
        P.keo_fill(mu, theta)
        P.jac_rebuild(g, x)
        t, its, _ = P.krylov(-dfdp(x, mu), lin_tol, lin_maxit)
        pd = 1.0 / np.sqrt(t @ t / length + 1.0)
        if xdot_old is None:
            if sign0 < 0:
                pd = -pd
        elif (t @ xdot_old) / length * pd + pd * pdot_old < 0.0:
            pd = -pd
        return pd * t, pd, its

    mu = p0
    P.keo_fill(mu, theta)
    x, steps, lin, fn = P.newton(g, np.array(psi0, np.float64), nl_tol, nl_maxit, lin_tol, lin_maxit)
    record(0, x, mu, steps, int(lin.sum()), 0, float(fn[-1]), 0.0, 0.0)
    if not fn[-1] < nl_tol or max_steps <= 0:
        return x, recs
    xdot, pdot, pred_its = tangent(x, mu, None, 0.0, ds0)
    ds = abs(ds0)
    x0, mu0 = x.copy(), mu
    k = 1
    while k <= max_steps:
        x = x0 + ds * xdot
        mu = mu0 + ds * pdot
        its = lin_total = 0
        ok = False
        while True:
            P.keo_fill(mu, theta)
            P.jac_rebuild(g, x)
            F = P.compute_f(g, x)
            gc = xdot @ (x - x0) / length + pdot * (mu - mu0) - ds
            nrm = np.sqrt(F @ F + gc * gc)
            if nrm < nl_tol:
                ok = True
                break
            if its >= nl_maxit or not np.isfinite(nrm):
                break
            Fp = dfdp(x, mu)
            a, ia, _ = P.krylov(-F, lin_tol, lin_maxit)
            b, ib, _ = P.krylov(-Fp, lin_tol, lin_maxit)
            lin_total += ia + ib
            dp = -(gc + xdot @ a / length) / (pdot + xdot @ b / length)
            x = x + a + dp * b
            mu = mu + dp
            its += 1
        if not ok:
            ds *= 0.5
            if ds < ds_min:
                x, mu = x0, mu0
                break
            continue
        ds_used = ds
        x0, mu0 = x.copy(), mu
        pred_prev = pred_its
        xdot, pdot, pred_its = tangent(x, mu, xdot, pdot, 0.0)
        record(k, x, mu, its, lin_total, pred_prev, float(nrm), ds_used, pdot)
        fac = (nl_maxit - its) / (nl_maxit - 1.0)
        ds = min(ds * (1.0 + aggressiveness * fac * fac), ds_max)
        if mu > p_max or mu < p_min:
            break
        k += 1
    return x, recs
