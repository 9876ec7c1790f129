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

Optional (scaling=True, hit_bound=True: LOCA's defaults "Enable Arc Length Scaling" and "Hit Continuation Bound",
which nosh-cont inherits because examples/conf.xml sets neither) [from the LOCA documentation and the derivation
below, unverified against its source]:

  parameter scale    the scaled dot product is <x,y>/len + s^2 p q with a scale factor s (initially 1).  The share
                     of the arc length taken by the parameter is c = s |pdot|.  Whenever a new tangent has
                     c > c_max (0.8), s is reset such that c = c_goal (0.5):  with T = <t,t>/len,
                     c^2 = s^2/(T + s^2)  =>  s_new = (c_goal/|pdot|) sqrt((1 - c^2)/(1 - c_goal^2)),  >= 1e-3;
                     the tangent is renormalised and ds, ds_min, ds_max are multiplied by |pdot_old/pdot_new|, which
                     keeps the predicted parameter increment ds pdot.
  step-size units    the initial / minimum / maximum step sizes are parameter increments: divided by |pdot| of the
                     first tangent to become arc lengths
  bounds             a step whose predictor would leave [p_min, p_max] is shortened to land on the bound,
                     ds = (bound - p0)/pdot, and is the last arc-length step; the run then ends with one
                     natural-continuation step to the bound itself (constant predictor, Newton on F alone)

The device driver nosh_continuation_arclength (nosh_b200/csrc/krylov.cu) is compared with this file.
"""
import numpy as np


def arclength(P, g, p0, psi0, ds0, ds_min, ds_max, aggressiveness, max_steps, theta=0.0, nl_tol=1e-8,
              nl_maxit=20, lin_tol=1e-10, lin_maxit=1000, p_min=-np.inf, p_max=np.inf, scaling=False,
              hit_bound=False, c_goal=0.5, c_max=0.8, scale_min=1e-3, scale0=1.0):
    """Continuation in mu on the OracleProblem P.  Returns (x, records)."""
    N = P.N
    length = 2.0 * N
    zeros = np.zeros(N)
    recs = []
    sc = scale0 if scaling else 1.0     # parameter scale factor s

    def record(k, x, mu, nsteps, lin, pred, fn, ds, pdot):
        recs.append(dict(step=k, param=mu, newton_steps=nsteps, linear_iterations=lin,
                         predictor_linear_iterations=pred, fnorm=fn, step_size=ds, dparam_ds=pdot,
                         gibbs_energy=P.gibbs_energy(x), norm=np.sqrt(P.inner_product(x, x)), scale=sc))

    def dfdp(x, mu):
        P.dkeo_fill(mu, theta, "mu")
        return P.compute_dfdp(x, False, zeros)

    def tangent(x, mu, xdot_old, pdot_old, sign0):
        """Returns (xdot, pdot, iterations, ratio): ratio = |pdot before / after a change of the scale factor|."""
        nonlocal sc
        P.keo_fill(mu, theta)
        P.jac_rebuild(g, x)
        t, its, _ = P.krylov(-dfdp(x, mu), lin_tol, lin_maxit)
        T = t @ t / length
        pd = 1.0 / np.sqrt(T + sc * sc)
        ratio = 1.0
        if scaling:
            c = sc * pd
            if c > c_max:
                sc_new = max(scale_min, c_goal / pd * np.sqrt((1.0 - c * c) / (1.0 - c_goal * c_goal)))
                pd_new = 1.0 / np.sqrt(T + sc_new * sc_new)
                ratio = pd / pd_new
                # the direction test below compares with the OLD tangent, normalised in the old scale
                if xdot_old is not None and (t @ xdot_old) / length * pd + sc * sc * pd * pdot_old < 0.0:
                    pd_new = -pd_new
                elif xdot_old is None and sign0 < 0:
                    pd_new = -pd_new
                sc = sc_new
                return pd_new * t, pd_new, its, ratio
        if xdot_old is None:
            if sign0 < 0:
                pd = -pd
        elif (t @ xdot_old) / length * pd + sc * sc * pd * pdot_old < 0.0:
            pd = -pd
        return pd * t, pd, its, ratio

    mu = p0
    P.keo_fill(mu, theta)
    x, steps, lin, fn = P.newton(g, np.array(psi0, np.float64), nl_tol, nl_maxit, lin_tol, lin_maxit)
    record(0, x, mu, steps, int(lin.sum()), 0, float(fn[-1]), 0.0, 0.0)
    if not fn[-1] < nl_tol or max_steps <= 0:
        return x, recs
    xdot, pdot, pred_its, ratio = tangent(x, mu, None, 0.0, ds0)
    ds = abs(ds0)
    if scaling:
        # parameter increments -> arc lengths, with the (rescaled) first tangent
        u = 1.0 / abs(pdot)
        ds, ds_min, ds_max = ds * u, ds_min * u, ds_max * u
    x0, mu0 = x.copy(), mu
    k = 1
    reached = None
    while k <= max_steps:
        capped = None
        if hit_bound:
            pred = mu0 + ds * pdot
            if pred > p_max:
                ds, capped = (p_max - mu0) / pdot, p_max
            elif pred < p_min:
                ds, capped = (p_min - mu0) / pdot, p_min
        x = x0 + ds * xdot
        mu = mu0 + ds * pdot
        its = lin_total = 0
        ok = False
        while True:
            P.keo_fill(mu, theta)
            P.jac_rebuild(g, x)
            F = P.compute_f(g, x)
            gc = xdot @ (x - x0) / length + sc * sc * pdot * (mu - mu0) - ds
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
            dp = -(gc + xdot @ a / length) / (sc * sc * pdot + xdot @ b / length)
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
        xdot, pdot, pred_its, ratio = tangent(x, mu, xdot, pdot, 0.0)
        record(k, x, mu, its, lin_total, pred_prev, float(nrm), ds_used, pdot)
        fac = (nl_maxit - its) / (nl_maxit - 1.0)
        ds = min(ds * (1.0 + aggressiveness * fac * fac), ds_max)
        ds, ds_min, ds_max = ds * ratio, ds_min * ratio, ds_max * ratio
        if hit_bound:
            reached = capped if capped is not None else (p_max if mu > p_max else p_min if mu < p_min else None)
            if reached is not None:
                break
        elif mu > p_max or mu < p_min:
            break
        k += 1
    if hit_bound and reached is not None and mu != reached:
        # the last step: natural continuation to the bound itself, constant predictor
        P.keo_fill(reached, theta)
        x, steps, lin, fn = P.newton(g, x, nl_tol, nl_maxit, lin_tol, lin_maxit)
        if fn[-1] < nl_tol:
            record(k + 1, x, reached, steps, int(lin.sum()), 0, float(fn[-1]), reached - mu, pdot)
    return x, recs
