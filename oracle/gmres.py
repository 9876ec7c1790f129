"""CPU restatement of restarted GMRES.  TEST INFRASTRUCTURE ONLY.

examples/conf.xml:104 selects Belos "Pseudo Block GMRES" (tolerance 1e-10, 1000 iterations, block size 1).
Belos is not in the reference tree (unpinned, no reference test pins an iteration count) => PARITY UNPINNED;
restated is GMRES(m) as Belos organises it: x0 = 0, two-pass iterated classical Gram-Schmidt ("ICGS", the
default orthogonalisation), Givens rotations, implicit relative residual |g_{j+1}| / ||r0|| <= tol, explicit
residual at every restart, right preconditioning (x = M y).  The device solver nosh_gmres
(nosh_b200/csrc/gmres.cu) is compared with this file.
"""
import numpy as np


def gmres(apply_A, apply_M, b, tol, maxit, restart=300):
    n = b.size
    x = np.zeros(n)
    m = min(restart, maxit) if maxit > 0 else 1
    r = b.copy()
    r0 = np.sqrt(r @ r)
    hist = [1.0]
    if r0 == 0.0:
        return x, 0, 0.0, hist
    beta = r0
    it = 0
    relres = 1.0
    converged = False
    while not converged and it < maxit and beta > 0.0:
        V = np.zeros((m + 1, n))
        H = np.zeros((m + 1, m))
        cs = np.zeros(m)
        sn = np.zeros(m)
        g = np.zeros(m + 1)
        V[0] = r / beta
        g[0] = beta
        j = 0
        while j < m and it < maxit:
            w = apply_A(apply_M(V[j]) if apply_M is not None else V[j])
            h1 = V[:j + 1] @ w
            w = w - V[:j + 1].T @ h1
            h2 = V[:j + 1] @ w
            w = w - V[:j + 1].T @ h2
            hn = np.sqrt(w @ w)
            H[:j + 1, j] = h1 + h2
            H[j + 1, j] = hn
            for i in range(j):
                a, c = H[i, j], H[i + 1, j]
                H[i, j] = cs[i] * a + sn[i] * c
                H[i + 1, j] = -sn[i] * a + cs[i] * c
            a, c = H[j, j], hn
            d = np.hypot(a, c)
            cs[j], sn[j] = (1.0, 0.0) if d == 0.0 else (a / d, c / d)
            H[j, j] = d
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            it += 1
            relres = abs(g[j + 1]) / r0
            hist.append(relres)
            if relres <= tol or hn == 0.0:
                converged = True
                j += 1
                break
            V[j + 1] = w / hn
            j += 1
        k = j
        y = np.zeros(k)
        for i in range(k - 1, -1, -1):
            y[i] = (g[i] - H[i, i + 1:k] @ y[i + 1:k]) / H[i, i]
        upd = V[:k].T @ y
        x = x + (apply_M(upd) if apply_M is not None else upd)
        if converged or it >= maxit:
            break
        r = b - apply_A(x)
        beta = np.sqrt(r @ r)
    return x, it, relres, hist
