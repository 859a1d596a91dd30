"""Independent numpy statement of the reference's operator and CG, written from
the mathematics rather than from oracle/pdeode_oracle.c, so
that a transcription slip in either shows up as a disagreement.  Test-only."""
import numpy as np


def dfunc(M, p):
    if p.dSelect == 1:
        return np.full_like(M, p.delta)
    if p.dSelect == 2:
        return p.delta * M ** p.alpha
    return p.delta * M ** p.alpha / (1.0 - M) ** p.beta


def ffunc(M, C, p):
    eps = C * 1e-8
    fs = p.fSelect
    if fs == 1: return C / (p.kappa + C) - p.nu
    if fs == 2: return C / (p.kappa + C)
    if fs == 3: return C / (p.kappa * M + C + eps) - p.nu
    if fs == 4: return C / (p.kappa * M + C + eps)
    if fs == 5: return np.ones_like(M)
    return C / (p.kappa + C) * (1.0 - M / (C + eps)) - p.nu


def dense_operator(M, C, row, col, p):
    """Dense n x n matrix of the implicit operator, built cell by cell from
    the ghost-node description: every cell has four faces; a face that would
    leave the grid is reflected onto the opposite neighbour.  The reference's
    quirk (P:576 '>=') makes cell n-col (1-based) behave as a last-row cell."""
    n = row * col
    d = dfunc(M, p)
    f = ffunc(M, C, p)
    h = 0.5 / (p.xDel * p.xDel)
    A = np.zeros((n, n))
    for i0 in range(n):
        i1 = i0 + 1                      # 1-based
        top = i1 <= col
        bottom = i1 >= n - col           # quirk Q1 included
        left = (i1 % col) == 1
        right = (i1 % col) == 0
        faces = [
            (i0 + col) if top else (i0 - col),
            (i0 + 1) if left else (i0 - 1),
            (i0 - 1) if right else (i0 + 1),
            (i0 - col) if bottom else (i0 + col),
        ]
        for k in faces:
            a = h * (d[k] + d[i0])
            A[i0, k] -= a
            A[i0, i0] += a
        A[i0, i0] += -f[i0] + 1.0 / p.tDel
    return A


def cg_reference(A, b, x0, tol2, maxit):
    """CG with the reference's stop rule."""
    n = b.size
    x = x0.copy()
    r = b - A @ x
    rho = 0.0
    pvec = None
    it = 0
    while True:
        it += 1
        rho_old = rho
        normx = np.sqrt(np.sum(x * x)) / n
        rho = float(r @ r)
        pvec = r.copy() if it == 1 else r + (rho / rho_old) * pvec
        q = A @ pvec
        alfa = rho / float(pvec @ q)
        x = x + alfa * pvec
        r = r - alfa * q
        err2 = abs(alfa) * np.sqrt(np.sum(pvec * pvec)) / n
        if it > maxit or err2 < tol2 * normx:
            return x, it


def solve_c(Mprev, Mnew, Cprev, p):
    aux = 0.5 * p.tDel * p.gama
    g = Cprev * Mprev / (p.kappa + Cprev)
    b = p.kappa - Cprev + aux * (Mnew + g)
    c = -p.kappa * Cprev + aux * p.kappa * g
    return 0.5 * (-b + np.sqrt(b * b - 4.0 * c))
