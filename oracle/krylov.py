"""Krylov methods with PETSc KSP conventions (oracle; test infrastructure only).

Restates what ``MultiBlockSystem.solve`` asks of PETSc
(preconditioner/preconditioner.py:732-772): ``KSP`` of type ``gmres`` / ``fgmres`` /
``minres``, ``setInitialGuessNonzero(True)`` (743), tolerances (738-742), restart (747-748),
monitor called with it = 0 for the initial residual (749-754), and the converged-reason
policy (768-772).  PETSc itself is third party, absent from /root/reference and
un-pinned; the conventions restated here are PETSc's documented defaults
(SURVEY.md Appendix A.3): default convergence test with reference norm ||b||
(||P^-1 b|| for the preconditioned norm), classical Gram-Schmidt without refinement,
restart 30, left PC + preconditioned norm for gmres/minres, right PC + true residual norm
for fgmres.  Vectors are flat numpy arrays; ``A`` and ``pc`` are callables.
"""
import numpy as np

CONVERGED_RTOL = 2
CONVERGED_ATOL = 3
CONVERGED_HAPPY_BREAKDOWN = 7
DIVERGED_ITS = -3
DIVERGED_DTOL = -4
DIVERGED_BREAKDOWN = -5
DIVERGED_NANORINF = -9
DIVERGED_INDEFINITE_PC = -8


class KSPResult:
    def __init__(self):
        self.its = 0
        self.reason = 0
        self.history = []          # residual norms as the KSP monitor sees them (it = 0 first)
        self.n_mult = 0
        self.n_pc = 0

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its


class _Conv:
    """KSPConvergedDefault."""

    def __init__(self, rtol, atol, divtol, ref_norm):
        self.rtol, self.atol, self.divtol = rtol, atol, divtol
        self.ttol = max(rtol * ref_norm, atol)
        self.rnorm0 = ref_norm

    def __call__(self, rnorm):
        if not np.isfinite(rnorm):
            return DIVERGED_NANORINF
        if rnorm <= self.ttol:
            return CONVERGED_ATOL if rnorm < self.atol else CONVERGED_RTOL
        if rnorm >= self.divtol * self.rnorm0:
            return DIVERGED_DTOL
        return 0


def gmres(A, b, x0, *, pc=None, flexible=False, restart=30, rtol=1e-5, atol=1e-50,
          divtol=1e4, max_it=1000, monitor=None):
    """GMRES(restart) (left PC, preconditioned norm) or FGMRES(restart) (right PC, true
    residual norm).  Returns (x, KSPResult)."""
    res = KSPResult()
    if pc is None:
        def pc(v):
            return v.copy()

    def op(v):
        res.n_mult += 1
        return A(v)

    def prec(v):
        res.n_pc += 1
        return pc(v)

    x = x0.copy()
    n = b.size
    ref = np.linalg.norm(b) if flexible else np.linalg.norm(prec(b))
    conv = _Conv(rtol, atol, divtol, ref)
    V = np.zeros((restart + 1, n))
    Z = np.zeros((restart, n)) if flexible else None
    H = np.zeros((restart + 1, restart))
    while res.reason == 0:
        r = b - op(x)
        if not flexible:
            r = prec(r)
        rnorm = np.linalg.norm(r)
        # KSPGMRESCycle: convergence test on the (re)computed residual of every cycle;
        # the history keeps one entry per iteration number (PETSc re-monitors the same
        # `its` with the recomputed norm at a restart; that duplicate is not recorded)
        if res.its == 0:
            res.history.append(rnorm)
            if monitor:
                monitor(res.its, rnorm)
        res.reason = conv(rnorm)
        if res.reason:
            break
        if rnorm == 0.0:
            res.reason = CONVERGED_ATOL
            break
        V[0] = r / rnorm
        g = np.zeros(restart + 1)
        g[0] = rnorm
        cs = np.zeros(restart)
        sn = np.zeros(restart)
        H[:] = 0.0
        it = 0
        while res.reason == 0 and it < restart and res.its < max_it:
            if flexible:
                Z[it] = prec(V[it])
                w = op(Z[it])
            else:
                w = prec(op(V[it]))
            # classical Gram-Schmidt: all dots against the unmodified w, then one update
            h = V[:it + 1] @ w
            w = w - h @ V[:it + 1]
            H[:it + 1, it] = h
            tt = np.linalg.norm(w)
            H[it + 1, it] = tt
            happy = tt == 0.0
            if not happy:
                V[it + 1] = w / tt
            # apply previous rotations, form the new one
            for k in range(it):
                t = H[k, it]
                H[k, it] = cs[k] * t + sn[k] * H[k + 1, it]
                H[k + 1, it] = -sn[k] * t + cs[k] * H[k + 1, it]
            denom = np.hypot(H[it, it], H[it + 1, it])
            if denom == 0.0:
                res.reason = DIVERGED_BREAKDOWN
                break
            cs[it] = H[it, it] / denom
            sn[it] = H[it + 1, it] / denom
            H[it, it] = denom
            H[it + 1, it] = 0.0
            g[it + 1] = -sn[it] * g[it]
            g[it] = cs[it] * g[it]
            rnorm = abs(g[it + 1])
            it += 1
            res.its += 1
            res.history.append(rnorm)
            if monitor:
                monitor(res.its, rnorm)
            res.reason = conv(rnorm)
            if happy and res.reason == 0:
                res.reason = CONVERGED_HAPPY_BREAKDOWN
        if res.its >= max_it and res.reason == 0:
            res.reason = DIVERGED_ITS
        if it > 0:
            y = np.linalg.solve(np.triu(H[:it, :it]), g[:it])
            x = x + (y @ Z[:it] if flexible else y @ V[:it])
    return x, res


def minres(A, b, x0, *, pc=None, rtol=1e-5, atol=1e-50, divtol=1e4, max_it=1000,
           monitor=None):
    """Preconditioned MINRES (Paige & Saunders), SPD preconditioner, monitored norm =
    sqrt(r^T P^-1 r) (PETSc: left PC, preconditioned norm).  Returns (x, KSPResult)."""
    res = KSPResult()
    if pc is None:
        def pc(v):
            return v.copy()

    def op(v):
        res.n_mult += 1
        return A(v)

    def prec(v):
        res.n_pc += 1
        return pc(v)

    x = x0.copy()
    yb = prec(b)
    bb = float(b @ yb)
    if bb < 0.0:
        res.reason = DIVERGED_INDEFINITE_PC
        return x, res
    conv = _Conv(rtol, atol, divtol, np.sqrt(bb))
    r1 = b - op(x)
    y = prec(r1)
    beta1 = float(r1 @ y)
    if beta1 < 0.0:
        res.reason = DIVERGED_INDEFINITE_PC
        return x, res
    beta1 = np.sqrt(beta1)
    res.history.append(beta1)
    if monitor:
        monitor(0, beta1)
    res.reason = conv(beta1)
    if res.reason or beta1 == 0.0:
        if beta1 == 0.0 and res.reason == 0:
            res.reason = CONVERGED_ATOL
        return x, res
    oldb = 0.0
    beta = beta1
    dbar = 0.0
    epsln = 0.0
    phibar = beta1
    cs, sn = -1.0, 0.0
    w = np.zeros_like(b)
    w2 = np.zeros_like(b)
    r2 = r1.copy()
    eps = np.finfo(float).eps
    while res.reason == 0 and res.its < max_it:
        s = 1.0 / beta
        v = s * y
        y = op(v)
        if res.its >= 1:
            y = y - (beta / oldb) * r1
        alfa = float(v @ y)
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        y = prec(r2)
        oldb = beta
        beta = float(r2 @ y)
        if beta < 0.0:
            res.reason = DIVERGED_INDEFINITE_PC
            break
        beta = np.sqrt(beta)
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        gamma = max(np.hypot(gbar, beta), eps)
        cs = gbar / gamma
        sn = beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) / gamma
        x = x + phi * w
        res.its += 1
        rnorm = abs(phibar)
        res.history.append(rnorm)
        if monitor:
            monitor(res.its, rnorm)
        res.reason = conv(rnorm)
        if beta == 0.0 and res.reason == 0:
            res.reason = CONVERGED_HAPPY_BREAKDOWN
    if res.its >= max_it and res.reason == 0:
        res.reason = DIVERGED_ITS
    return x, res
