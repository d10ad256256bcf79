"""Chebyshev semi-iteration (oracle; test infrastructure only).

Restates PETSc ``KSPCHEBYSHEV`` + ``PCJACOBI`` as configured by the reference for
``solver_0`` (control/control.py:1967-1982): fixed eigenvalue bounds of ``D^-1 A``, no
estimation, ``max_it`` steps, tolerances 0.  PETSc itself is not in /root/reference (third
party, un-pinned); the recurrence follows PETSc's published first-kind Chebyshev
iteration.  The same recurrence, started from a non-zero guess,
is the smoother of the aggregation AMG (``oracle/amg.py``).
"""
import numpy as np

from .fastmv import mv


def chebyshev_coefficients(e_min, e_max, steps):
    """Scalars of the three-term recurrence: (scale, [omega_2 .. omega_steps])."""
    scale = 2.0 / (e_max + e_min)
    alpha = 1.0 - scale * e_min
    mu = 1.0 / alpha
    omegaprod = 2.0 / alpha
    c_prev, c_cur = 1.0, mu
    omegas = []
    for _ in range(2, steps + 1):
        c_next = 2.0 * mu * c_cur - c_prev
        omegas.append(omegaprod * c_cur / c_next)
        c_prev, c_cur = c_cur, c_next
    return scale, omegas


def chebyshev(A, dinv, b, e_min, e_max, steps, x0=None):
    """``steps`` Chebyshev iterations on ``D^-1 A`` for ``A x = b``.

    ``b`` (and ``x0``) may be a vector (n,) or a batch (n, cols) / handled column-wise.
    Zero initial guess when ``x0`` is None (steps-1 products with A), else steps products.
    """
    scale, omegas = chebyshev_coefficients(e_min, e_max, steps)
    d = dinv if b.ndim == 1 else dinv[:, None]
    if x0 is None:
        p_prev = np.zeros_like(b)
        r = b
    else:
        p_prev = x0
        r = b - mv(A, x0)
    p_cur = p_prev + scale * (d * r)
    for omega in omegas:
        r = b - mv(A, p_cur)
        p_next = (1.0 - omega) * p_prev + omega * p_cur + (omega * scale) * (d * r)
        p_prev, p_cur = p_cur, p_next
    return p_cur
