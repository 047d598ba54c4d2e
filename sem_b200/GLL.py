"""Host-side 1-D Gauss-Legendre-Lobatto tables -- same public functions as the reference's ``Solvers/GLL.py``.

These O(P^2) tables are the only thing the GPU path takes from the host: they are uploaded once per polynomial order
into constant memory (``sem_ctx_create``).  Values follow the reference's formulas (file:line cited per function) so
the device operators agree with the reference's matrices to rounding; the code is an independent, vectorised and
cached implementation.
"""
import functools

import numpy as np


@functools.lru_cache(maxsize=None)
def _tables(P: int):
    """nodes, weights, Legendre values L_k(x_i) -- Newton iteration of GLL.py:13-30."""
    if P < 1:
        raise ValueError("polynomial order must be >= 1")
    idx = np.arange(P + 1)
    x = -np.cos(np.pi * idx / P)                      # Chebyshev-Lobatto start (GLL.py:15)
    L = np.zeros((P + 1, P + 1), dtype=np.float64)
    tiny = np.finfo(np.float64).eps
    delta = np.ones_like(x)
    while np.abs(delta).max() > tiny:                 # GLL.py:20-28
        L[:, 0] = 1.0
        L[:, 1] = x
        for k in range(2, P + 1):                     # Bonnet recursion
            L[:, k] = ((2 * k - 1) * x * L[:, k - 1] - (k - 1) * L[:, k - 2]) / k
        delta = -(x * L[:, P] - L[:, P - 1]) / ((P + 1) * L[:, P])
        x = x + delta
    w = 2. / (P * (P + 1) * L[:, P] ** 2)             # GLL.py:31
    for a in (x, w, L):
        a.setflags(write=False)
    return x, w, L


def standard_nodes(P: int):
    """Quadrature nodes in [-1,1], weights and the Legendre Vandermonde matrix  (GLL.py:7-33)."""
    x, w, L = _tables(P)
    return x.copy(), w.copy(), L.copy()


def standard_mass_matrix(P: int):
    """diag(w)  (GLL.py:36-42)."""
    return np.diag(_tables(P)[1])


@functools.lru_cache(maxsize=None)
def _diff(P: int):
    x, _, L = _tables(P)
    LP = L[:, -1]
    with np.errstate(divide="ignore", invalid="ignore"):
        D = (LP[:, None] / LP[None, :]) * 1 / (x[:, None] - x[None, :])   # GLL.py:56
    D[np.diag_indices(P + 1)] = 0.0
    D[0, 0] = -P * (P + 1) / 4                                            # GLL.py:57-58
    D[-1, -1] = P * (P + 1) / 4
    D.setflags(write=False)
    return D


def standard_differentiation_matrix(P: int):
    """D[i, j] = l_j'(xi_i)  (GLL.py:45-59)."""
    return _diff(P).copy()


def standard_gradient_matrix(P: int):
    """G[i, j] = w_i D[i, j]  (GLL.py:62-70)."""
    return _tables(P)[1][:, None] * _diff(P)


@functools.lru_cache(maxsize=None)
def _stiff(P: int):
    w, D = _tables(P)[1], _diff(P)
    K = np.einsum('k,ki,kj->ij', w, D, D)             # same contraction order as GLL.py:81
    K.setflags(write=False)
    return K


def standard_stiffness_matrix(P: int):
    """K[i, j] = sum_k w_k D[k, i] D[k, j]  (GLL.py:73-81)."""
    return _stiff(P).copy()


def standard_product_matrix(P: int):
    """F[i, j, k] = w_i delta_ij delta_ik  (GLL.py:84-91)."""
    F = np.zeros((P + 1,) * 3)
    r = np.arange(P + 1)
    F[r, r, r] = _tables(P)[1]
    return F


def standard_convection_matrix(P: int):
    """C[i, j, k] = w_i delta_ij D[i, k]  (GLL.py:94-102)."""
    Cm = np.zeros((P + 1,) * 3)
    r = np.arange(P + 1)
    Cm[r, r, :] = standard_gradient_matrix(P)
    return Cm


def standard_evaluation_matrix(P: int, xi: np.ndarray):
    """S[i, j] = l_j(xi[i]) as the product formula of GLL.py:105-116."""
    x = _tables(P)[0]
    xi = np.asarray(xi, dtype=np.float64).ravel()
    S = np.ones((xi.size, P + 1))
    for j in range(P + 1):
        for k in range(P + 1):
            if k != j:
                S[:, j] *= (xi - x[k]) / (x[j] - x[k])
    return S
