"""Continuous-Galerkin SEM on the uniform Cartesian element grid -- public functions of the reference's
``Solvers/SEM.py`` on top of the GPU path.

Mesh bookkeeping (nodes, index map) stays on the host.  Assembly (``assemble`` of 4-index arrays), ``scatter`` and the
global operators run on the device: ``global_*_matrix`` return matrix-free operator objects whose ``@`` launches the
fused sum-factorised kernel instead of materialising SciPy/``sparse`` matrices (the dense ``(P+1)^6``-per-element
temporaries of SEM.py:243-244 are never formed).
"""
import typing

import numpy as np
import torch

from . import GLL
from .device import SemDevice


def xi2x(e, xi, dx):
    """Physical coordinate of standard coordinate xi in element e  (SEM.py:11-20)."""
    if np.any(np.asarray(xi) > 1) or np.any(np.asarray(xi) < -1):
        raise ValueError('xi out of range')
    return dx / 2 * (xi + 1) + dx * e


def x2xi(x, dx):
    """(element, xi) of a physical coordinate; shared nodes go to the left element  (SEM.py:23-36)."""
    frac, e = np.modf(np.asarray(x, dtype=np.float64) / dx)
    xi = 2 * frac - 1
    shift = np.isclose(xi, -1) * (e > 0)
    e[shift] -= 1
    xi[shift] = 1
    return e.astype(int), xi


def element_nodes_1d(P, N_ex, dx):
    """x[m, k]  (SEM.py:39-48)."""
    xi = GLL.standard_nodes(P)[0]
    return np.vstack([xi2x(m, xi, dx) for m in range(N_ex)])


def global_nodes_1d(P, N_ex, dx):
    """Global 1-D nodes: element nodes without the duplicated first node  (SEM.py:51-60)."""
    xe = element_nodes_1d(P, N_ex, dx)
    return np.insert(np.ravel(xe[:, 1:]), 0, 0)


def element_nodes(P, N_ex, N_ey, dx, dy):
    """[x[m,n,k,l], y[m,n,k,l]]  (SEM.py:63-79), built by broadcasting instead of the (m, n) loop."""
    xe = element_nodes_1d(P, N_ex, dx)
    ye = element_nodes_1d(P, N_ey, dy)
    pts = np.empty((2, N_ex, N_ey, P + 1, P + 1))
    pts[0] = xe[:, None, :, None]
    pts[1] = ye[None, :, None, :]
    return pts


def global_nodes(P, N_ex, N_ey, dx, dy):
    """[x_p, y_p], x slow / y fast  (SEM.py:82-94)."""
    x1 = global_nodes_1d(P, N_ex, dx)
    y1 = global_nodes_1d(P, N_ey, dy)
    return np.reshape(np.array(np.meshgrid(x1, y1, indexing='ij')), (2, x1.size * y1.size))


def global_index(P, N_ex, N_ey, m, n, i, j):
    """Local -> global index  (SEM.py:97-110)."""
    if np.any(m >= N_ex) or np.any(n >= N_ey) or np.any(i > P) or np.any(j > P):
        raise ValueError('Indices out of range')
    return n * P + j + (N_ey * P + 1) * (m * P + i)


def _device_for(P, N_ex, N_ey, dx=1.0, dy=1.0, _cache={}):
    key = (P, N_ex, N_ey, float(dx), float(dy))
    if key not in _cache:
        _cache[key] = SemDevice(P, N_ex, N_ey, dx, dy)
    return _cache[key]


def assemble(A_e: np.ndarray):
    """Gather-scatter of an element array into global numbering  (SEM.py:113-146).

    4-index ``A[m,n,i,j]`` -> global vector: the colour-ordered, atomic-free device kernel (bitwise reproducible).
    6-index ``A[m,n,i,j,k,l]`` -> SciPy CSR matrix and 8-index ``A[m,n,i,j,r,s,k,l]`` -> a sparse 3-tensor (:class:`COO3`, the
    stand-in for ``sparse.COO``; contract it with :func:`tensordot`): host-side convenience paths for callers outside the two
    solvers -- the solvers themselves never form these arrays (``global_*_matrix`` return matrix-free operators).
    """
    A_e = np.asarray(A_e, dtype=np.float64)
    if A_e.ndim not in (4, 6, 8):
        raise ValueError("assemble expects a 4-, 6- or 8-index element array")
    N_ex, N_ey, P = A_e.shape[0], A_e.shape[1], A_e.shape[2] - 1
    if A_e.ndim == 4:
        dev = _device_for(P, N_ex, N_ey)
        elem = torch.from_numpy(np.ascontiguousarray(A_e)).to(dev.tdev)
        return dev.to_host(dev.gather_scatter(elem))
    import scipy.sparse as sp_sparse
    N = (P * N_ex + 1) * (P * N_ey + 1)
    nz = np.nonzero(A_e)
    m, n = nz[0], nz[1]
    idx = [global_index(P, N_ex, N_ey, m, n, nz[2 + 2 * t], nz[3 + 2 * t]) for t in range(A_e.ndim // 2 - 1)]
    data = A_e[nz]
    if A_e.ndim == 6:                                  # duplicates (shared nodes) are summed by the conversion
        return sp_sparse.coo_matrix((data, (idx[0], idx[1])), shape=(N, N)).tocsr()
    return COO3(np.vstack(idx), data, (N, N, N))


class COO3:
    """Minimal sparse 3-tensor in coordinate format (duplicates are summed on contraction): what ``SEM.assemble`` returns for
    8-index element arrays in place of ``sparse.COO`` (SEM.py:139-145; the ``sparse`` package is not a dependency here)."""

    def __init__(self, coords, data, shape):
        self.coords, self.data, self.shape = np.asarray(coords), np.asarray(data, dtype=np.float64), tuple(shape)

    def tensordot(self, v, axis):
        """Contract index ``axis`` (0, 1 or 2) with the vector ``v`` -> SciPy CSR matrix over the remaining two indices."""
        import scipy.sparse as sp_sparse
        v = np.asarray(v, dtype=np.float64)
        keep = [a for a in range(3) if a != axis]
        return sp_sparse.coo_matrix((self.data * v[self.coords[axis]], (self.coords[keep[0]], self.coords[keep[1]])),
                                    shape=(self.shape[keep[0]], self.shape[keep[1]])).tocsr()

    def __matmul__(self, v):          # numpy semantics of C @ v: the last index
        return self.tensordot(v, 2)

    def __rmatmul__(self, v):         # numpy semantics of v @ C: the second-to-last index
        return self.tensordot(v, 1)


def tensordot(C, v, axes, return_type=None):
    """``sparse.tensordot(C, v, (axis, 0), return_type=sparse.COO)`` of the reference (CD:82-83,101-102; NS:103-104,130-136) for
    the 3-tensors of this package (:class:`COO3`, :class:`_Convection3`): contracts index ``axes[0]`` of ``C`` with ``v``."""
    return C.tensordot(v, int(axes[0]))


def scatter(u: np.ndarray, P: int, N_ex: int, N_ey: int):
    """Global vector -> element array u[m,n,i,j]  (SEM.py:149-167) through the device scatter kernel."""
    u = np.asarray(u, dtype=np.float64)
    if u.shape[0] != (P * N_ex + 1) * (P * N_ey + 1):
        raise ValueError('Not a valid combination of global coefficients vector, P, N_ex, and N_ey')
    dev = _device_for(P, N_ex, N_ey)
    return dev.scatter(dev.to_device(u)).cpu().numpy()


class _MatrixFree:
    """Stand-in for the CSR matrices of SEM.py:170-223: ``op @ x`` launches the fused device kernel."""

    def __init__(self, dev, kind):
        self._dev, self._kind = dev, kind
        self.shape = (dev.NX * dev.NY,) * 2

    def __matmul__(self, x):
        dev = self._dev
        xd = dev.to_device(x)
        y = dev.zeros()
        if self._kind == 'M':
            dev.apply_mass(xd, y)
        elif self._kind == 'K':
            dev.apply_stiffness(xd, y)
        elif self._kind == 'Gx':
            dev.apply_gradient(xd, y, None)
        else:
            dev.apply_gradient(xd, None, y)
        return dev.to_host(y)

    def diagonal(self):
        if self._kind != 'M':
            raise NotImplementedError
        return self._dev.to_host(self._dev.mass_diag())


def global_mass_matrix(P, N_ex, N_ey, dx, dy):
    """Matrix-free M  (SEM.py:170-183)."""
    return _MatrixFree(_device_for(P, N_ex, N_ey, dx, dy), 'M')


def global_stiffness_matrix(P, N_ex, N_ey, dx, dy):
    """Matrix-free K  (SEM.py:186-203)."""
    return _MatrixFree(_device_for(P, N_ex, N_ey, dx, dy), 'K')


def global_gradient_matrices(P, N_ex, N_ey, dx, dy):
    """Matrix-free G_x, G_y  (SEM.py:206-223)."""
    dev = _device_for(P, N_ex, N_ey, dx, dy)
    return _MatrixFree(dev, 'Gx'), _MatrixFree(dev, 'Gy')


class _RowScaled:
    """``diag(u) G``: what ``u @ C`` is for a convection tensor; ``@ x`` launches the matrix-free gradient."""

    def __init__(self, u, G):
        self._u, self._G, self.shape = np.asarray(u, dtype=np.float64), G, G.shape

    def __matmul__(self, x):
        return self._u * (self._G @ x)


class _Convection3:
    """Stand-in for the N x N x N convection tensors of SEM.py:226-245.  With the Kronecker deltas of GLL.py:84-102 the tensor
    is diagonal in its first two indices, ``C[a, b, c] = delta_ab G[a, c]``, so it is never formed:
    ``tensordot(C, u, (1, 0)) == u @ C == diag(u) G`` (matrix-free operator) and ``tensordot(C, T, (2, 0)) == C @ T ==
    diag(G T)`` (SciPy sparse diagonal matrix)."""

    def __init__(self, G):
        self._G = G
        self.shape = (G.shape[0],) * 3

    def tensordot(self, v, axis):
        import scipy.sparse as sp_sparse
        if axis == 1:
            return _RowScaled(v, self._G)
        if axis == 2:
            return sp_sparse.diags(self._G @ v, format='csr')
        raise NotImplementedError("a convection tensor is contracted over its velocity (1) or its field (2) index")

    def __matmul__(self, T):
        return self.tensordot(T, 2)

    def __rmatmul__(self, u):
        return self.tensordot(u, 1)

    __array_ufunc__ = None            # let ``ndarray @ C`` reach __rmatmul__


def global_convection_matrices(P, N_ex, N_ey, dx, dy):
    """C_x, C_y of SEM.py:226-245 as matrix-free stand-ins (:class:`_Convection3`); contract them with :func:`tensordot`, ``@``
    or ``u @ C``.  The solvers use the same identities inside the fused kernels."""
    Gx, Gy = global_gradient_matrices(P, N_ex, N_ey, dx, dy)
    return _Convection3(Gx), _Convection3(Gy)


def interp_matrix_1d(P: int, N_e: int, h: float, pts: np.ndarray) -> np.ndarray:
    """Dense 1-D interpolation matrix I[a, q] = l_q(pts[a]) onto arbitrary points from the N_e*P+1 global nodes of a line of
    N_e elements of size h: the Lagrange basis of the element that holds the point (x2xi, SEM.py:23-36; evaluation matrix
    GLL.py:105-116), zero elsewhere.  The SEM interpolant on an ij-meshgrid is then I_x F I_y^T (SEM.py:248-273)."""
    pts = np.asarray(pts, dtype=np.float64).ravel()
    e, xi = x2xi(pts, h)
    e = np.clip(e, 0, N_e - 1)
    S = GLL.standard_evaluation_matrix(P, xi)                    # [a, k]
    I = np.zeros((pts.size, N_e * P + 1))
    cols = e[:, None] * P + np.arange(P + 1)[None, :]
    np.put_along_axis(I, cols, S, axis=1)
    return I


def eval_interpolation(u_e: np.ndarray, points_e: np.ndarray, points_plot: typing.Tuple[np.ndarray, np.ndarray]):
    """Evaluate the SEM interpolant of u_e[m,n,k,l] on an ij-meshgrid  (SEM.py:248-273).

    Host-side tensor-product evaluation, one gathered einsum instead of the (m, n) double loop.
    """
    P = u_e.shape[2] - 1
    x_e = points_e[0, :, 0, :, 0]
    y_e = points_e[1, 0, :, 0, :]
    dx = x_e[0, -1] - x_e[0, 0]
    dy = y_e[0, -1] - y_e[0, 0]
    m_p, xi_p = x2xi(points_plot[0][:, 0], dx)
    n_p, eta_p = x2xi(points_plot[1][0, :], dy)
    Sx = GLL.standard_evaluation_matrix(P, xi_p)      # [a, k]
    Sy = GLL.standard_evaluation_matrix(P, eta_p)     # [b, l]
    tmp = np.einsum('ankl,ak->anl', u_e[m_p], Sx)     # contract x inside each plot column's element column
    return np.einsum('abl,bl->ab', tmp[:, n_p, :], Sy)


# ---- boundary block of the pressure rows (experimental NS preconditioner stage) ---------------------
def assembled_1d(A_s: np.ndarray, N_e: int) -> np.ndarray:
    """Dense 1-D assembly of N_e copies of a (P+1) x (P+1) element matrix with one shared node (the x or y factor of the
    Kronecker-structured global matrices of SEM.py:170-223)."""
    n = A_s.shape[0]
    P = n - 1
    A = np.zeros((N_e * P + 1, N_e * P + 1))
    for m in range(N_e):
        A[m * P:m * P + n, m * P:m * P + n] += A_s
    return A


def pressure_boundary_block(P: int, N_ex: int, N_ey: int, dx: float, dy: float, pin: int = None):
    """Nodes and matrix of the pressure-Neumann rows restricted to the boundary: returns ``(ix, iy, K_BB)`` with the boundary
    nodes of the mesh (W line, E line, then the S / N nodes of the lines in between; the node with global index ``pin`` is
    left out) and ``K_BB[a, b] = K[node_a, node_b]`` for ``K = K1x (x) M1y + M1x (x) K1y`` (SEM.py:186-203, the rows the
    reference puts into the continuity block at boundary nodes, NS:119,157)."""
    NX, NY = N_ex * P + 1, N_ey * P + 1
    Ks, w = GLL.standard_stiffness_matrix(P), GLL.standard_nodes(P)[1]
    K1x, K1y = assembled_1d(2.0 / dx * Ks, N_ex), assembled_1d(2.0 / dy * Ks, N_ey)
    M1x, M1y = np.diag(assembled_1d(dx / 2.0 * np.diag(w), N_ex)), np.diag(assembled_1d(dy / 2.0 * np.diag(w), N_ey))
    ix = np.concatenate([np.zeros(NY, int), np.full(NY, NX - 1), np.repeat(np.arange(1, NX - 1), 2)])
    iy = np.concatenate([np.arange(NY), np.arange(NY), np.tile([0, NY - 1], NX - 2)])
    if pin is not None:
        keep = (iy + NY * ix) != pin
        ix, iy = ix[keep], iy[keep]
    same_x = ix[:, None] == ix[None, :]
    same_y = iy[:, None] == iy[None, :]
    KBB = K1x[ix[:, None], ix[None, :]] * (M1y[iy][:, None] * same_y) + (M1x[ix][:, None] * same_x) * K1y[iy[:, None], iy[None, :]]
    return ix, iy, KBB


def ring_chebyshev_parameters(P: int, N_ex: int, N_ey: int, dx: float, dy: float, pin: int = None, target: float = 0.03,
                              max_steps: int = 8):
    """Spectrum bounds ``(lo, hi)`` of the diagonally scaled boundary-ring block ``D^-1/2 K_BB D^-1/2`` of the pressure rows
    (rows K[mask,:] of NS:119,157, pin node left out) and the degree of the fixed Chebyshev polynomial that replaces its inverse
    in the NS preconditioner: the smallest degree whose error bound ``2 r^k / (1 + r^2k)``, ``r = (sqrt(hi/lo) - 1) /
    (sqrt(hi/lo) + 1)``, is below ``target`` (3 for square elements, where the spectrum lies in about [0.56, 1.50] for every
    mesh and order), at most ``max_steps``.  The block is assembled sparsely on the host from the 1-D matrices (O(ring P)
    entries) and its extreme eigenvalues come from Lanczos; a 2 % margin is added on both sides."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    NX, NY = N_ex * P + 1, N_ey * P + 1
    Ks, w = GLL.standard_stiffness_matrix(P), GLL.standard_nodes(P)[1]

    def line(A_s, N_e):
        n = A_s.shape[0]
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
        rows = np.concatenate([(ii + m * P).ravel() for m in range(N_e)])
        cols = np.concatenate([(jj + m * P).ravel() for m in range(N_e)])
        return sps.coo_matrix((np.tile(A_s.ravel(), N_e), (rows, cols)), shape=(N_e * P + 1,) * 2).tocsr()

    K1x, K1y = line(2.0 / dx * Ks, N_ex), line(2.0 / dy * Ks, N_ey)
    M1x, M1y = line(dx / 2.0 * np.diag(w), N_ex).diagonal(), line(dy / 2.0 * np.diag(w), N_ey).diagonal()
    ix = np.concatenate([np.zeros(NY, int), np.full(NY, NX - 1), np.repeat(np.arange(1, NX - 1), 2)])
    iy = np.concatenate([np.arange(NY), np.arange(NY), np.tile([0, NY - 1], NX - 2)])
    if pin is not None:
        keep = (iy + NY * ix) != pin
        ix, iy = ix[keep], iy[keep]
    nb = ix.size
    # selection matrices ring <- x line / y line; K_BB = sum over the y values {0, NY-1, ...} ... assembled through the masks
    Sx = sps.csr_matrix((np.ones(nb), (np.arange(nb), ix)), shape=(nb, NX))
    Sy = sps.csr_matrix((np.ones(nb), (np.arange(nb), iy)), shape=(nb, NY))
    same_y = (Sy @ Sy.T).tocsr()                       # 1 where two ring nodes share iy
    same_x = (Sx @ Sx.T).tocsr()
    A = (Sx @ K1x @ Sx.T).multiply(same_y).multiply(sps.csr_matrix(M1y[iy][:, None])) \
        + (Sy @ K1y @ Sy.T).multiply(same_x).multiply(sps.csr_matrix(M1x[ix][:, None]))
    A = sps.csr_matrix(A)
    sd = 1.0 / np.sqrt(A.diagonal())
    B = sps.diags(sd) @ A @ sps.diags(sd)
    B = 0.5 * (B + B.T)
    if nb <= 400:
        ev = np.linalg.eigvalsh(B.toarray())
        lo, hi = ev[0], ev[-1]
    else:
        v0 = np.cos(np.arange(nb) * 0.7) + 1.5          # fixed start vector: reproducible bounds
        hi = spla.eigsh(B, k=1, which='LA', v0=v0, tol=1e-10, return_eigenvectors=False)[0]
        lo = spla.eigsh(B, k=1, which='SA', v0=v0, tol=1e-10, return_eigenvectors=False)[0]
    lo, hi = 0.98 * float(lo), 1.02 * float(hi)
    r = (np.sqrt(hi / lo) - 1.0) / (np.sqrt(hi / lo) + 1.0)
    steps = 1
    while steps < max_steps and 2.0 * r ** steps / (1.0 + r ** (2 * steps)) > target:
        steps += 1
    return lo, hi, steps
