"""CPU mirror (numpy) of the device Navier-Stokes preconditioner -- TEST INFRASTRUCTURE, never imported by the product.

The device solver (sem_b200/csrc/sem_capi.cu, ``ns_precond_apply``) solves the 3-field linearised system of NS:138-160 by
right-preconditioned GMRES with the block lower-triangular preconditioner  [[P_a, 0], [C, S~]]^-1.  This module restates
every stage of it with the oracle's matrices and the same matrix-free building blocks (1-D eigen-decompositions, ring
Chebyshev, tensor-product projectors), stage by stage, so that the GPU tests can compare ONE application of every stage
with the CPU (tests/test_gpu_parity.py) and so that iteration counts can be studied without a GPU.

Stages of  z_p = S~^-1 y  (``level`` as in ``sem_krylov.precond``):
  2 'fdm'      z_p = y / M_p                                            (the reference's own Schur preconditioner, NS:208-212)
  3 'fdm+bb'   interior as above, boundary ring by block elimination   z_B = K_BB^-1 (y_B - K_BI z_I)  (fixed Chebyshev polynomial)
  4 'full'     two-level: z1 = Pi Shat^+ Pi^T y on the interior (+ ring), r1 = y - S_0 z1,
               z2 = M^-1 F_p K_N^+ r1 on the interior (+ ring) (pressure convection-diffusion), z = z1 + z2,
               then the rank-one member-selection correction  z -= l_c (m_c.z - l_c.y) / (m_c.l_c).
"""
import numpy as np
import scipy.linalg as sla
from numpy.polynomial import legendre as npl

from . import sem_oracle as so

import scipy.sparse as sps
import scipy.sparse.linalg as spla

CHEB_TARGET, CHEB_MAX_STEPS = 0.03, 8


def ring_chebyshev(K, ring):
    """Spectrum bounds (2 % margin) of the diagonally scaled ring block K[ring][:, ring] and the Chebyshev degree whose error
    bound 2 r^k / (1 + r^2k) is below CHEB_TARGET (at most CHEB_MAX_STEPS) -- what sem_b200.SEM.ring_chebyshev_parameters
    computes from the 1-D matrices."""
    idx = np.where(ring)[0]
    A = K[idx][:, idx].tocsr()
    sd = 1.0 / np.sqrt(A.diagonal())
    B = sps.diags(sd) @ A @ sps.diags(sd)
    B = 0.5 * (B + B.T)
    if idx.size <= 400:
        ev = np.linalg.eigvalsh(B.toarray())
        lo, hi = ev[0], ev[-1]
    else:
        hi = spla.eigsh(B, k=1, which='LA', tol=1e-10, return_eigenvectors=False)[0]
        lo = spla.eigsh(B, k=1, which='SA', tol=1e-10, return_eigenvectors=False)[0]
    lo, hi = 0.98 * float(lo), 1.02 * float(hi)
    r = (np.sqrt(hi / lo) - 1.0) / (np.sqrt(hi / lo) + 1.0)
    steps = 1
    while steps < CHEB_MAX_STEPS and 2.0 * r ** steps / (1.0 + r ** (2 * steps)) > CHEB_TARGET:
        steps += 1
    return lo, hi, steps


def _dense_1d(P, ne, h):
    M1 = so._assembled_1d(h / 2 * so.mass_1d(P), ne).toarray().diagonal().copy()
    K1 = so._assembled_1d(2 / h * so.stiff_1d(P), ne).toarray()
    G1 = so._assembled_1d(so.grad_1d(P), ne).toarray()
    return M1, K1, G1


def pencil_eig(K, M, lo, hi):
    """Generalised eigenpairs of (K[lo:hi, lo:hi], diag(M[lo:hi])): Q [n, hi-lo] with zero rows outside, lam."""
    n = K.shape[0]
    s = 1.0 / np.sqrt(M[lo:hi])
    A = K[lo:hi, lo:hi] * s[:, None] * s[None, :]
    lam, V = np.linalg.eigh(0.5 * (A + A.T))
    Q = np.zeros((n, hi - lo))
    Q[lo:hi] = V * s[:, None]
    lam[np.abs(lam) < 1e-11 * np.abs(lam).max()] = 0.0
    return Q, lam


class Dir1D:
    """1-D pieces of one direction."""

    def __init__(self, P, ne, h):
        self.P, self.ne, self.h, self.n = P, ne, h, ne * P + 1
        self.M, self.K, self.G = _dense_1d(P, ne, h)
        n = self.n
        self.Qd, self.ld = pencil_eig(self.K, self.M, 1, n - 1)      # Dirichlet both ends (velocity block)
        self.Qn, self.ln = pencil_eig(self.K, self.M, 0, n)          # Neumann both ends (pressure Laplacian)
        # near-null pattern s = L_P(xi) element by element (sign alternating for odd P), hats on the element vertices
        xi = so.gll(P)[0]
        LP = npl.legval(xi, [0] * P + [1])
        self.s = np.zeros(n)
        for m in range(ne):
            self.s[m * P:m * P + P + 1] = (1.0 if P % 2 == 0 else (-1.0) ** m) * LP
        xn = (xi + 1) / 2
        W = np.zeros((n, ne + 1))
        for m in range(ne):
            W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1 - xn)
            W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xn)
        self.W = W * self.s[:, None]
        # coarse pencil on the interior nodes: R(sigma) = G E (K_II + sigma M_II)^-1 E^T G^T, sigma = lambda_s / 4
        self.lam_s = (self.s @ self.K @ self.s) / (self.s @ (self.M * self.s))
        KII = self.K[1:-1, 1:-1]
        MII = self.M[1:-1]
        X = np.linalg.solve(KII + 0.25 * self.lam_s * np.diag(MII), self.G[:, 1:-1].T)     # (K_II + s M_II)^-1 E^T G^T
        R = self.G[:, 1:-1] @ X
        self.R = 0.5 * (R + R.T)
        Rf = np.zeros((n, n))
        Rf[1:-1, 1:-1] = self.R[1:-1, 1:-1]
        self.Qc, self.lc_ = pencil_eig(Rf, self.M, 1, n - 1)          # pencil (E^T R E, M_II)
        # projector onto span(W) restricted to the interior nodes, M_II-orthogonal: P_W = W_I T^-1 W_I^T M_II
        WI = self.W[1:-1]
        self.WI = WI
        self.T = WI.T @ (MII[:, None] * WI)
        self.Tinv = np.linalg.pinv(self.T)
        # 1-D factor of the pressure part of the Jacobian's left null vector: 1 - L_P element by element for even P,
        # 1 + (-1)^m L_P for odd P (continuous; vanishes at both ends of the line only when the element count is even)
        self.lfac = np.zeros(n)
        for m in range(ne):
            self.lfac[m * P:m * P + P + 1] = 1.0 - LP if P % 2 == 0 else 1.0 + (-1.0) ** m * LP

    def PW(self, X, axis, transpose=False):
        """P_W = W_I T^-1 W_I^T M_II (transpose: P_W^T = M_II W_I T^-1 W_I^T) applied along `axis`, which runs over the
        INTERIOR nodes of this direction."""
        MII = self.M[1:-1]
        X = np.moveaxis(X, axis, 0)
        shp = (-1,) + (1,) * (X.ndim - 1)
        C = self.WI.T @ (X if transpose else MII.reshape(shp) * X)
        out = self.WI @ np.tensordot(self.Tinv, C, axes=(1, 0))
        if transpose:
            out = MII.reshape(shp) * out
        return np.moveaxis(out, 0, axis)


class NSPrecondMirror:
    def __init__(self, ns: so.NSOracle):
        self.ns = ns
        P, nex, ney = ns._P, ns._N_ex, ns._N_ey
        self.dx_, self.dy_ = Dir1D(P, nex, ns._dx), Dir1D(P, ney, ns._dy)
        self.NX, self.NY, self.N = self.dx_.n, self.dy_.n, ns.N
        self.pin = ns._pin
        self.bnd = ns._mask_bound.copy()
        self.ring = self.bnd.copy()
        self.ring[self.pin] = False                     # the pin node keeps its identity row (JVP form, NS:157-158)
        self.inner = ~self.bnd
        self.inner[self.pin] = False                    # interior pressure nodes without the pin
        self.Mp = ns._M.copy()
        self.Mp[self.pin] = 1.0
        self.Kdiag = ns._K.diagonal()
        self.cheb_lo, self.cheb_hi, self.cheb_steps = ring_chebyshev(ns._K, self.ring)
        self.lc = np.outer(self.dx_.lfac, self.dy_.lfac).ravel()
        self.mc = self.Mp * self.lc
        # the Jacobian is singular (left null vector (l_u, l_v, l_c)) iff this candidate vanishes on the boundary and at the pin
        self.singular = bool(np.all(np.abs(self.lc[self.bnd]) < 1e-12) and abs(self.lc[self.pin]) < 1e-12)

    # ---- building blocks ------------------------------------------------------------------------------------------------
    def _2d(self, v):
        return v.reshape(self.NX, self.NY)

    def fdm(self, r, Qx, lx, Qy, ly):
        T = Qx.T @ self._2d(r) @ Qy
        den = lx[:, None] + ly[None, :]
        T = np.where(den > 0, T / np.where(den > 0, den, 1.0), 0.0)
        return (Qx @ T @ Qy.T).ravel()

    def vel_fdm(self, r):
        """z = K^-1 r on the interior nodes, z = r on the boundary (the Stokes velocity block with its identity rows)."""
        z = self.fdm(r, self.dx_.Qd, self.dx_.ld, self.dy_.Qd, self.dy_.ld)
        z[self.bnd] = r[self.bnd]
        return z

    def neu_fdm(self, r):
        return self.fdm(r, self.dx_.Qn, self.dx_.ln, self.dy_.Qn, self.dy_.ln)

    def coarse_fdm(self, r):
        return self.fdm(r, self.dx_.Qc, self.dx_.lc_, self.dy_.Qc, self.dy_.lc_)

    def project(self, v, transpose):
        """Pi v = v - Q v Q^T (transpose: Pi^T v = v - Q^T v Q), Q = I - P_W, on the interior grid; zero outside."""
        X = self._2d(v)[1:-1, 1:-1]
        U = X - self.dx_.PW(X, 0, transpose)
        U = U - self.dy_.PW(U, 1, transpose)
        out = np.zeros((self.NX, self.NY))
        out[1:-1, 1:-1] = X - U
        return out.ravel()

    def coarse(self, y):
        """z1 = Pi Shat^+ Pi^T y on the interior pressure nodes (pin excluded), zero elsewhere."""
        Y = np.where(self.inner, y, 0.0)
        X = self.coarse_fdm(self.project(Y, True))
        return np.where(self.inner, self.project(X, False), 0.0)

    def neumann_rows(self, q):
        """(K q) on the ring nodes, zero elsewhere."""
        out = np.zeros(self.N)
        out[self.ring] = (self.ns._K @ q)[self.ring]
        return out

    def ring_solve(self, rhs, steps=None):
        """K_BB^-1 rhs on the ring by a fixed Chebyshev polynomial of the diagonally scaled block (spectrum bounds from ring_chebyshev)."""
        ring = self.ring
        sd = np.zeros(self.N)
        sd[ring] = 1.0 / np.sqrt(self.Kdiag[ring])
        steps = self.cheb_steps if steps is None else steps
        theta, delta = 0.5 * (self.cheb_hi + self.cheb_lo), 0.5 * (self.cheb_hi - self.cheb_lo)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = sd * rhs
        y = np.zeros(self.N)
        d = r / theta
        for _ in range(steps):
            y = y + d
            r = r - sd * self.neumann_rows(sd * d)
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + (2.0 * rho_new / delta) * r
            rho = rho_new
        return sd * y

    def add_ring(self, z, y):
        """z given on the interior (+ pin): z_B = K_BB^-1 (y_B - K_BI z_I), z_pin = y_pin."""
        z = np.where(self.inner, z, 0.0)
        z[self.pin] = y[self.pin]
        zb = self.ring_solve(np.where(self.ring, y, 0.0) - self.neumann_rows(z))
        return z + zb

    def div_inner(self, a, b):
        """continuity rows of the Jacobian: G_x a + G_y b on the interior pressure nodes, zero on ring and pin"""
        return np.where(self.inner, self.ns._G_x @ a + self.ns._G_y @ b, 0.0)

    def schur_stokes(self, q):
        """S_0 q = D q - C A_0^-1 G q with the Stokes velocity block (what the device inverts exactly)."""
        ns = self.ns
        gx = np.where(self.bnd, 0.0, ns._G_x @ q)
        gy = np.where(self.bnd, 0.0, ns._G_y @ q)
        out = self.neumann_rows(q) - self.div_inner(self.vel_fdm(gx), self.vel_fdm(gy))
        out[self.pin] = q[self.pin]
        return out

    def pcd(self, r):
        """M^-1 F_p K_N^+ r on the interior pressure nodes."""
        ns = self.ns
        q = self.neu_fdm(np.where(self.inner, r, 0.0))
        Fq = ns._K @ q + ns._Re * (ns._u * (ns._G_x @ q) + ns._v * (ns._G_y @ q))
        return np.where(self.inner, Fq / self.Mp, 0.0)

    def member_correction(self, z, y):
        if not self.singular:
            return z
        return z - self.lc * ((self.mc @ z - self.lc @ y) / (self.mc @ self.lc))

    # ---- Schur-block preconditioners ---------------------------------------------------------------------------------------
    def schur_inv(self, y, level, correct=True):
        if level == 2:
            return y / self.Mp
        if level == 3:
            return self.add_ring(y / self.Mp, y)
        z1 = self.add_ring(self.coarse(y), y)
        r1 = y - self.schur_stokes(z1)
        z2 = self.add_ring(self.pcd(r1), r1)
        z = z1 + z2
        return self.member_correction(z, y) if correct else z

    def apply(self, r, level=4):
        """Block lower-triangular preconditioner applied to r = (r_u, r_v, r_c)."""
        N = self.N
        zu, zv = self.vel_fdm(r[:N]), self.vel_fdm(r[N:2 * N])
        y = r[2 * N:] - self.div_inner(zu, zv)
        return np.hstack((zu, zv, self.schur_inv(y, level)))


def gmres_right(A, b, Pinv, tol, maxit):
    """Right-preconditioned full GMRES (CGS2); returns x, iterations."""
    n = b.size
    V = np.zeros((maxit + 1, n))
    Z = np.zeros((maxit, n))
    H = np.zeros((maxit + 1, maxit))
    beta = np.linalg.norm(b)
    V[0] = b / beta
    g = np.zeros(maxit + 1)
    g[0] = beta
    cs, sn = np.zeros(maxit), np.zeros(maxit)
    k = -1
    for k in range(maxit):
        Z[k] = Pinv(V[k])
        w = A(Z[k])
        for _ in range(2):
            h = V[:k + 1] @ w
            w -= h @ V[:k + 1]
            H[:k + 1, k] += h
        H[k + 1, k] = np.linalg.norm(w)
        V[k + 1] = w / H[k + 1, k]
        for i in range(k):
            t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
            H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
            H[i, k] = t
        d = np.hypot(H[k, k], H[k + 1, k])
        cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
        H[k, k], H[k + 1, k] = d, 0.0
        g[k + 1] = -sn[k] * g[k]
        g[k] = cs[k] * g[k]
        if abs(g[k + 1]) <= tol:
            break
    y = np.linalg.solve(np.triu(H[:k + 1, :k + 1]), g[:k + 1])
    return y @ Z[:k + 1], k + 1
