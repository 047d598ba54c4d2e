"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy/scipy) for the SEM hot path.

A plain restatement of the reference's algorithm for the path named in BASELINE.json (``north_star``):
GLL tables, local->global index map, gather-scatter assembly, the global M/K/G_x/G_y operators, the
convection--diffusion (CD) and Navier--Stokes (NS) residual / Jacobian-vector products with their boundary rows,
and the Newton / linear solves.  Every function cites the reference ``file:line`` it follows
(paths relative to ``/root/reference``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import this module, and only as the checker / the CPU baseline -- never as the product path.  The product
(``sem_b200``) never imports it and fails loudly when its CUDA library is missing.

PINNING.  The reference ships no tests and no golden vectors.  This oracle is pinned instead
against outputs of the reference itself, run in the build container through ``oracle/ref_shim.py``; the vectors
are committed under ``tests/golden/*.npz`` together with the generating script ``tests/golden/make_golden.py``,
and ``tests/test_oracle.py`` re-checks the oracle against them on every run (operator applies to <= 1e-13 relative,
converged fields to <= 1e-9 relative), plus the analytic GLL known answers and the README Helmholtz example
(``Solvers/README.md:49-96``).  The coupled Boussinesq couplers (OpenMDAO, not installable here) are "parity
unpinned" at the coupler level: only the coupled fixed point is checked (block Gauss-Seidel over the reference's
own solver calls, see ``make_golden.py``).

Matrices are built with Kronecker products of 1-D assembled matrices.  That is value-identical to the reference's
dense ``einsum`` + COO duplicate summation (``SEM.py:170-223``) -- checked in ``tests/test_oracle.py`` -- but needs
O(nnz) memory instead of O(N_e^2 (P+1)^4), so the oracle also reaches mesh sizes the reference cannot build.
The 3-index convection tensors (``SEM.py:226-245``) are never formed: ``u @ C_x == diag(u) G_x`` and
``C_x @ T == diag(G_x T)`` hold exactly because F_s and C_s carry Kronecker deltas (``GLL.py:84-102``).
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla


# ----------------------------------------------------------------------------------------------------------------
# 1-D GLL tables                                                                                     (Solvers/GLL.py)
# ----------------------------------------------------------------------------------------------------------------
def gll(P):
    """Nodes, weights and Legendre Vandermonde matrix L_k(xi_i) of the (P+1)-point GLL rule.

    Follows ``GLL.py:7-33``: Chebyshev--Lobatto first guess ``-cos(pi i / P)``, Newton updates
    ``-(x L_P - L_{P-1}) / ((P+1) L_P)`` until the largest update is <= machine eps, weights ``2 / (P (P+1) L_P^2)``.
    """
    x = -np.cos(np.pi * np.arange(P + 1) / P)
    V = np.zeros((P + 1, P + 1))
    eps = np.finfo(np.float64).eps
    while True:
        V[:, 0] = 1.0
        V[:, 1] = x
        for k in range(2, P + 1):
            V[:, k] = ((2 * k - 1) * x * V[:, k - 1] - (k - 1) * V[:, k - 2]) / k
        step = -(x * V[:, P] - V[:, P - 1]) / ((P + 1) * V[:, P])
        x = x + step
        if np.max(np.abs(step)) <= eps:
            break
    w = 2.0 / (P * (P + 1) * V[:, P] ** 2)
    return x, w, V


def diff_matrix(P):
    """D[i, j] = l_j'(xi_i)  (``GLL.py:45-59``): L_P(xi_i) / L_P(xi_j) / (xi_i - xi_j), corners -+P(P+1)/4."""
    x, _, V = gll(P)
    LP = V[:, -1]
    D = np.zeros((P + 1, P + 1))
    for i in range(P + 1):
        for j in range(P + 1):
            if i != j:
                D[i, j] = LP[i] / LP[j] * 1 / (x[i] - x[j])
    D[0, 0] = -P * (P + 1) / 4
    D[-1, -1] = P * (P + 1) / 4
    return D


def mass_1d(P):
    """diag(w)  (``GLL.py:36-42``)."""
    return np.diag(gll(P)[1])


def grad_1d(P):
    """G_s = diag(w) D  (``GLL.py:62-70``)."""
    return np.einsum('i,ij->ij', gll(P)[1], diff_matrix(P))


def stiff_1d(P):
    """K_s = D^T diag(w) D  (``GLL.py:73-81``), summed over the quadrature index in that einsum's order."""
    D = diff_matrix(P)
    return np.einsum('k,ki,kj->ij', gll(P)[1], D, D)


def eval_matrix(P, xi):
    """S[i, j] = l_j(xi[i])  (``GLL.py:105-116``), product form of the Lagrange basis."""
    x = gll(P)[0]
    xi = np.asarray(xi, dtype=np.float64)
    S = np.ones((xi.size, P + 1))
    for j in range(P + 1):
        for k in range(P + 1):
            if k != j:
                S[:, j] *= (xi - x[k]) / (x[j] - x[k])
    return S


# ----------------------------------------------------------------------------------------------------------------
# mesh, index map, gather-scatter                                                                   (Solvers/SEM.py)
# ----------------------------------------------------------------------------------------------------------------
def nodes_1d(P, Ne, h):
    """Global 1-D node coordinates (``SEM.py:39-60``): x = h/2 (xi+1) + h m, first node of elements m>0 dropped."""
    xi = gll(P)[0]
    xe = np.vstack([h / 2 * (xi + 1) + h * m for m in range(Ne)])
    return np.insert(np.ravel(xe[:, 1:]), 0, 0)


def global_nodes(P, N_ex, N_ey, dx, dy):
    """2 x N coordinates, x slow / y fast  (``SEM.py:82-94``)."""
    x1 = nodes_1d(P, N_ex, dx)
    y1 = nodes_1d(P, N_ey, dy)
    return np.reshape(np.array(np.meshgrid(x1, y1, indexing='ij')), (2, x1.size * y1.size))


def element_nodes(P, N_ex, N_ey, dx, dy):
    """2 x N_ex x N_ey x (P+1) x (P+1) element coordinates  (``SEM.py:63-79``)."""
    xi = gll(P)[0]
    xe = np.vstack([dx / 2 * (xi + 1) + dx * m for m in range(N_ex)])
    ye = np.vstack([dy / 2 * (xi + 1) + dy * n for n in range(N_ey)])
    pts = np.zeros((2, N_ex, N_ey, P + 1, P + 1))
    pts[0] = xe[:, None, :, None]
    pts[1] = ye[None, :, None, :]
    return pts


def global_index(P, N_ex, N_ey, m, n, i, j):
    """``n P + j + (N_ey P + 1)(m P + i)``  (``SEM.py:97-110``), raising on out-of-range like the reference."""
    if np.any(m >= N_ex) or np.any(n >= N_ey) or np.any(i > P) or np.any(j > P):
        raise ValueError('Indices out of range')
    return n * P + j + (N_ey * P + 1) * (m * P + i)


def element_index_array(P, N_ex, N_ey):
    """int64 array [m, n, i, j] -> global index."""
    m, n, i, j = np.meshgrid(np.arange(N_ex), np.arange(N_ey), np.arange(P + 1), np.arange(P + 1), indexing='ij')
    return global_index(P, N_ex, N_ey, m, n, i, j)


def assemble_vector(A_e):
    """Gather-scatter of a 4-index element array into a global vector  (``SEM.py:126-131``).

    Duplicates are summed in row-major element order (the order ``np.nonzero`` emits and COO->dense accumulates).
    """
    N_ex, N_ey, n1, _ = A_e.shape
    P = n1 - 1
    out = np.zeros((P * N_ex + 1) * (P * N_ey + 1))
    np.add.at(out, element_index_array(P, N_ex, N_ey).ravel(), np.asarray(A_e, dtype=np.float64).ravel())
    return out


def scatter(u, P, N_ex, N_ey):
    """Global vector -> element array copy  (``SEM.py:149-167``)."""
    if u.shape[0] != (P * N_ex + 1) * (P * N_ey + 1):
        raise ValueError('Not a valid combination of global coefficients vector, P, N_ex, and N_ey')
    return np.asarray(u)[element_index_array(P, N_ex, N_ey)]


def _assembled_1d(Ae, Ne):
    """Sum Ne copies of the (P+1)x(P+1) element matrix along the diagonal with one shared node (1-D assembly)."""
    n1 = Ae.shape[0]
    P = n1 - 1
    rows, cols, vals = [], [], []
    ii, jj = np.meshgrid(np.arange(n1), np.arange(n1), indexing='ij')
    for m in range(Ne):
        rows.append((ii + m * P).ravel())
        cols.append((jj + m * P).ravel())
        vals.append(Ae.ravel())
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    keep = vals != 0  # the reference drops exact zeros before assembly (np.nonzero, SEM.py:127,133)
    return sps.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(Ne * P + 1,) * 2).tocsr()


def global_operators(P, N_ex, N_ey, dx, dy):
    """M (diagonal, returned as vector), K, G_x, G_y as the reference assembles them  (``SEM.py:170-223``).

    ``M_e = (dx/2 M_s) x (dy/2 M_s)``; ``K_e = (2/dx K_s) x (dy/2 M_s) + (dx/2 M_s) x (2/dy K_s)``;
    ``G_x_e = G_s x (dy/2 M_s)``; ``G_y_e = (dx/2 M_s) x G_s``  (x = Kronecker, x index slow).
    """
    Ms, Ks, Gs = mass_1d(P), stiff_1d(P), grad_1d(P)
    Mx = _assembled_1d(dx / 2 * Ms, N_ex)
    My = _assembled_1d(dy / 2 * Ms, N_ey)
    Kx = _assembled_1d(2 / dx * Ks, N_ex)
    Ky = _assembled_1d(2 / dy * Ks, N_ey)
    Gx1 = _assembled_1d(Gs, N_ex)
    Gy1 = _assembled_1d(Gs, N_ey)
    M = sps.kron(Mx, My, format='csr').diagonal()
    K = (sps.kron(Kx, My, format='csr') + sps.kron(Mx, Ky, format='csr')).tocsr()
    G_x = sps.kron(Gx1, My, format='csr')
    G_y = sps.kron(Mx, Gy1, format='csr')
    return M, K, G_x, G_y


class KronOps:
    """The reference's global operators on meshes it cannot assemble (BASELINE config 5: 67 M nodes), through the Kronecker
    identities of ``global_operators`` -- ``K = Kx (x) My + Mx (x) Ky``, ``G_x = Gx1 (x) My``, ``G_y = Mx (x) Gy1`` (value
    identical to SEM.py:170-223, checked in tests/test_oracle.py) -- applied to a field stored as the 2-D array X[ix, iy]:
    ``kron(A, B) vec(X) = vec(A X B^T)``.  Only sparse 1-D matrices and dense 2-D fields: seconds at 67 M nodes."""

    def __init__(self, P, N_ex, N_ey, dx, dy):
        Ms, Ks, Gs = mass_1d(P), stiff_1d(P), grad_1d(P)
        self.Mx = _assembled_1d(dx / 2 * Ms, N_ex).diagonal()
        self.My = _assembled_1d(dy / 2 * Ms, N_ey).diagonal()
        self.Kx = _assembled_1d(2 / dx * Ks, N_ex)
        self.Ky = _assembled_1d(2 / dy * Ks, N_ey)
        self.Gx1 = _assembled_1d(Gs, N_ex)
        self.Gy1 = _assembled_1d(Gs, N_ey)
        self.shape = (N_ex * P + 1, N_ey * P + 1)

    def K(self, X):
        return (self.Kx @ X) * self.My[None, :] + self.Mx[:, None] * (self.Ky @ X.T).T

    def Gx(self, X):
        return (self.Gx1 @ X) * self.My[None, :]

    def Gy(self, X):
        return self.Mx[:, None] * (self.Gy1 @ X.T).T

    def M(self, X):
        return self.Mx[:, None] * X * self.My[None, :]

    def cd_jvp(self, Pe, dT, u, v, dirichlet_wesn):
        """CD:104-121 without the du/dv terms: Sys dT with Dirichlet rows = dT."""
        Y = self.K(dT) + Pe * (u * self.Gx(dT) + v * self.Gy(dT))
        W, E, S, N = dirichlet_wesn
        if W: Y[0, :] = dT[0, :]
        if E: Y[-1, :] = dT[-1, :]
        if S: Y[:, 0] = dT[:, 0]
        if N: Y[:, -1] = dT[:, -1]
        return Y

    def ns_jvp(self, Re, u, v, du, dv, dp):
        """NS:138-160 (dT = None) about the state (u, v): Jacobian diagonals Re G u etc. (NS:123-136), velocity Dirichlet rows,
        pressure-Neumann rows K[mask,:] dp, then the pin at node int(N/2)."""
        sys_ = lambda X: self.K(X) + Re * (u * self.Gx(X) + v * self.Gy(X))
        ru = sys_(du) + Re * self.Gx(u) * du + Re * self.Gy(u) * dv + self.Gx(dp)
        rv = Re * self.Gx(v) * du + sys_(dv) + Re * self.Gy(v) * dv + self.Gy(dp)
        rc = self.Gx(du) + self.Gy(dv)
        Kp = self.K(dp)
        for sl_ in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
            ru[sl_], rv[sl_], rc[sl_] = du[sl_], dv[sl_], Kp[sl_]
        pin = (self.shape[0] * self.shape[1]) // 2
        rc[divmod(pin, self.shape[1])] = dp[divmod(pin, self.shape[1])]
        return ru, rv, rc


def x2xi(x, h):
    """Physical coordinate -> (element, xi)  (``SEM.py:23-36``): left-element convention at shared nodes."""
    frac, e = np.modf(np.asarray(x, dtype=np.float64) / h)
    xi = 2 * frac - 1
    left = np.isclose(xi, -1) * (e > 0)
    e[left] -= 1
    xi[left] = 1
    return e.astype(int), xi


def interpolate(f, P, N_ex, N_ey, dx, dy, points_plot):
    """Evaluate the SEM interpolant of the global vector f on an ij-meshgrid  (``SEM.py:248-273``)."""
    f_e = scatter(f, P, N_ex, N_ey)
    xs = points_plot[0][:, 0]
    ys = points_plot[1][0, :]
    m_p, xi_p = x2xi(xs, dx)
    n_p, eta_p = x2xi(ys, dy)
    Sx = eval_matrix(P, xi_p)  # [a, k]
    Sy = eval_matrix(P, eta_p)  # [b, l]
    val = np.zeros((xs.size, ys.size))
    for m in range(N_ex):
        for n in range(N_ey):
            ia = np.nonzero(m_p == m)[0]
            ib = np.nonzero(n_p == n)[0]
            if ia.size and ib.size:
                val[np.ix_(ia, ib)] = np.einsum('kl,ik,jl->ij', f_e[m, n], Sx[ia], Sy[ib])
    return val


# ----------------------------------------------------------------------------------------------------------------
# solvers                                                  (Solvers/ConvectionDiffusion_Solver.py, NavierStokes_Solver.py)
# ----------------------------------------------------------------------------------------------------------------
class CDOracle:
    """Steady convection--diffusion, ``Pe [u,v].grad T = lap T``  (``ConvectionDiffusion_Solver.py:9-203``)."""

    def __init__(self, L_x, L_y, Pe, P, N_ex, N_ey, T_W=None, T_E=None, T_S=None, T_N=None, mtol=1e-7):
        self._Pe, self._mtol = Pe, mtol
        self._L_x, self._L_y, self._P, self._N_ex, self._N_ey = L_x, L_y, P, N_ex, N_ey
        self._dx, self._dy = L_x / N_ex, L_y / N_ey
        self.points = global_nodes(P, N_ex, N_ey, self._dx, self._dy)
        self.N = (N_ex * P + 1) * (N_ey * P + 1)
        self._M, self._K, self._G_x, self._G_y = global_operators(P, N_ex, N_ey, self._dx, self._dy)
        # Dirichlet rows: W, E, S, N in that order, later assignments win at corners  (CD:62-71)
        d = np.full(self.N, np.nan)
        if T_W is not None:
            d[np.isclose(self.points[0], 0)] = T_W
        if T_E is not None:
            d[np.isclose(self.points[0], L_x)] = T_E
        if T_S is not None:
            d[np.isclose(self.points[1], 0)] = T_S
        if T_N is not None:
            d[np.isclose(self.points[1], L_y)] = T_N
        self._dirichlet = d
        self._mask_dir = ~np.isnan(d)
        self._Sys = self._gxT = self._gyT = None

    def _get_residuals(self, T, u, v):
        """``Sys = K + Pe (diag(u) G_x + diag(v) G_y)``; ``res = Sys T``; Dirichlet rows ``T - T_dir``  (CD:73-92)."""
        Conv = self._Pe * (sps.diags(u) @ self._G_x + sps.diags(v) @ self._G_y)
        self._Sys = (Conv + self._K).tocsr()
        res = self._Sys @ T
        res[self._mask_dir] = T[self._mask_dir] - self._dirichlet[self._mask_dir]
        return res

    def _calc_jacobians(self, T):
        """``Jac_T_u = Pe diag(G_x T)``, ``Jac_T_v = Pe diag(G_y T)``  (CD:94-102)."""
        self._gxT = self._Pe * (self._G_x @ T)
        self._gyT = self._Pe * (self._G_y @ T)

    def _get_dresiduals(self, dT, du=None, dv=None):
        """``Sys dT (+ Jac_T_u du + Jac_T_v dv)``; Dirichlet rows ``dT``  (CD:104-121)."""
        dres = self._Sys @ dT
        if du is not None:
            dres += self._gxT * du
        if dv is not None:
            dres += self._gyT * dv
        dres[self._mask_dir] = dT[self._mask_dir]
        return dres

    def jacobian_matrix(self):
        """The matrix whose action is ``_get_dresiduals(dT)``: Sys with Dirichlet rows replaced by identity."""
        keep = sps.diags((~self._mask_dir).astype(float))
        return (keep @ self._Sys + sps.diags(self._mask_dir.astype(float))).tocsc()

    def _get_update(self, dres, dT0=None):
        """Solve ``J dT = dres``  (CD:123-156).  The reference iterates LGMRES to ``mtol sqrt(N)``; the oracle
        solves the same system directly (sparse LU), i.e. the limit the reference converges to."""
        return spla.splu(self.jacobian_matrix()).solve(dres)

    def _get_solution(self, u, v, T0=None):
        """One Newton step from T0 (problem is linear)  (CD:158-170)."""
        T = T0 if T0 is not None else np.zeros(self.N)
        res = self._get_residuals(T, u, v)
        return T + self._get_update(-res)

    def _get_solution_lgmres(self, u, v, T0=None):
        """The reference's OWN algorithm for the CD solve, restated (CD:123-170): matrix-free operator through
        ``_get_dresiduals``, SciPy LGMRES without preconditioner, ``inner_m = int(0.3 N)``, ``atol = mtol sqrt(N)``.  Host timing
        baseline of bench.py (`kind: port`).  Returns (T, operator evaluations)."""
        T = np.zeros(self.N) if T0 is None else T0
        res = self._get_residuals(T, u, v)
        count = [0]

        def mv(x):
            count[0] += 1
            return self._get_dresiduals(x)

        A = spla.LinearOperator((self.N, self.N), mv, dtype=float)
        dT, info = spla.lgmres(A, -res, atol=self._mtol * np.sqrt(self.N), rtol=0, inner_m=int(self.N * 0.3))
        if info != 0:
            raise RuntimeError(f'ConvectionDiffusion LGMRES: Failed to converge in {info} iterations')
        return T + dT, count[0]

    def _get_vector(self, f_func):
        return f_func(self.points[0], self.points[1])

    def _get_interpol(self, f, points_plot):
        return interpolate(f, self._P, self._N_ex, self._N_ey, self._dx, self._dy, points_plot)


class NSOracle:
    """Steady incompressible Navier--Stokes with Boussinesq buoyancy  (``NavierStokes_Solver.py:10-303``)."""

    def __init__(self, L_x, L_y, Re, Gr, P, N_ex, N_ey, v_W=0, v_E=0, u_S=0, u_N=0, mtol=1e-7, mtol_newton=1e-5):
        if Re == 0 and Gr != 0:
            raise ValueError('Cannot have Re == 0 and Gr != 0')
        self._Re, self._Gr = Re, Gr
        self._Gr_over_Re = Gr / Re if Re != 0 else 0.
        self._mtol, self._mtol_newton = mtol, mtol_newton
        self._L_x, self._L_y, self._P, self._N_ex, self._N_ey = L_x, L_y, P, N_ex, N_ey
        self._dx, self._dy = L_x / N_ex, L_y / N_ey
        self.points = global_nodes(P, N_ex, N_ey, self._dx, self._dy)
        self.N = (N_ex * P + 1) * (N_ey * P + 1)
        self._M, self._K, self._G_x, self._G_y = global_operators(P, N_ex, N_ey, self._dx, self._dy)
        # W, E, S, N in that order; the later assignment wins at corners (lid value at the top corners)  (NS:78-91)
        du_ = np.full(self.N, np.nan)
        dv_ = np.full(self.N, np.nan)
        W, E = np.isclose(self.points[0], 0), np.isclose(self.points[0], L_x)
        S, Nn = np.isclose(self.points[1], 0), np.isclose(self.points[1], L_y)
        dv_[W] = v_W
        du_[W] = 0
        dv_[E] = v_E
        du_[E] = 0
        du_[S] = u_S
        dv_[S] = 0
        du_[Nn] = u_N
        dv_[Nn] = 0
        self._dirichlet_u, self._dirichlet_v = du_, dv_
        self._mask_bound = ~np.isnan(du_)
        self._pin = int(self.N / 2)  # reference pressure node  (NS:89)
        self._K_bound = self._K[self._mask_bound, :]
        self._u = self._v = None
        self._k = 0

    def _get_residuals(self, u, v, p, T):
        """Momentum and continuity residuals with their boundary rows  (NS:93-121)."""
        self._u, self._v = np.array(u), np.array(v)
        Sys_u = self._K @ u + self._Re * (u * (self._G_x @ u) + v * (self._G_y @ u))
        Sys_v = self._K @ v + self._Re * (u * (self._G_x @ v) + v * (self._G_y @ v))
        res_u = Sys_u + self._G_x @ p
        res_v = Sys_v + self._G_y @ p - self._Gr_over_Re * self._M * T
        res_c = self._G_x @ u + self._G_y @ v
        mb = self._mask_bound
        res_u[mb] = u[mb] - self._dirichlet_u[mb]
        res_v[mb] = v[mb] - self._dirichlet_v[mb]
        res_c[self._pin] = p[self._pin] - 0.0          # pin first ...
        res_c[mb] = self._K_bound @ p                  # ... then the pressure-Neumann rows  (NS:116-119)
        return res_u, res_v, res_c

    def _calc_jacobians(self, u, v):
        """Diagonals ``Re G_x u``, ``Re G_y u``, ``Re G_x v``, ``Re G_y v`` about the last residual point (NS:123-136).

        Like the reference, ``Sys`` (hence the advecting velocity) is the one cached by the last ``_get_residuals``.
        """
        self._gxu = self._Re * (self._G_x @ u)
        self._gyu = self._Re * (self._G_y @ u)
        self._gxv = self._Re * (self._G_x @ v)
        self._gyv = self._Re * (self._G_y @ v)

    def _sys(self, x):
        return self._K @ x + self._Re * (self._u * (self._G_x @ x) + self._v * (self._G_y @ x))

    def _get_dresiduals(self, du, dv, dp, dT=None):
        """3-field JVP with boundary rows: Neumann rows first, then the pin  (NS:138-160)."""
        dres_u = self._sys(du) + self._gxu * du + self._gyu * dv + self._G_x @ dp
        dres_v = self._gxv * du + self._sys(dv) + self._gyv * dv + self._G_y @ dp
        dres_c = self._G_x @ du + self._G_y @ dv
        if dT is not None:
            dres_v += -self._Gr_over_Re * self._M * dT
        mb = self._mask_bound
        dres_u[mb] = du[mb]
        dres_v[mb] = dv[mb]
        dres_c[mb] = self._K_bound @ dp
        dres_c[self._pin] = dp[self._pin]
        return dres_u, dres_v, dres_c

    def jacobian_matrix(self):
        """3N x 3N matrix whose action is ``_get_dresiduals(du, dv, dp)``."""
        N = self.N
        I = sps.identity(N, format='csr')
        Sys = self._K + self._Re * (sps.diags(self._u) @ self._G_x + sps.diags(self._v) @ self._G_y)
        inner = sps.diags((~self._mask_bound).astype(float))
        bnd = sps.diags(self._mask_bound.astype(float))
        Juu = inner @ (Sys + sps.diags(self._gxu)) + bnd
        Juv = inner @ sps.diags(self._gyu)
        Jvu = inner @ sps.diags(self._gxv)
        Jvv = inner @ (Sys + sps.diags(self._gyv)) + bnd
        Gxp = inner @ self._G_x
        Gyp = inner @ self._G_y
        cmask = np.ones(N)
        cmask[self._mask_bound] = 0
        cmask[self._pin] = 0
        cin = sps.diags(cmask)
        Cu = cin @ self._G_x
        Cv = cin @ self._G_y
        bp = self._mask_bound.astype(float)
        bp[self._pin] = 0  # pin overrides the Neumann row in the JVP
        pin = np.zeros(N)
        pin[self._pin] = 1
        Cp = sps.diags(bp) @ self._K + sps.diags(pin)
        return sps.bmat([[Juu, Juv, Gxp], [Jvu, Jvv, Gyp], [Cu, Cv, Cp]], format='csc')

    def _left_null(self, J):
        """Left null vector of the Jacobian, or None when J is regular.

        With equal-order velocity/pressure spaces and the boundary rows of NS:155-158 the 3N x 3N Jacobian has, on
        most meshes, a one-dimensional null space (a spurious pressure mode); its left null vector ``l`` vanishes on
        the interior momentum rows and does not depend on the linearisation point (cached).  Found by one sparse LU
        of ``J^T`` with one (redundant) equation replaced by the normalisation ``l[r*] = 1``.
        """
        if getattr(self, "_l_cache", None) is not None:
            return self._l_cache if self._l_cache is not False else None
        n = J.shape[0]
        r_star = 2 * self.N + self._interior_probe()
        JT = J.T.tolil()
        JT[r_star, :] = 0
        JT[r_star, r_star] = 1.0
        rhs = np.zeros(n)
        rhs[r_star] = 1.0
        l = spla.splu(JT.tocsc()).solve(rhs)
        l /= np.linalg.norm(l)
        if np.linalg.norm(J.T @ l) > 1e-9 * spla.norm(J, np.inf):
            self._l_cache = False
            return None
        self._l_cache = l
        return l

    def _interior_probe(self):
        """Index of the mesh-interior node (ix, iy) = (1, 1) (never the pin, never on the boundary)."""
        NY = self._N_ey * self._P + 1
        k = NY + 1
        if k == self._pin:
            k += 1
        return k

    def _get_update(self, dres_u, dres_v, dres_cont, du0=None, dv0=None, dp0=None):
        """Solve the linearised system  (NS:162-236).

        The reference factorises the velocity block and runs right-preconditioned LGMRES on the pressure Schur
        complement ``S`` from ``dp0`` (or 0) with the diagonal-mass preconditioner of NS:208-212 (pin row passed
        through).  ``S`` is singular but the system is consistent, and that iteration converges to the one solution
        with ``M_p (dp - dp0)`` in ``range(S)``, i.e. ``l_c . (M_p (dp - dp0)) = 0`` for the left null vector
        ``l = (l_u, l_v, l_c)`` of the Jacobian.  The oracle computes that same solution directly: the continuity
        equation of one interior node is redundant (``l`` is non-zero there); it is replaced by ``p[node] = 0`` to get
        a particular solution and the right null vector from one sparse LU, and the solution is shifted along the
        null vector until ``m . (x - x0) = 0`` with ``m = (0, 0, M_p l_c)``.  (Checked against the
        reference's own converged output in tests/golden/ns.npz.)  When J is regular it is solved as is.
        """
        N = self.N
        J = self.jacobian_matrix()
        b = np.hstack((dres_u, dres_v, dres_cont))
        l = self._left_null(J)
        if l is None:
            return tuple(np.split(spla.splu(J).solve(b), 3))
        Mp = self._M.copy()
        Mp[self._pin] = 1.0
        m = np.hstack((np.zeros(2 * N), Mp * l[2 * N:]))
        x0 = np.hstack([np.zeros(N) if a is None else a for a in (du0, dv0, dp0)])
        # regularise with a sparse row instead of the dense constraint row: p[probe] = 0 gives a particular solution
        # x1 and the same LU yields the right null vector q (J_r q = e_r*); then shift along q onto the constraint.
        r_star = 2 * N + self._interior_probe()
        keep = np.ones(3 * N)
        keep[r_star] = 0.0
        unit = np.zeros(3 * N)
        unit[r_star] = 1.0
        lu = spla.splu((sps.diags(keep) @ J + sps.diags(unit)).tocsc())
        x1 = lu.solve(b * keep)
        q = lu.solve(unit)
        alpha = (m @ (x1 - x0)) / (m @ q)
        return tuple(np.split(x1 - alpha * q, 3))

    def _get_update_schur(self, dres_u, dres_v, dres_cont, du0=None, dv0=None, dp0=None):
        """The reference's OWN algorithm for the linearised system, restated line by line (NS:162-236): SuperLU of the
        boundary-modified 2N x 2N velocity Jacobian (NS:178-184), right-hand side and matrix-free operator of the pressure Schur
        complement through ``_get_dresiduals`` (NS:196-205), diagonal-mass preconditioner with the pin row passed through
        (NS:208-212), SciPy LGMRES with ``inner_m = int(0.3 N)`` and ``atol = mtol sqrt(N)`` (NS:222-224), back-substitution
        of the velocities (NS:233-234).  Used as the host timing baseline of bench.py (`kind: port`) and to cross-check
        ``_get_update``; ``self.schur_matvecs`` counts the Schur operator evaluations."""
        N = self.N
        Sys = self._K + self._Re * (sps.diags(self._u) @ self._G_x + sps.diags(self._v) @ self._G_y)
        Jac = sps.bmat([[Sys + sps.diags(self._gxu), sps.diags(self._gyu)],
                        [sps.diags(self._gxv), Sys + sps.diags(self._gyv)]], format='lil')
        mask = np.hstack((self._mask_bound,) * 2)
        Jac[mask, :] = 0
        Jac[mask, mask] = 1
        lu = spla.splu(Jac.tocsc())
        zero = np.zeros(N)

        def solve_velo(ru, rv):
            return np.split(lu.solve(np.hstack((ru, rv))), 2)

        b_schur = dres_cont - self._get_dresiduals(*solve_velo(dres_u, dres_v), zero)[2]
        count = [0]

        def schur_mv(dp):
            count[0] += 1
            f_x, f_y = solve_velo(*self._get_dresiduals(zero, zero, dp)[:2])
            return self._get_dresiduals(-f_x, -f_y, dp)[2]

        def precon_mv(c):
            z = c / self._M
            z[self._pin] = c[self._pin]
            return z

        A = spla.LinearOperator((N, N), schur_mv, dtype=float)
        M = spla.LinearOperator((N, N), precon_mv, dtype=float)
        dp, info = spla.lgmres(A, b_schur, M=M, x0=dp0, atol=self._mtol * np.sqrt(N), rtol=0, inner_m=int(N * 0.3))
        self.schur_matvecs = getattr(self, "schur_matvecs", 0) + count[0]
        if info != 0:
            raise RuntimeError(f'NavierStokes LGMRES: Failed to converge in {info} iterations')
        b_u, b_v = self._get_dresiduals(zero, zero, dp)[:2]
        du, dv = solve_velo(dres_u - b_u, dres_v - b_v)
        return du, dv, dp

    def _get_solution(self, T, u0=None, v0=None, p0=None, max_newton=50, algorithm='direct'):
        """Newton loop, stop on the spectral norm of the 3 x N residual array <= mtol_newton sqrt(3N)  (NS:238-270).
        algorithm: 'direct' (sparse LU of the 3-field Jacobian + member selection, ``_get_update``) or 'reference' (the
        reference's Schur-complement LGMRES, ``_get_update_schur``)."""
        u = u0 if u0 is not None else np.zeros(self.N)
        v = v0 if v0 is not None else np.zeros(self.N)
        p = p0 if p0 is not None else np.zeros(self.N)
        self._k = 0
        while True:
            res_u, res_v, res_c = self._get_residuals(u, v, p, T)
            norm = np.linalg.norm((res_u, res_v, res_c), ord=2)
            if norm <= self._mtol_newton * np.sqrt(self.N * 3) or self._k >= max_newton:
                break
            self._calc_jacobians(u, v)
            solve = self._get_update_schur if algorithm == 'reference' else self._get_update
            du, dv, dp = solve(-res_u, -res_v, -res_c)
            u += du
            v += dv
            p += dp
            self._k += 1
        return u, v, p

    def _get_vector(self, f_func):
        return f_func(self.points[0], self.points[1])

    def _get_interpol(self, f, points_plot):
        return interpolate(f, self._P, self._N_ex, self._N_ey, self._dx, self._dy, points_plot)
