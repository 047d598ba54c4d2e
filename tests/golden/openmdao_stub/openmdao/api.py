"""``openmdao.api`` stand-in: the names used by the reference's OpenMDAO layer --
ImplicitComponent (options.declare / add_input / add_output and the five protocol methods), Group (add_subsystem, connect,
nonlinear_solver, linear_solver, get_val), Problem (model, setup, run_model) and the solver classes NonlinearBlockGS,
NewtonSolver, ArmijoGoldsteinLS, LinearBlockJac, ScipyKrylov with the options the coupler passes
(Boussinesq_SequentialCoupler.py:66-97).  The solution strategies follow the published semantics of those classes as the
coupler configures them (block Gauss-Seidel over solve_nonlinear; Newton with an initial sub-system solve, GMRES(restart) on
apply_linear with one block-Jacobi sweep of solve_linear as preconditioner, or the block-Jacobi sweep alone with an
Armijo-Goldstein backtracking line search)."""
import numpy as np
import scipy.sparse.linalg as spla


class _Options(dict):
    def declare(self, name, default=None, desc=''):
        self.setdefault(name, default)


class _Vec(dict):
    """name -> array view; assignment copies into the existing storage like an OpenMDAO vector."""

    def __setitem__(self, key, value):
        if key in self and isinstance(dict.__getitem__(self, key), np.ndarray):
            dict.__getitem__(self, key)[...] = np.asarray(value, dtype=float).reshape(dict.__getitem__(self, key).shape)
        else:
            dict.__setitem__(self, key, np.array(value, dtype=float))


class ImplicitComponent:
    def __init__(self, **kwargs):
        self.options = _Options()
        self.initialize()
        for k, v in kwargs.items():
            self.options[k] = v
        self._in, self._out = {}, {}

    def initialize(self):
        pass

    def setup(self):
        pass

    def add_input(self, name, val=1.0, desc=''):
        self._in[name] = np.array(val, dtype=float)

    def add_output(self, name, val=1.0, desc=''):
        self._out[name] = np.array(val, dtype=float)


class _Solver:
    def __init__(self, **opts):
        self.options = dict(opts)
        self.precon = None
        self.linesearch = None


class NonlinearBlockGS(_Solver):
    pass


class NewtonSolver(_Solver):
    pass


class ArmijoGoldsteinLS(_Solver):
    pass


class LinearBlockJac(_Solver):
    pass


class ScipyKrylov(_Solver):
    pass


class Group:
    def __init__(self):
        self._subs, self._conn = [], []
        self.nonlinear_solver = None
        self.linear_solver = None

    def add_subsystem(self, name, sub):
        sub.name = name
        self._subs.append(sub)
        return sub

    def connect(self, src, tgt):
        self._conn.append((src, tgt))

    # ---- what Problem.setup / run_model drive ------------------------------------------------------------------------------
    def _setup(self):
        for s in self._subs:
            if isinstance(s, Group):
                s._setup()
        self._comps = [s for s in self._subs if isinstance(s, ImplicitComponent)]
        for c in self._comps:
            c.setup()
        self.out = {c.name: _Vec({k: v.copy() for k, v in c._out.items()}) for c in self._comps}
        self.inp = {c.name: _Vec({k: v.copy() for k, v in c._in.items()}) for c in self._comps}
        self.res = {c.name: _Vec({k: np.zeros_like(v) for k, v in c._out.items()}) for c in self._comps}
        self.src = {}
        for s, t in self._conn:
            self.src[tuple(t.split('.'))] = tuple(s.split('.'))

    def get_val(self, name):
        c, v = name.split('.')
        return self.out[c][v]

    def _transfer(self, comp, out=None, inp=None):
        out = self.out if out is None else out
        inp = self.inp if inp is None else inp
        for k in inp[comp.name]:
            sc, sv = self.src[(comp.name, k)]
            inp[comp.name][k] = out[sc][sv]

    def _apply_nonlinear(self):
        for c in self._comps:
            self._transfer(c)
            c.apply_nonlinear(self.inp[c.name], self.out[c.name], self.res[c.name])
        return np.sqrt(sum(float(v @ v) for c in self._comps for v in self.res[c.name].values()))

    def _layout(self):
        return [(c.name, k, v.size) for c in self._comps for k, v in self.out[c.name].items()]

    def _pack(self, d):
        return np.concatenate([d[c][k] for c, k, _ in self._layout()])

    def _unpack(self, x):
        d, o = {}, 0
        for c, k, n in self._layout():
            d.setdefault(c, _Vec())[k] = x[o:o + n].copy()
            o += n
        return d

    def _jvp(self, dx):
        d_out = self._unpack(dx)
        d_res = {c.name: _Vec({k: np.zeros_like(v) for k, v in self.out[c.name].items()}) for c in self._comps}
        for c in self._comps:
            d_in = {c.name: _Vec({k: np.zeros_like(v) for k, v in self.inp[c.name].items()})}
            self._transfer(c, out=d_out, inp=d_in)
            c.apply_linear(self.inp[c.name], self.out[c.name], d_in[c.name], d_out[c.name], d_res[c.name], 'fwd')
        return self._pack(d_res)

    def _block_jacobi(self, r):
        d_res = self._unpack(r)
        d_out = {c.name: _Vec({k: np.zeros_like(v) for k, v in self.out[c.name].items()}) for c in self._comps}
        for c in self._comps:
            c.solve_linear(d_out[c.name], d_res[c.name], 'fwd')
        return self._pack(d_out)

    def _run(self):
        nl = self.nonlinear_solver
        atol = nl.options.get('atol', 1e-10)
        maxiter = nl.options.get('maxiter', 10)
        if isinstance(nl, NonlinearBlockGS):
            for it in range(maxiter):
                for c in self._comps:
                    self._transfer(c)
                    c.solve_nonlinear(self.inp[c.name], self.out[c.name])
                if self._apply_nonlinear() <= atol:
                    return it + 1
            raise RuntimeError('NonlinearBlockGS failed to converge')
        # NewtonSolver(solve_subsystems=True): sub-system solve first, then Newton on the coupled residual
        if nl.options.get('solve_subsystems'):
            for c in self._comps:
                self._transfer(c)
                c.solve_nonlinear(self.inp[c.name], self.out[c.name])
        rn = self._apply_nonlinear()
        for it in range(maxiter):
            if rn <= atol:
                return it
            for c in self._comps:
                c.linearize(self.inp[c.name], self.out[c.name], None)
            r = -self._pack(self.res)
            ls = self.linear_solver
            if isinstance(ls, ScipyKrylov):
                n = r.size
                A = spla.LinearOperator((n, n), matvec=self._jvp, dtype=float)
                M = spla.LinearOperator((n, n), matvec=self._block_jacobi, dtype=float) if ls.precon is not None else None
                dx, info = spla.gmres(A, r, M=M, restart=ls.options.get('restart', 20), maxiter=ls.options.get('maxiter', 1000),
                                      atol=ls.options.get('atol', 1e-12), rtol=0.0)
                if info != 0 and ls.options.get('err_on_non_converge'):
                    raise RuntimeError('ScipyKrylov failed to converge')
            else:
                dx = self._block_jacobi(r)
            x0 = self._pack(self.out)
            alpha = 1.0
            if nl.linesearch is not None:
                o = nl.linesearch.options
                for _ in range(o.get('maxiter', 5) + 1):
                    self._set_outputs(x0 + alpha * dx)
                    rt = self._apply_nonlinear()
                    if rt <= rn - o.get('c', 0.1) * alpha * rn:
                        break
                    alpha *= o.get('rho', 0.5)
                rn = rt
            else:
                self._set_outputs(x0 + dx)
                rn = self._apply_nonlinear()
        if rn <= atol:
            return maxiter
        raise RuntimeError('NewtonSolver failed to converge')

    def _set_outputs(self, x):
        d = self._unpack(x)
        for c, k, _ in self._layout():
            self.out[c][k] = d[c][k]


class Problem:
    def __init__(self):
        self.model = Group()

    def setup(self):
        self.model._setup()

    def run_model(self):
        for s in self.model._subs:
            if isinstance(s, Group) and s.nonlinear_solver is not None:
                s.iterations = s._run()
