"""Record what the reference's OWN callers ask of the two solver classes (run in the build container only).

    python tests/golden/make_component_trace.py        # writes tests/golden/component_trace.npz

The UNMODIFIED adapters ``OpenMDAO/ConvectionDiffusion_Component.py`` / ``NavierStokes_Component.py`` and the UNMODIFIED coupler
script ``OpenMDAO/Boussinesq_SequentialCoupler.py`` are imported from /root/reference and run over the reference's solvers
(through oracle/ref_shim.py) with the ``openmdao.api`` stand-in of tests/golden/openmdao_stub (openmdao ~=3.9.2 is not
installable here).  A recording proxy around the two solver objects logs every call that crosses the drop-in boundary: method
name, array arguments, array results.  ``tests/test_gpu_parity.py::test_reference_callers_on_the_drop_in`` replays the trace on
the GPU classes (applies <= 1e-12, solves <= 1e-8); ``tests/test_boundary.py`` drives the same unmodified-in-spirit adapters
over the drop-in.  The mesh is small (P = 3, 4 x 4 elements for NS, 2 x 2 for CD -- the study's N_e / 2 rule,
study/Boussinesq_run.py:50) so that the trace stays a small fixture; mode GS, then mode JNK."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "openmdao_stub"))
from oracle import ref_shim  # noqa: E402

GLL, SEM, CDS, NSS = ref_shim.install()
import Solvers.ConvectionDiffusion_Solver as cd_mod  # noqa: E402
import Solvers.NavierStokes_Solver as ns_mod  # noqa: E402

TRACED = ('_get_residuals', '_calc_jacobians', '_get_dresiduals', '_get_update', '_get_solution', '_get_interpol')
LOG = []


def traced_class(base, tag):
    """Subclass of the reference solver whose boundary methods append (tag, method, args, results) to LOG."""
    ns = {}
    for name in TRACED:
        def make(name):
            orig = getattr(base, name)

            def wrapper(self, *args, **kwargs):
                if getattr(self, '_depth', 0):                 # calls the solver makes to itself are not boundary calls
                    return orig(self, *args, **kwargs)
                self._depth = 1
                try:
                    a_in = [None if a is None else np.array(a, dtype=float) for a in args]
                    k_in = {k: (None if v is None else np.array(v, dtype=float)) for k, v in kwargs.items()}
                    out = orig(self, *args, **kwargs)
                finally:
                    self._depth = 0
                outs = () if out is None else (out if isinstance(out, tuple) else (out,))
                LOG.append((tag, name, a_in, k_in, [np.array(o, dtype=float) for o in outs]))
                return out
            return wrapper
        ns[name] = make(name)
    return type(base.__name__, (base,), ns)


def main():
    # the coupler script builds its own solver objects: hand it the recording subclasses under the reference's names
    cd_mod.ConvectionDiffusionSolver = traced_class(CDS, 'cd')
    ns_mod.NavierStokesSolver = traced_class(NSS, 'ns')
    sys.modules.pop('OpenMDAO.Boussinesq_SequentialCoupler', None)
    from OpenMDAO import Boussinesq_SequentialCoupler as bsc      # the reference's file, unmodified
    xp, yp = np.meshgrid(np.linspace(0, 1, 7), np.linspace(0, 1, 5), indexing='ij')
    out = {}
    kw = dict(Re=50.0, Ra=200.0, Pr=0.71, P_cd=3, N_ex_cd=2, N_ey_cd=2, P_ns=3, N_ex_ns=4, N_ey_ns=4)
    for mode in ('GS', 'JNK'):
        LOG.clear()
        T, u, v = bsc.run((xp, yp), 1.0, 1.0, mode=mode, mtol_nonlin=1e-9, mtol_gmres=1e-11, **kw)
        print(f"mode {mode}: {len(LOG)} boundary calls; T range {T.min():.4f} .. {T.max():.4f}")
        out[f"{mode}/n"] = np.array(len(LOG))
        for i, (tag, name, a_in, k_in, outs) in enumerate(LOG):
            out[f"{mode}/{i}/who"] = np.array(f"{tag}.{name}")
            for j, a in enumerate(a_in):
                if a is not None:
                    out[f"{mode}/{i}/arg{j}"] = a
            for k, a in k_in.items():
                if a is not None:
                    out[f"{mode}/{i}/kw_{k}"] = a
            for j, o in enumerate(outs):
                out[f"{mode}/{i}/out{j}"] = o
        out[f"{mode}/T_plot"], out[f"{mode}/u_plot"], out[f"{mode}/v_plot"] = T, u, v
    out["xp"], out["yp"] = xp, yp
    for k, v in kw.items():
        out[f"kw/{k}"] = np.array(v)
    path = os.path.join(HERE, "component_trace.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
