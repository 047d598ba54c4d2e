"""sem_b200 -- B200-native (sm_100a) hot path of the 2-D spectral element solver Tangxiaotian11/SEM.

Drop-in classes (same names and signatures as the reference's Solvers package):
    from sem_b200 import ConvectionDiffusionSolver, NavierStokesSolver
or, unchanged reference-style imports through the top-level ``Solvers`` shim package of this repository:
    from Solvers.NavierStokes_Solver import NavierStokesSolver
"""
import importlib

from . import GLL  # noqa: F401

__all__ = ["GLL", "SEM", "ConvectionDiffusionSolver", "NavierStokesSolver", "SemDevice", "Boussinesq_SequentialCoupler"]

_LAZY = {
    "SEM": ("sem_b200.SEM", None),
    "SemDevice": ("sem_b200.device", "SemDevice"),
    "ConvectionDiffusionSolver": ("sem_b200.ConvectionDiffusion_Solver", "ConvectionDiffusionSolver"),
    "NavierStokesSolver": ("sem_b200.NavierStokes_Solver", "NavierStokesSolver"),
    "Boussinesq_SequentialCoupler": ("sem_b200.Boussinesq_SequentialCoupler", None),
}


def __getattr__(name):
    # torch / CUDA are only imported when the device-backed pieces are touched
    if name in _LAZY:
        mod, attr = _LAZY[name]
        m = importlib.import_module(mod)
        return m if attr is None else getattr(m, attr)
    raise AttributeError(name)
