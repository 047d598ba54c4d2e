"""TEST INFRASTRUCTURE ONLY -- makes the read-only upstream reference importable in the build container.

The reference (``/root/reference``) imports two things this image does not provide:

* the third-party package ``sparse`` (pydata/sparse ~=0.12.0, ``requirements.txt:3``), used only for
  ``sparse.COO(coords, data, shape=...)`` (``Solvers/SEM.py:145``) and
  ``sparse.tensordot(C, vec, (axis, 0), return_type=sparse.COO).tocsr()``
  (``Solvers/ConvectionDiffusion_Solver.py:82-83,101-102``, ``Solvers/NavierStokes_Solver.py:103-104,132-136``);
* ``scipy.sparse.linalg.lgmres(..., tol=0, ...)`` -- scipy >= 1.14 renamed ``tol`` to ``rtol``.

This module installs a minimal stand-in for both (published semantics restated, nothing copied) and puts the
reference on ``sys.path``.  It is used ONLY by ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` exists) to generate the committed golden vectors.  Nothing in the product, in ``bench.py`` or in
the ``-m gpu`` tests may import it: ``/root/reference`` does not exist on the GPU box.
"""
import sys
import types

import numpy as np
import scipy.sparse as sp_sparse
import scipy.sparse.linalg as sp_linalg

REFERENCE_ROOT = "/root/reference"


class _COO:
    """Three-index COO tensor with duplicate entries allowed (they are summed on contraction/conversion)."""

    def __init__(self, coords, data=None, shape=None):
        self.coords = np.asarray(coords).astype(np.int64)
        self.data = np.asarray(data, dtype=np.float64)
        self.shape = tuple(shape)
        self.ndim = len(self.shape)

    def tocsr(self):
        if self.ndim != 2:
            raise ValueError("tocsr needs a 2-d COO")
        return sp_sparse.coo_matrix((self.data, (self.coords[0], self.coords[1])), shape=self.shape).tocsr()


def _tensordot(a, b, axes=2, return_type=None):
    """Contract axis ``axes[0]`` of the COO tensor ``a`` with axis 0 of the dense vector ``b``."""
    ax_a, ax_b = axes
    if ax_b != 0 or np.ndim(b) != 1:
        raise NotImplementedError("stand-in covers tensor-by-vector contraction only")
    keep = [k for k in range(a.ndim) if k != ax_a]
    data = a.data * np.asarray(b, dtype=np.float64)[a.coords[ax_a]]
    return _COO(a.coords[keep], data, tuple(a.shape[k] for k in keep))


def install():
    """Install the stand-ins and return the imported reference modules (GLL, SEM, CD class, NS class)."""
    if "sparse" not in sys.modules:
        mod = types.ModuleType("sparse")
        mod.COO = _COO
        mod.tensordot = _tensordot
        sys.modules["sparse"] = mod

    if not getattr(sp_linalg.lgmres, "_tol_adapter", False):
        _orig = sp_linalg.lgmres

        def lgmres(A, b, x0=None, *, tol=None, rtol=1e-5, atol=0.0, **kw):
            if tol is not None:
                rtol = tol
            return _orig(A, b, x0=x0, rtol=rtol, atol=atol, **kw)

        lgmres._tol_adapter = True
        sp_linalg.lgmres = lgmres

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # The reference's ``Solvers`` directory has no __init__.py (a namespace package), so the repository's own ``Solvers/`` shim
    # package (a regular package) would win the import no matter the order of sys.path: bind the name to the reference's
    # directory explicitly before anything is imported from it.
    import os
    pkg = sys.modules.get("Solvers")
    ref_dir = os.path.join(REFERENCE_ROOT, "Solvers")
    if pkg is None or list(getattr(pkg, "__path__", [])) != [ref_dir]:
        for name in [m for m in sys.modules if m == "Solvers" or m.startswith("Solvers.")]:
            del sys.modules[name]
        pkg = types.ModuleType("Solvers")
        pkg.__path__ = [ref_dir]
        sys.modules["Solvers"] = pkg
    from Solvers import GLL, SEM
    from Solvers.ConvectionDiffusion_Solver import ConvectionDiffusionSolver
    from Solvers.NavierStokes_Solver import NavierStokesSolver
    return GLL, SEM, ConvectionDiffusionSolver, NavierStokesSolver
