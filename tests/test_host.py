"""CPU: host-side logic of the product and the C ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, relerr
from tests.golden.make_golden_cases import MESHES


def test_library_exports_every_declared_symbol():
    """libsem_b200.so loads and exports exactly the entry points include/sem_b200.h declares."""
    from sem_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sem_b200.h")).read()
    declared = set(re.findall(r"\b(sem_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found in the header"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in sem_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert _lib.load().sem_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sem_b200
    from sem_b200._lib import SemError
    with pytest.raises(SemError):
        sem_b200.ConvectionDiffusionSolver(1, 1, 40, 4, 4, 4)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: no module of the product may import it or touch a name called `oracle` (docstrings
    may cite oracle/*.py as the place where a stage is restated)."""
    import ast
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sem_b200")):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                if isinstance(node, ast.Import):
                    assert not any(a.name.split(".")[0] == "oracle" for a in node.names), f"{f} imports the oracle"
                elif isinstance(node, ast.ImportFrom):
                    assert (node.module or "").split(".")[0] != "oracle", f"{f} imports from the oracle"
                    assert not any(a.name == "oracle" for a in node.names), f"{f} imports the oracle"
                elif isinstance(node, ast.Name):
                    assert node.id != "oracle", f"{f} uses a name `oracle`"
                elif isinstance(node, ast.Call) and getattr(node.func, "id", getattr(node.func, "attr", "")) in ("__import__", "import_module"):
                    assert not any(isinstance(a, ast.Constant) and str(a.value).startswith("oracle") for a in node.args), f"{f}"


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16])
def test_host_gll_matches_reference(golden, P):
    from sem_b200 import GLL
    g = golden("gll")
    x, w, _ = GLL.standard_nodes(P)
    assert np.array_equal(x, g[f"x{P}"]) and np.array_equal(w, g[f"w{P}"])
    assert np.array_equal(GLL.standard_differentiation_matrix(P), g[f"D{P}"])
    assert np.array_equal(GLL.standard_stiffness_matrix(P), g[f"K{P}"])
    assert np.array_equal(GLL.standard_gradient_matrix(P), g[f"G{P}"])
    assert np.array_equal(GLL.standard_mass_matrix(P), np.diag(g[f"w{P}"]))
    assert np.allclose(GLL.standard_evaluation_matrix(P, np.linspace(-1, 1, 7)), g[f"S{P}"], rtol=0, atol=1e-13)
    F, Cm = GLL.standard_product_matrix(P), GLL.standard_convection_matrix(P)
    assert np.array_equal(np.einsum('iii->i', F), w) and abs(F.sum() - w.sum()) < 1e-15
    assert np.array_equal(np.einsum('iik->ik', Cm), g[f"G{P}"])


@pytest.mark.parametrize("tag,P,nx,ny,Lx,Ly", MESHES)
def test_host_mesh_functions_match_reference(golden, tag, P, nx, ny, Lx, Ly):
    from sem_b200 import SEM
    g = golden("operators")
    assert np.array_equal(SEM.global_nodes(P, nx, ny, Lx / nx, Ly / ny), g[f"{tag}/points"])
    pts_e = SEM.element_nodes(P, nx, ny, Lx / nx, Ly / ny)
    val = SEM.eval_interpolation(g[f"{tag}/scattered"], pts_e, (g[f"{tag}/xp"], g[f"{tag}/yp"]))
    assert relerr(val, g[f"{tag}/interp"]) < 1e-13
    assert SEM.global_index(P, nx, ny, nx - 1, ny - 1, P, P) == (nx * P + 1) * (ny * P + 1) - 1
    with pytest.raises(ValueError):
        SEM.global_index(P, nx, ny, nx, 0, 0, 0)
    with pytest.raises(ValueError):
        SEM.xi2x(0, np.array([1.5]), 1.0)
    e, xi = SEM.x2xi(np.array([0.0, Lx / nx, Lx]), Lx / nx)
    assert list(e) == [0, 0, nx - 1] and np.allclose(xi, [-1, 1, 1])


def test_pinned_result_pool_recycles_only_dead_arrays(monkeypatch):
    """Result arrays are fresh per call; their page-locked block returns to the pool only after the array AND all of its
    views are gone (sem_b200/device.py)."""
    import gc
    import torch
    import sem_b200.device as D
    monkeypatch.setattr(D, "_alloc_pinned", lambda n: torch.empty(n, dtype=torch.float64))   # no CUDA on the CPU box
    monkeypatch.setattr(D, "_PINNED_POOL", {})
    a = D._pinned_result(12)
    a[:] = 3.0
    view = a[2:5].reshape(3, 1)
    b = D._pinned_result(12)
    assert not np.shares_memory(a, b)
    del a
    gc.collect()
    assert D._PINNED_POOL.get(12, []) == []          # the view keeps the block out of the pool
    assert float(view[0, 0]) == 3.0
    del view
    gc.collect()
    assert len(D._PINNED_POOL[12]) == 1
    c = D._pinned_result(12)                         # recycled block
    assert D._PINNED_POOL[12] == [] and c.shape == (12,)


@pytest.mark.parametrize("P,nx,ny,Lx,Ly", [(4, 5, 3, 1.3, 0.7), (3, 4, 4, 1.0, 1.0), (8, 2, 3, 2.0, 1.0), (1, 3, 2, 1.0, 1.0)])
def test_pressure_boundary_block_matches_oracle(P, nx, ny, Lx, Ly):
    """Boundary-ring block of the NS preconditioner: node list and K_BB against the
    boundary rows / columns of the oracle's assembled Jacobian (rows K[mask,:] of NS:119,157)."""
    from oracle import sem_oracle as so   # checker
    from sem_b200 import SEM
    ns = so.NSOracle(Lx, Ly, 10.0, 0.0, P, nx, ny, u_N=1.0)
    z = np.zeros(ns.N)
    ns._get_residuals(z, z, z, z)
    ns._calc_jacobians(z, z)
    D = ns.jacobian_matrix().tocsr()[2 * ns.N:, 2 * ns.N:]
    ix, iy, KBB = SEM.pressure_boundary_block(P, nx, ny, Lx / nx, Ly / ny, pin=ns._pin)
    g = iy + (ny * P + 1) * ix
    assert sorted(g.tolist()) == sorted(np.where(ns._mask_bound & (np.arange(ns.N) != ns._pin))[0].tolist())
    ref = D[g][:, g].toarray()
    assert np.abs(KBB - ref).max() <= 1e-13 * np.abs(ref).max()
    # the device replaces K_BB^-1 by a fixed Chebyshev polynomial of the diagonally scaled block: the bounds handed to
    # sem_ctx_set_ns_schur (SEM.ring_chebyshev_parameters, sparse assembly from the 1-D matrices + Lanczos) must enclose its
    # spectrum, and the CPU mirror's restatement must agree
    from oracle.ns_precond import ring_chebyshev
    sd = 1.0 / np.sqrt(np.diag(KBB))
    ev = np.linalg.eigvalsh(KBB * sd[:, None] * sd[None, :])
    lo, hi, steps = SEM.ring_chebyshev_parameters(P, nx, ny, Lx / nx, Ly / ny, pin=ns._pin)
    assert lo <= ev.min() <= 1.03 * lo and 0.97 * hi <= ev.max() <= hi and 1 <= steps <= 8
    ring = ns._mask_bound.copy()
    ring[ns._pin] = False
    lo2, hi2, steps2 = ring_chebyshev(ns._K, ring)
    assert abs(lo - lo2) < 1e-8 and abs(hi - hi2) < 1e-8 and steps == steps2


def test_header_is_plain_c():
    """include/sem_b200.h is the drop-in boundary: it must compile as C99 on its own (no C++ or torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "sem_b200.h")
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_assemble_matrix_and_tensor_element_arrays():
    """SEM.assemble for 6- and 8-index element arrays (SEM.py:132-145) -- host paths, no GPU: against a brute-force dense
    accumulation over the global index map, and the 6-index stiffness against the oracle's assembled matrix."""
    from sem_b200 import SEM, GLL
    from oracle import sem_oracle as so
    P, nx, ny, dx, dy = 3, 3, 2, 0.7, 1.3
    n, N = P + 1, (P * nx + 1) * (P * ny + 1)
    rng = np.random.default_rng(5)
    A6 = rng.standard_normal((nx, ny, n, n, n, n)) * (rng.random((nx, ny, n, n, n, n)) < 0.3)
    dense = np.zeros((N, N))
    A8 = rng.standard_normal((nx, ny, n, n, n, n, n, n)) * (rng.random((nx, ny, n, n, n, n, n, n)) < 0.05)
    dense3 = np.zeros((N, N, N))
    gi = lambda m, e, i, j: int(SEM.global_index(P, nx, ny, m, e, i, j))
    for m in range(nx):
        for e in range(ny):
            for i in range(n):
                for j in range(n):
                    a = gi(m, e, i, j)
                    for k in range(n):
                        for l in range(n):
                            b = gi(m, e, k, l)
                            dense[a, b] += A6[m, e, i, j, k, l]
                            for r in range(n):
                                for q in range(n):
                                    dense3[a, b, gi(m, e, r, q)] += A8[m, e, i, j, k, l, r, q]
    assert np.abs(SEM.assemble(A6).toarray() - dense).max() < 1e-14
    C = SEM.assemble(A8)
    u = rng.standard_normal(N)
    assert np.abs(SEM.tensordot(C, u, (1, 0)).toarray() - np.einsum('abc,b->ac', dense3, u)).max() < 1e-12
    assert np.abs(SEM.tensordot(C, u, (2, 0)).toarray() - np.einsum('abc,c->ab', dense3, u)).max() < 1e-12
    assert np.abs((C @ u).toarray() - np.einsum('abc,c->ab', dense3, u)).max() < 1e-12
    # the element stiffness array of SEM.py:186-203, assembled, is the oracle's K
    w, Ks = GLL.standard_nodes(P)[1], GLL.standard_stiffness_matrix(P)
    Ke = (2.0 / dx) * (dy / 2.0) * np.einsum('ik,j,jl->ijkl', Ks, w, np.eye(n)) \
        + (dx / 2.0) * (2.0 / dy) * np.einsum('i,ik,jl->ijkl', w, np.eye(n), Ks)
    K = SEM.assemble(np.broadcast_to(Ke, (nx, ny) + Ke.shape).copy())
    Ko = so.global_operators(P, nx, ny, dx, dy)[1]
    assert abs(K - Ko).max() < 1e-12
    with pytest.raises(ValueError):
        SEM.assemble(np.zeros((2, 2, 3)))


def test_public_signatures_match_the_reference():
    """Drop-in boundary: every public function / method of the reference's hot-path modules (tests/golden/api_signatures.json,
    recorded from /root/reference with `ast` by tests/golden/make_api_signatures.py) exists here with the same parameter names in
    the same order and the same defaults; the drop-in may only append keyword parameters (device, partition, precond ...)."""
    import ast
    import inspect
    import json
    import sem_b200
    from sem_b200 import GLL, SEM, Boussinesq_SequentialCoupler
    api = json.load(open(os.path.join(ROOT, "tests", "golden", "api_signatures.json")))
    homes = {"ConvectionDiffusion_Solver": sem_b200, "NavierStokes_Solver": sem_b200, "SEM": SEM, "GLL": GLL,
             "Boussinesq_SequentialCoupler": Boussinesq_SequentialCoupler}

    def same_default(ref_src, got):
        if ref_src is None:
            return got is inspect.Parameter.empty
        if got is inspect.Parameter.empty:
            return False
        try:
            return ast.literal_eval(ref_src) == got
        except (ValueError, SyntaxError):
            return True                           # a non-literal default of the reference (none on this path today)

    missing, wrong = [], []
    for mod, entry in api.items():
        home = homes[mod]
        items = [(f"{mod}.{n}", getattr(home, n, None), ps) for n, ps in entry["functions"].items()]
        for cname, methods in entry["classes"].items():
            cls = getattr(home, cname, None)
            items += [(f"{cname}.{n}", getattr(cls, n, None) if cls else None, ps) for n, ps in methods.items()]
        for name, obj, ref_params in items:
            if obj is None:
                missing.append(name)
                continue
            got = list(inspect.signature(obj).parameters.values())
            for i, rp in enumerate(ref_params):
                if i >= len(got) or got[i].name != rp["name"] or not same_default(rp["default"], got[i].default):
                    wrong.append(f"{name}: parameter {i} is {got[i] if i < len(got) else None}, reference {rp}")
                    break
            else:
                extra = got[len(ref_params):]
                if any(p.default is inspect.Parameter.empty and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL) for p in extra):
                    wrong.append(f"{name}: extra parameters without defaults {extra}")
    assert not missing, f"missing: {missing}"
    assert not wrong, "\n".join(wrong)
