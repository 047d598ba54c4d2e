"""ctypes binding of libsem_b200.so (the C ABI declared in include/sem_b200.h).

The CUDA library is the product: there is NO CPU fallback.  Importing this module without the built library, or
calling into it without a CUDA device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEM_B200_LIB") or os.path.join(_HERE, "libsem_b200.so")   # override: A/B builds only


class SemError(RuntimeError):
    pass


class sem_mesh_desc(C.Structure):
    _fields_ = [("P", C.c_int), ("N_ex", C.c_int), ("N_ey", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("D", C.c_void_p), ("Ks", C.c_void_p), ("w", C.c_void_p), ("device", C.c_int),
                ("m_begin", C.c_int), ("m_end", C.c_int)]


class sem_cd_bc(C.Structure):
    _fields_ = [("active", C.c_int * 4), ("value", C.c_double * 4)]


class sem_ns_bc(C.Structure):
    _fields_ = [("v_W", C.c_double), ("v_E", C.c_double), ("u_S", C.c_double), ("u_N", C.c_double)]


class sem_cd_state(C.Structure):
    _fields_ = [("bc", sem_cd_bc), ("Pe", C.c_double), ("u", C.c_void_p), ("v", C.c_void_p),
                ("gxT", C.c_void_p), ("gyT", C.c_void_p)]


class sem_ns_state(C.Structure):
    _fields_ = [("bc", sem_ns_bc), ("Re", C.c_double), ("Gr_over_Re", C.c_double), ("u", C.c_void_p),
                ("v", C.c_void_p), ("gxu", C.c_void_p), ("gyu", C.c_void_p), ("gxv", C.c_void_p), ("gyv", C.c_void_p)]


class sem_fdm_dir(C.Structure):
    _fields_ = [("lo", C.c_int), ("cnt", C.c_int), ("fold", C.c_int), ("Qe", C.c_void_p), ("Qo", C.c_void_p),
                ("lam", C.c_void_p), ("nmodes", C.c_int)]


class sem_ns_schur_desc(C.Structure):
    _fields_ = [("wl", C.c_void_p), ("wr", C.c_void_p),
                ("Tinv_x", C.c_void_p), ("Tinv_y", C.c_void_p),
                ("lfx", C.c_void_p), ("lfy", C.c_void_p), ("singular", C.c_int), ("two_level", C.c_int),
                ("inv_den", C.c_double), ("cheb_lo", C.c_double), ("cheb_hi", C.c_double), ("cheb_steps", C.c_int)]


class sem_transfer(C.Structure):
    _fields_ = [("nxp", C.c_int), ("nyp", C.c_int), ("mx", C.c_void_p), ("ny", C.c_void_p), ("Sx", C.c_void_p), ("Sy", C.c_void_p)]


class sem_krylov(C.Structure):
    _fields_ = [("atol", C.c_double), ("restart", C.c_int), ("max_iters", C.c_int), ("precond", C.c_int),
                ("verbose", C.c_int), ("iters", C.c_int), ("resnorm", C.c_double)]


class sem_coupled(C.Structure):
    _fields_ = [("ns", C.c_void_p), ("cd", C.c_void_p), ("ns_state", C.POINTER(sem_ns_state)), ("cd_state", C.POINTER(sem_cd_state)),
                ("ns_to_cd", sem_transfer), ("cd_to_ns", sem_transfer), ("kr_ns", C.POINTER(sem_krylov)),
                ("kr_cd", C.POINTER(sem_krylov)), ("ns_work", C.c_void_p), ("ns_work_len", C.c_longlong),
                ("cd_work", C.c_void_p), ("cd_work_len", C.c_longlong), ("ns_null", C.c_void_p), ("ns_null_nrm2", C.c_double),
                ("iters_cd", C.c_int), ("iters_ns", C.c_int), ("solves", C.c_int)]


_P = C.c_void_p
_LL = C.c_longlong
# name -> (restype, argtypes); every symbol include/sem_b200.h declares
SIGNATURES = {
    "sem_last_error": (C.c_char_p, []),
    "sem_version": (C.c_int, []),
    "sem_ctx_create": (C.c_int, [C.POINTER(_P), C.POINTER(sem_mesh_desc)]),
    "sem_ctx_destroy": (None, [_P]),
    "sem_ctx_ld": (C.c_int, [_P]),
    "sem_ctx_nx": (C.c_int, [_P]),
    "sem_ctx_ny": (C.c_int, [_P]),
    "sem_ctx_vec_len": (_LL, [_P]),
    "sem_ctx_set_tiling": (C.c_int, [_P, C.c_int, C.c_int]),
    "sem_nccl_unique_id": (C.c_int, [_P]),
    "sem_ctx_attach_comm": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "sem_ctx_comm_mode": (C.c_int, [_P]),
    "sem_ctx_attach_loopback": (C.c_int, [_P]),
    "sem_ctx_partitioned_applies": (_LL, [_P, C.c_int]),
    "sem_h2d": (C.c_int, [_P, _P, _P, _P]),
    "sem_d2h": (C.c_int, [_P, _P, _P, _P]),
    "sem_apply_stiffness": (C.c_int, [_P, _P, _P, _P]),
    "sem_apply_gradient": (C.c_int, [_P, _P, C.c_double, _P, _P, _P]),
    "sem_apply_mass": (C.c_int, [_P, _P, _P, _P]),
    "sem_mass_diag": (C.c_int, [_P, _P, _P]),
    "sem_gather_scatter": (C.c_int, [_P, _P, _P, _P]),
    "sem_scatter": (C.c_int, [_P, _P, _P, _P]),
    "sem_interpolate": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P]),
    "sem_cd_residual": (C.c_int, [_P, C.POINTER(sem_cd_state), _P, _P, _P]),
    "sem_cd_jacobians": (C.c_int, [_P, C.c_double, _P, _P, _P, _P]),
    "sem_cd_jvp": (C.c_int, [_P, C.POINTER(sem_cd_state), _P, _P, _P, _P, _P]),
    "sem_ctx_set_fdm": (C.c_int, [_P, C.c_int, C.POINTER(sem_fdm_dir), C.POINTER(sem_fdm_dir), C.c_int, C.c_double]),
    "sem_fdm_apply": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "sem_ctx_set_ns_schur": (C.c_int, [_P, C.POINTER(sem_ns_schur_desc)]),
    "sem_ns_precond_debug": (C.c_int, [_P, C.POINTER(sem_ns_state), C.c_int, C.c_int, _P, _P, _P, _P]),
    "sem_cd_jvp_host": (C.c_int, [_P, C.POINTER(sem_cd_state), _P, _P, _P, _P, _P]),
    "sem_cd_work_len": (_LL, [_P, C.c_int]),
    "sem_cd_solve": (C.c_int, [_P, C.POINTER(sem_cd_state), _P, _P, C.POINTER(sem_krylov), _P, _LL, _P]),
    "sem_ns_residual": (C.c_int, [_P, C.POINTER(sem_ns_state), _P, _P, _P, _P, _P, _P, _P, _P]),
    "sem_ns_jacobians": (C.c_int, [_P, C.c_double, _P, _P, _P, _P, _P, _P, _P]),
    "sem_ns_jvp": (C.c_int, [_P, C.POINTER(sem_ns_state), _P, _P, _P, _P, _P, _P, _P, _P]),
    "sem_ns_work_len": (_LL, [_P, C.c_int]),
    "sem_ns_solve": (C.c_int, [_P, C.POINTER(sem_ns_state), _P, _P, C.POINTER(sem_krylov), _P, _LL, _P]),
    "sem_coupled_vec_len": (_LL, [C.POINTER(sem_coupled)]),
    "sem_coupled_work_len": (_LL, [C.POINTER(sem_coupled), C.c_int]),
    "sem_coupled_jvp": (C.c_int, [C.POINTER(sem_coupled), _P, _P, _P, _P]),
    "sem_coupled_solve": (C.c_int, [_P, C.POINTER(sem_coupled), _P, _P, C.POINTER(sem_krylov), _P, _LL, _P]),
    "sem_dot": (C.c_int, [_P, _P, _P, _LL, C.POINTER(C.c_double), _P]),
    "sem_axpby": (C.c_int, [_P, C.c_double, _P, C.c_double, _P, _LL, _P]),
}

_lib = None


def load():
    """Load libsem_b200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SemError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C sem_b200/csrc -j8` (sem_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what):
    """Map a negative return code to an exception carrying the library's error text."""
    if code < 0:
        raise SemError(f"{what} failed ({code}): {load().sem_last_error().decode(errors='replace')}")
    return code
