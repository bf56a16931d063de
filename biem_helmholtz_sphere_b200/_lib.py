"""ctypes binding of libbhs.so (include/bhs.h).  No fallback: a missing library raises."""

from __future__ import annotations

import ctypes as C
import os
import weakref

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbhs.so")

_lib = None

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f64 = C.c_double

# name -> (restype, argtypes); every symbol declared in include/bhs.h
SIGNATURES = {
    "bhs_version": (i32, []),
    "bhs_device_sm_count": (i32, [C.POINTER(i32)]),
    "bhs_plan_create": (i32, [i32, i32, C.POINTER(vp)]),
    "bhs_plan_create_tree": (i32, [i32, i32, i32, C.POINTER(vp)]),
    "bhs_plan_destroy": (None, [vp]),
    "bhs_plan_harm": (i32, [vp]),
    "bhs_plan_harm2": (i32, [vp]),
    "bhs_plan_quad_points": (i32, [vp]),
    "bhs_plan_index_table": (i32, [vp, vp]),
    "bhs_plan_quadrature": (i32, [vp, vp, vp]),
    "bhs_plan_coupling_stats": (i32, [vp, C.POINTER(i64), C.POINTER(i64)]),
    "bhs_bessel": (i32, [i32, i32, i32, i32, vp, i64, vp, vp]),
    "bhs_bessel_z": (i32, [i32, i32, i32, i32, vp, vp, i64, vp, vp]),
    "bhs_harmonics": (i32, [vp, i32, vp, i64, vp, vp]),
    "bhs_rhs_expand": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "bhs_assemble_workspace": (i64, [vp, i32, i32]),
    "bhs_assemble": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp]),
    "bhs_assemble_rows": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i64, i64, vp, vp]),
    "bhs_diag_coef_workspace": (i64, [vp, i32, i32]),
    "bhs_diag_coef": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "bhs_zgesv_workspace": (i64, [i64, i32]),
    "bhs_zgesv": (i32, [i64, i32, vp, i64, vp, vp, vp, vp, vp]),
    "bhs_zgesv_batched_workspace": (i64, [i64, i32, i32]),
    "bhs_zgesv_batched": (i32, [i64, i32, i32, vp, i64, i64, vp, i64, vp, vp, vp, vp]),
    "bhs_zgetrf": (i32, [i64, vp, i64, vp, vp, vp, vp]),
    "bhs_zgetrs": (i32, [i64, i32, vp, i64, vp, vp, vp, vp]),
    "bhs_zgemm_workspace": (i64, [i64, i64, i64]),
    "bhs_zgemm_sub": (i32, [i64, i64, i64, vp, i64, vp, i64, vp, i64, vp, vp]),
    "bhs_uscat_workspace": (i64, [vp, i32]),
    "bhs_uscat": (i32, [vp, i32, vp, vp, f64, f64, f64, vp, vp, i64, i32, vp, vp, vp]),
    "bhs_fp64_peak": (i32, [i32, i32, C.POINTER(f64)]),
    "bhs_launch_count": (i64, [i32]),
    "bhs_profile": (i32, [i32]),
    "bhs_profile_read": (i32, [i32, C.POINTER(f64), C.POINTER(f64), C.POINTER(i64)]),
}

KIND_J, KIND_Y, KIND_H1 = 0, 1, 2
PROF_CATEGORIES = ("lu_gemm", "lu_panel", "lu_trsm", "lu_pack", "lu_rhs", "asm_main", "asm_pre", "uscat", "rhs_expand", "lu_gemm_inner")
FLAG_PER_BALL, FLAG_FAR_FIELD, FLAG_INNER = 1, 2, 4
TREE_CHAIN, TREE_HOPF = 0, 1


class BhsError(RuntimeError):
    pass


def load():
    """Load libbhs.so (once) and attach the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BhsError(
            f"{LIB_PATH} not found: build it with `python -m biem_helmholtz_sphere_b200.build` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "libbhs call") -> None:
    if rc == 0:
        return
    if rc == -1:
        raise ValueError(f"{what}: invalid argument")
    if rc == -2:
        raise NotImplementedError(f"{what}: outside the implemented range")
    if rc == -3:
        raise MemoryError(f"{what}: allocation failed")
    raise BhsError(f"{what}: CUDA error {rc}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Plan:
    """Owner of a bhs_plan_t* for (d, n_end[, tree]) on the current device."""

    def __init__(self, d: int, n_end: int, tree: int = TREE_CHAIN):
        lib = load()
        h = vp()
        check(lib.bhs_plan_create_tree(d, n_end, tree, C.byref(h)), "bhs_plan_create_tree")
        self.handle = h
        self.tree = tree
        self.d = d
        self.n_end = n_end
        self.H = lib.bhs_plan_harm(h)
        self.H2 = lib.bhs_plan_harm2(h)
        self.Q = lib.bhs_plan_quad_points(h)
        self._fin = weakref.finalize(self, lib.bhs_plan_destroy, h)

    def index_table(self):
        import numpy as np

        out = np.empty((self.H, self.d - 1), dtype=np.int32)
        check(load().bhs_plan_index_table(self.handle, out.ctypes.data_as(vp)), "bhs_plan_index_table")
        return out

    def quadrature(self):
        import numpy as np

        dirs = np.empty((self.d, self.Q), dtype=np.float64)
        w = np.empty((self.Q,), dtype=np.float64)
        check(
            load().bhs_plan_quadrature(self.handle, dirs.ctypes.data_as(vp), w.ctypes.data_as(vp)),
            "bhs_plan_quadrature",
        )
        return dirs, w

    def coupling_stats(self):
        n, b = i64(), i64()
        check(load().bhs_plan_coupling_stats(self.handle, C.byref(n), C.byref(b)), "bhs_plan_coupling_stats")
        return n.value, b.value


_plans: dict = {}
_MAX_PLANS = 48


def get_plan(d: int, n_end: int, tree: int = TREE_CHAIN) -> Plan:
    import torch

    key = (torch.cuda.current_device(), d, n_end) if tree == TREE_CHAIN else (torch.cuda.current_device(), d, n_end, tree)
    p = _plans.get(key)
    if p is None:
        # plans own device tables (coupling coefficients: ~300 MB at 3-D n_end = 39): keep the most recent _MAX_PLANS
        while len(_plans) >= _MAX_PLANS:
            _plans.pop(next(iter(_plans)))
        p = Plan(d, n_end, tree)
        _plans[key] = p
    else:
        _plans[key] = _plans.pop(key)
    return p
