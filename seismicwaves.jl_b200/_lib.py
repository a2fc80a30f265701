"""ctypes binding of libswb200.so (the C ABI declared in include/swb200.h).

There is no CPU fallback: importing this module fails loudly if the shared library has not been
built (run `python -c "import __graft_entry__ as g; g.build()"` or `make -C seismicwaves.jl_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWB200_LIB") or os.path.join(_HERE, "libswb200.so")  # SWB200_LIB: a differently built library (experiments)

SWB_F32, SWB_F64 = 0, 1
SWB_ACOU_CD, SWB_ACOU_VD, SWB_ELA_ISO = 1, 2, 3
SWB_FLAG_FAST_F32, SWB_FLAG_NO_FUSION, SWB_FLAG_NO_GRAPH = 1, 2, 4


class SwbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libswb200 error {code}: {msg}")
        self.code = code


class swb_cpml_axis(C.Structure):
    _fields_ = [("a", C.c_void_p), ("a_h", C.c_void_p), ("b", C.c_void_p), ("b_h", C.c_void_p)]


class swb_points(C.Structure):
    _fields_ = [("n", C.c_int64), ("pos", C.c_void_p), ("tf", C.c_void_p), ("nt", C.c_int64)]


class swb_acou_cd_step_args(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("ndim", C.c_int32), ("halo", C.c_int32), ("flags", C.c_int32),
        ("n", C.c_int64 * 3), ("spacing", C.c_double * 3),
        ("pold", C.c_void_p), ("pcur", C.c_void_p), ("pnew", C.c_void_p), ("fact", C.c_void_p),
        ("psi", C.c_void_p * 3), ("xi", C.c_void_p * 3),
        ("cpml", swb_cpml_axis * 3),
        ("src", swb_points), ("rec", swb_points),
        ("it", C.c_int64), ("stream", C.c_void_p),
    ]


class swb_acou_vd_step_args(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("halo", C.c_int32), ("flags", C.c_int32), ("_pad", C.c_int32),
        ("n", C.c_int64 * 2), ("spacing", C.c_double * 2),
        ("pcur", C.c_void_p), ("vcur", C.c_void_p * 2),
        ("fact_m0", C.c_void_p), ("fact_m1_stag", C.c_void_p * 2),
        ("psi", C.c_void_p * 2), ("xi", C.c_void_p * 2),
        ("cpml", swb_cpml_axis * 2),
        ("src", swb_points), ("rec", swb_points),
        ("it", C.c_int64), ("stream", C.c_void_p),
    ]


class swb_sinc_points(C.Structure):
    _fields_ = [("n", C.c_int64), ("off", C.c_void_p), ("ij", C.c_void_p), ("coef", C.c_void_p)]


swb_sinc_points_host = swb_sinc_points  # same layout, host pointers


class swb_ela_step_args(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("halo", C.c_int32), ("flags", C.c_int32), ("freetop", C.c_int32),
        ("n", C.c_int64 * 2), ("spacing", C.c_double * 2), ("dt", C.c_double),
        ("uold", C.c_void_p * 2), ("ucur", C.c_void_p * 2), ("unew", C.c_void_p * 2),
        ("sigma", C.c_void_p * 3),
        ("lambda_", C.c_void_p), ("mu", C.c_void_p), ("rho_ihalf", C.c_void_p), ("rho_jhalf", C.c_void_p), ("mu_ihalf_jhalf", C.c_void_p),
        ("psi_dsdx", C.c_void_p * 2), ("psi_dsdz", C.c_void_p * 2), ("psi_dudx", C.c_void_p * 2), ("psi_dudz", C.c_void_p * 2),
        ("cpml", swb_cpml_axis * 2),
        ("src_kind", C.c_int32), ("_pad", C.c_int32),
        ("src_pts", swb_sinc_points * 2), ("srctf", C.c_void_p), ("nt_tf", C.c_int64),
        ("Mxx", C.c_void_p), ("Mzz", C.c_void_p), ("Mxz", C.c_void_p),
        ("rec_pts", swb_sinc_points * 2), ("traces", C.c_void_p), ("nt_tr", C.c_int64),
        ("it", C.c_int64), ("stream", C.c_void_p),
    ]


class swb_ela_correlate_args(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("flags", C.c_int32), ("freetop", C.c_int32), ("_pad", C.c_int32),
        ("n", C.c_int64 * 2), ("spacing", C.c_double * 2), ("dt", C.c_double),
        ("adjucur", C.c_void_p * 2),
        ("u_itm2", C.c_void_p * 2), ("u_itm1", C.c_void_p * 2), ("u_it", C.c_void_p * 2),
        ("lambda_", C.c_void_p), ("mu", C.c_void_p),
        ("grad_rho_ihalf", C.c_void_p), ("grad_rho_jhalf", C.c_void_p), ("grad_lambda", C.c_void_p), ("grad_mu", C.c_void_p),
        ("grad_mu_ihalf_jhalf", C.c_void_p),
        ("stream", C.c_void_p),
    ]


class swb_sim_desc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dtype", C.c_int32), ("ndim", C.c_int32), ("device", C.c_int32),
        ("n", C.c_int64 * 3), ("spacing", C.c_double * 3), ("dt", C.c_double), ("nt", C.c_int64),
        ("halo", C.c_int32), ("freetop", C.c_int32), ("gradient", C.c_int32), ("check_freq", C.c_int32),
        ("flags", C.c_int32), ("_pad", C.c_int32),
    ]


class swb_l2_spec(C.Structure):
    _fields_ = [("observed", C.c_void_p), ("invcov_diag", C.c_void_p), ("mask", C.c_void_p)]


class swb_slab_handle(C.Structure):
    _fields_ = [("ipc", C.c_uint8 * 192), ("raw", C.c_uint64 * 3), ("nz", C.c_int64), ("plane_elems", C.c_int64), ("device", C.c_int32), ("pid", C.c_int32)]


# every symbol include/swb200.h declares (tests check the header against this list and the .so against both)
EXPORTS = [
    "swb_last_error", "swb_abi_version", "swb_abi_layout", "swb_device_count", "swb_launch_count", "swb_diag_cd_partition",
    "swb_set_device", "swb_malloc", "swb_free", "swb_memcpy_h2d", "swb_memcpy_d2h", "swb_memcpy_d2d", "swb_fill", "swb_synchronize",
    "swb_acou_cd_forward_onestep", "swb_acou_cd_adjoint_onestep", "swb_acou_cd_correlate_gradient", "swb_prescale_residuals",
    "swb_acou_vd_forward_onestep", "swb_acou_vd_adjoint_onestep", "swb_acou_vd_correlate_gradient_m0", "swb_acou_vd_correlate_gradient_m1",
    "swb_ela_forward_onestep", "swb_ela_adjoint_onestep", "swb_ela_correlate_gradients",
    "swb_sim_create", "swb_sim_destroy", "swb_sim_device_bytes", "swb_sim_set_material", "swb_sim_set_material_device", "swb_sim_set_cpml",
    "swb_sim_bind_scalar_shot", "swb_sim_bind_elastic_shot", "swb_sim_forward", "swb_sim_get_snapshot",
    "swb_sim_gradient_forward", "swb_sim_gradient_adjoint", "swb_sim_gradient_l2", "swb_sim_gradient_l2_ex", "swb_sim_get_raw_gradient",
    "swb_sim_accumulate_gradient", "swb_sim_zero_total_gradient", "swb_sim_total_gradient_ptr", "swb_sim_get_total_gradient",
    "swb_sim_cell_updates", "swb_sim_get_field", "swb_sim_stream", "swb_sim_kernel_timing", "swb_sim_kernel_timing_class",
    "swb_comm_unique_id", "swb_comm_create", "swb_comm_destroy", "swb_comm_allreduce_sum", "swb_sim_allreduce_total_gradient", "swb_sim_set_slab",
    "swb_sim_slab_export", "swb_sim_slab_connect",
]

_lib = None


def load() -> C.CDLL:
    """Load libswb200.so; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "seismicwaves.jl_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.swb_last_error.restype = C.c_char_p
    lib.swb_abi_version.restype = C.c_int32
    lib.swb_abi_layout.restype = C.c_char_p
    lib.swb_device_count.restype = C.c_int32
    lib.swb_launch_count.restype = C.c_int64
    lib.swb_sim_device_bytes.restype = C.c_int64
    lib.swb_sim_device_bytes.argtypes = [C.c_void_p]
    lib.swb_sim_cell_updates.restype = C.c_int64
    lib.swb_sim_cell_updates.argtypes = [C.c_void_p]
    vp, i32, i64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t
    sigs = {
        "swb_diag_cd_partition": [i32, i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(i64)],
        "swb_set_device": [i32],
        "swb_malloc": [C.POINTER(vp), sz],
        "swb_free": [vp],
        "swb_memcpy_h2d": [vp, vp, sz],
        "swb_memcpy_d2h": [vp, vp, sz],
        "swb_memcpy_d2d": [vp, vp, sz],
        "swb_fill": [vp, i32, dbl, sz, sz],
        "swb_synchronize": [],
        "swb_acou_cd_forward_onestep": [C.POINTER(swb_acou_cd_step_args)],
        "swb_acou_cd_adjoint_onestep": [C.POINTER(swb_acou_cd_step_args)],
        "swb_acou_cd_correlate_gradient": [i32, i32, sz, vp, vp, vp, vp, vp, dbl, vp],
        "swb_prescale_residuals": [i32, i32, C.POINTER(i64), vp, i64, i64, vp, vp, vp],
        "swb_acou_vd_forward_onestep": [C.POINTER(swb_acou_vd_step_args)],
        "swb_acou_vd_adjoint_onestep": [C.POINTER(swb_acou_vd_step_args)],
        "swb_acou_vd_correlate_gradient_m0": [i32, i32, sz, vp, vp, vp, vp, dbl, vp],
        "swb_acou_vd_correlate_gradient_m1": [i32, i32, C.POINTER(i64), C.POINTER(dbl), C.POINTER(vp), C.POINTER(vp), vp, vp],
        "swb_ela_forward_onestep": [C.POINTER(swb_ela_step_args)],
        "swb_ela_adjoint_onestep": [C.POINTER(swb_ela_step_args)],
        "swb_ela_correlate_gradients": [C.POINTER(swb_ela_correlate_args)],
        "swb_sim_create": [C.POINTER(swb_sim_desc), C.POINTER(vp)],
        "swb_sim_destroy": [vp],
        "swb_sim_set_material": [vp, i32, C.POINTER(vp), i32],
        "swb_sim_set_material_device": [vp, i32, C.POINTER(vp), i32],
        "swb_sim_set_cpml": [vp, i32, vp, vp, vp, vp],
        "swb_sim_bind_scalar_shot": [vp, i64, vp, vp, i64, vp],
        "swb_sim_bind_elastic_shot": [vp, i32, C.POINTER(swb_sinc_points), vp, vp, vp, vp, C.POINTER(swb_sinc_points)],
        "swb_sim_forward": [vp, vp, i32],
        "swb_sim_get_snapshot": [vp, i64, i32, vp],
        "swb_sim_gradient_forward": [vp, vp],
        "swb_sim_gradient_adjoint": [vp, vp],
        "swb_sim_gradient_l2": [vp, vp, vp, C.POINTER(dbl)],
        "swb_sim_gradient_l2_ex": [vp, C.POINTER(swb_l2_spec), vp, C.POINTER(dbl)],
        "swb_sim_get_raw_gradient": [vp, i32, vp],
        "swb_sim_accumulate_gradient": [vp, i64, vp, i32, i64, vp, i32],
        "swb_sim_zero_total_gradient": [vp],
        "swb_sim_total_gradient_ptr": [vp, i32, C.POINTER(vp), C.POINTER(sz)],
        "swb_sim_get_total_gradient": [vp, i32, vp],
        "swb_sim_get_field": [vp, C.c_char_p, vp, sz],
        "swb_sim_stream": [vp, C.POINTER(vp)],
        "swb_sim_kernel_timing": [vp, i32, C.POINTER(dbl), C.POINTER(i64)],
        "swb_sim_kernel_timing_class": [vp, i32, C.POINTER(dbl), C.POINTER(i64)],
        "swb_comm_unique_id": [vp],
        "swb_comm_create": [vp, i32, i32, i32, C.POINTER(vp)],
        "swb_comm_destroy": [vp],
        "swb_comm_allreduce_sum": [vp, vp, sz, i32, vp],
        "swb_sim_allreduce_total_gradient": [vp, vp],
        "swb_sim_set_slab": [vp, vp, i32, i32],
        "swb_sim_slab_export": [vp, C.POINTER(swb_slab_handle)],
        "swb_sim_slab_connect": [vp, C.POINTER(swb_slab_handle), C.POINTER(swb_slab_handle)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int32
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise SwbError(status, load().swb_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(load().swb_device_count())


def require_device() -> None:
    if device_count() < 1:
        raise RuntimeError("seismicwaves.jl_b200 needs an sm_100 (B200) device; there is no CPU fallback")
