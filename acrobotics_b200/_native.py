"""ctypes binding of ``libacro_b200.so`` (C ABI: ``include/acro_b200.h``).

There is no CPU fallback: if the library is missing and cannot be built, or a
call fails, this module raises.  PyTorch is used by the callers only for device
memory and streams; nothing here touches torch.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ACRO_B200_LIB selects an alternative build of the same library (tuning variants)
LIB_PATH = os.environ.get("ACRO_B200_LIB") or os.path.join(_HERE, "libacro_b200.so")

MAX_DOF = 8
MAX_ROBOT_BOXES = 24
MAX_SELF_PAIRS = 160
MAX_SCENE_BOXES = 64
CARRIER_BASE = -1
CARRIER_TOOL = -2
Q_AOS = 0
Q_SOA = 1


class AcroBox(C.Structure):
    _fields_ = [("half", C.c_double * 3), ("tf", C.c_double * 12)]


class AcroRobotBox(C.Structure):
    _fields_ = [
        ("carrier", C.c_int32),
        ("reserved", C.c_int32),
        ("half", C.c_double * 3),
        ("tf", C.c_double * 12),
    ]


class AcroRobotDesc(C.Structure):
    _fields_ = [
        ("ndof", C.c_int32),
        ("joint_type", C.c_int32 * MAX_DOF),
        ("a", C.c_double * MAX_DOF),
        ("d", C.c_double * MAX_DOF),
        ("theta", C.c_double * MAX_DOF),
        ("cos_alpha", C.c_double * MAX_DOF),
        ("sin_alpha", C.c_double * MAX_DOF),
        ("tf_base", C.c_double * 12),
        ("has_tool_tip", C.c_int32),
        ("check_self", C.c_int32),
        ("tf_tool_tip", C.c_double * 12),
        ("n_boxes", C.c_int32),
        ("n_self_pairs", C.c_int32),
        ("boxes", AcroRobotBox * MAX_ROBOT_BOXES),
        ("self_pairs", (C.c_int32 * 2) * MAX_SELF_PAIRS),
    ]


class NativeError(RuntimeError):
    """A libacro_b200 call returned a non-zero status."""


# every symbol include/acro_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_D = C.c_double
SIGNATURES = {
    "acro_abi_version": (C.c_int, []),
    "acro_last_error": (C.c_char_p, []),
    "acro_device_count": (C.c_int, []),
    "acro_launch_count": (_I64, []),
    "acro_set_option": (C.c_int, [C.c_char_p, C.c_char_p]),
    "acro_robot_create": (C.c_int, [C.POINTER(AcroRobotDesc), C.POINTER(_P)]),
    "acro_robot_destroy": (None, [_P]),
    "acro_scene_create": (C.c_int, [C.POINTER(AcroBox), _I32, C.POINTER(_P)]),
    "acro_scene_destroy": (None, [_P]),
    "acro_scene_prepare": (C.c_int, [_P, _P]),
    "acro_collide_boxes_f64": (C.c_int, [_P, _P, _I64, _P, _P]),
    "acro_fk_f64": (C.c_int, [_P, _P, _I64, _I32, _P, _I32, _P]),
    "acro_fk_f32": (C.c_int, [_P, _P, _I64, _I32, _P, _I32, _P]),
    "acro_fk_rpy_f64": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "acro_fk_cc_f64": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _D, _P]),
    "acro_fk_cc_workspace_bytes": (_I64, [_I64]),
    "acro_fk_cc_f64_ws": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _D, _P, _I64, _P]),
    "acro_fk_cc_f32": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    "acro_fk_cc_f64_host": (C.c_int, [_P, _P, _P, _I64, _P, _I64]),
    "acro_fk_cc_f32_host": (C.c_int, [_P, _P, _P, _I64, _P, _I64]),
    "acro_jac_pos_f64": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "acro_jac_rpy_f64": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "acro_ik_kuka_f64": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "acro_ik_kuka_cc_f64": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P]),
    "acro_joint_solutions_csr_f64": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _P, _P, _I64, _P, _P]),
    "acro_shortest_path_f64": (C.c_int, [_P, _I32, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "acro_ik_arm2_f64": (C.c_int, [_D, _D, _D, _P, _I64, _P, _P, _P]),
    "acro_ik_anthropomorphic_f64": (C.c_int, [_D, _D, _P, _I64, _P, _P, _P]),
    "acro_ik_wrist_f64": (C.c_int, [_P, C.POINTER(_D), _I64, _P, _P]),
    "acro_path_cc_f64": (C.c_int, [_P, _P, _P, _P, _I64, _D, _P, _P]),
    "acro_swept_boxes_f64": (C.c_int, [_P, _P, _P, _I64, C.c_int32, _P, _P]),
    "acro_swept_cc_f64": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int32, _P, _P, _D, _P]),
    "acro_filter_tolerance": (C.c_int, [_P, _P, C.POINTER(_D)]),
    "acro_measure_filter_error": (C.c_int, [_P, _P, _P, _I64, _P, _P]),
    "acro_measure_fma_peak": (C.c_int, [_I32, _I32, C.POINTER(_D), C.POINTER(_D)]),
    "acro_math_probe_f64": (C.c_int, [_I32, _P, _P, _P, _P, _I64, _P]),
}

_lib = None


def load():
    """Load (building it first if needed) the CUDA library; raise if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build

        _build.build()
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} is missing and could not be built; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.acro_abi_version() != 1:
        raise NativeError("libacro_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().acro_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'libacro_b200'} failed with status {rc}: {msg}")


def launch_count():
    return int(load().acro_launch_count())


def set_option(name, value):
    """Process-wide tuning knob (see ``acro_set_option`` in include/acro_b200.h)."""
    check(load().acro_set_option(str(name).encode(), str(value).encode()), "acro_set_option")
