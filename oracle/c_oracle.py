"""Build + ctypes wrapper of oracle/c/acro_oracle.c (TEST INFRASTRUCTURE).

``fk_cc(spec, scene, q)`` evaluates the C restatement over N configurations
(optionally across processes), ``box_box`` is the single native call the
per-configuration port uses where the reference calls python-fcl."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import np_oracle as no

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "acro_oracle.c")
LIB = os.path.join(HERE, "c", "libacro_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC, "-lm"],
            check=True,
        )
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        lib.oracle_box_box.restype = C.c_int
        lib.oracle_box_box.argtypes = [dp, dp, dp, dp, dp]
        lib.oracle_fk_cc.restype = None
        lib.oracle_fk_cc.argtypes = [
            C.c_int, dp, C.POINTER(C.c_int32), dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int32),
            C.c_int, dp, C.c_int, dp, C.c_int64, C.POINTER(C.c_uint8), dp,
        ]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def box_box_fn():
    """Raw native box-box: f(side1, tf1, side2, tf2) -> 0/1 on contiguous float64 arrays."""
    lib = load()
    fn = lib.oracle_box_box

    def call(side1, tf1, side2, tf2):
        return fn(_dp(side1), _dp(tf1), _dp(side2), _dp(tf2), None)

    return call


def pack(spec: no.RobotSpec, scene):
    """RobotSpec + scene -> the flat arrays oracle_fk_cc takes."""
    boxes = np.zeros((max(len(spec.boxes), 1), 20))
    for k, b in enumerate(spec.boxes):
        boxes[k, 0] = b.group
        boxes[k, 1:4] = 2.0 * b.half
        boxes[k, 4:] = b.tf.reshape(16)
    groups = [b.group for b in spec.boxes]
    pairs = [
        (b1, b2)
        for gi, gj in spec.self_pairs
        for b1 in range(len(groups)) if groups[b1] == gi
        for b2 in range(len(groups)) if groups[b2] == gj
    ]
    pairs = np.array(pairs if pairs else [(0, 0)], dtype=np.int32)
    if scene is None:
        sc = np.zeros((1, 19))
        ns = 0
    else:
        sh, sR, sT = scene if isinstance(scene, tuple) else no.scene_arrays(scene)
        ns = len(sh)
        sc = np.zeros((max(ns, 1), 19))
        for k in range(ns):
            tf = np.eye(4)
            tf[:3, :3], tf[:3, 3] = sR[k], sT[k]
            sc[k, :3] = 2.0 * sh[k]
            sc[k, 3:] = tf.reshape(16)
    return dict(
        ndof=spec.ndof, dh=np.ascontiguousarray(spec.dh), prismatic=spec.prismatic.astype(np.int32),
        base=np.ascontiguousarray(spec.tf_base.reshape(16)), n_boxes=len(spec.boxes), boxes=boxes,
        n_pairs=len(spec.self_pairs) and len(pairs), pairs=pairs, n_scene=ns, scene=sc,
        check_self=int(spec.check_self),
    )


def fk_cc_packed(p, q, want_g=False):
    lib = load()
    q = np.ascontiguousarray(q, dtype=np.float64)
    n = q.shape[0]
    hit = np.zeros(n, dtype=np.uint8)
    g = np.zeros(n) if want_g else None
    lib.oracle_fk_cc(
        p["ndof"], _dp(p["dh"]), p["prismatic"].ctypes.data_as(C.POINTER(C.c_int32)), _dp(p["base"]),
        p["n_boxes"], _dp(p["boxes"]), p["n_pairs"], p["pairs"].ctypes.data_as(C.POINTER(C.c_int32)),
        p["n_scene"], _dp(p["scene"]), p["check_self"], _dp(q), n,
        hit.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(g) if want_g else None,
    )
    return (hit.astype(bool), g) if want_g else hit.astype(bool)


def fk_cc(spec, scene, q, want_g=False):
    return fk_cc_packed(pack(spec, scene), q, want_g)
