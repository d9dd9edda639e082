"""Batched closed-form inverse kinematics (GPU).

Module-level ``ik`` helpers with the reference's call shapes:
``arm_2.ik(T, links)`` (``inverse_kinematics/arm_2.py:5-104``),
``spherical_wrist.ik(T, tf_base)`` (``spherical_wrist.py:5-30``),
``anthropomorphic_arm.ik(T, links)`` (``anthropomorphic_arm.py:5-55``), plus
``*_batch`` versions over N poses.  ``planar_arm.ik`` (a 2-D toy that depends on
acrolib) is out of scope.
"""
import numpy as np
import torch

from . import _native, device
from .model import IKResult


def _positions(p):
    t, host = device.to_device(p, torch.float64, ndim=2)
    return t.reshape(-1, 3), host


def arm2_batch(p, a1, a2, a3):
    """p (N,3) -> (sol (N,4,3), valid (N,4) bool); slots [pos_a,pos_b,neg_a,neg_b]."""
    pd, host = _positions(p)
    n = pd.shape[0]
    sol = torch.empty((n, 4, 3), dtype=torch.float64, device="cuda")
    bits = torch.empty(n, dtype=torch.uint8, device="cuda")
    _native.check(
        _native.load().acro_ik_arm2_f64(
            float(a1), float(a2), float(a3), pd.data_ptr(), n, sol.data_ptr(), bits.data_ptr(),
            device.current_stream_ptr(),
        ),
        "acro_ik_arm2_f64",
    )
    return device.to_host(sol, host), device.to_host(_bits_to_bool(bits, 4), host)


def anthropomorphic_batch(p, l2, l3):
    """p (N,3) -> (sol (N,4,3), reachable (N,) bool)."""
    pd, host = _positions(p)
    n = pd.shape[0]
    sol = torch.empty((n, 4, 3), dtype=torch.float64, device="cuda")
    bits = torch.empty(n, dtype=torch.uint8, device="cuda")
    _native.check(
        _native.load().acro_ik_anthropomorphic_f64(
            float(l2), float(l3), pd.data_ptr(), n, sol.data_ptr(), bits.data_ptr(),
            device.current_stream_ptr(),
        ),
        "acro_ik_anthropomorphic_f64",
    )
    return device.to_host(sol, host), device.to_host(bits != 0, host)


def wrist_batch(R, R_base):
    """R (N,3,3), R_base (3,3) -> (N,2,3)."""
    Rd, host = device.to_device(R, torch.float64, ndim=3)
    Rd = Rd.reshape(-1, 3, 3)
    n = Rd.shape[0]
    sol = torch.empty((n, 2, 3), dtype=torch.float64, device="cuda")
    rb = (_native.C.c_double * 9)(*np.asarray(R_base, dtype=np.float64).reshape(9).tolist())
    _native.check(
        _native.load().acro_ik_wrist_f64(
            Rd.data_ptr(), rb, n, sol.data_ptr(), device.current_stream_ptr()
        ),
        "acro_ik_wrist_f64",
    )
    return device.to_host(sol, host)


def _bits_to_bool(bits, nbits):
    shifts = torch.arange(nbits, device=bits.device, dtype=torch.uint8)
    return ((bits.unsqueeze(1) >> shifts) & 1).to(torch.bool)


# -- scalar forms with the reference's signatures --------------------------------
class arm_2:
    @staticmethod
    def ik(T, links) -> IKResult:
        sol, valid = arm2_batch(np.asarray(T)[:3, 3].reshape(1, 3), links[0].dh.a, links[1].dh.a,
                                links[2].dh.a)
        q = [list(s) for s, ok in zip(sol[0], valid[0]) if ok]
        return IKResult(True, q) if q else IKResult(False)


class spherical_wrist:
    @staticmethod
    def ik(T, tf_base) -> IKResult:
        sol = wrist_batch(np.asarray(T)[:3, :3].reshape(1, 3, 3), np.asarray(tf_base)[:3, :3])
        return IKResult(True, sol[0])


class anthropomorphic_arm:
    @staticmethod
    def ik(T, links) -> IKResult:
        sol, ok = anthropomorphic_batch(np.asarray(T)[:3, 3].reshape(1, 3), links[1].dh.a,
                                        links[2].dh.a)
        return IKResult(True, sol[0]) if ok[0] else IKResult(False)
