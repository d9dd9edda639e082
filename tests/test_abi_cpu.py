"""No-GPU checks of the boundary: the C-ABI library loads and exports every symbol
include/acro_b200.h declares, handles that need no device can be created, and the
product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
from conftest import ROOT

import acrobotics_b200 as ab
from acrobotics_b200 import _native, device


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "acro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acro_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = _declared_symbols()
    assert len(names) >= 24
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/acro_b200.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_native.SIGNATURES) == set(names)
    assert lib.acro_abi_version() == 1


def test_struct_layout_matches_header():
    # sizes computed from the header's field list (natural alignment)
    assert ctypes.sizeof(_native.AcroBox) == 15 * 8
    assert ctypes.sizeof(_native.AcroRobotBox) == 8 + 15 * 8
    expected = 4 + 8 * 4 + 4 + 5 * 8 * 8 + 12 * 8 + 8 + 12 * 8 + 8 + 24 * 128 + 160 * 8
    assert ctypes.sizeof(_native.AcroRobotDesc) == expected


def test_robot_handle_and_error_convention():
    lib = _native.load()
    d = ab.Kuka()._describe()
    h = ctypes.c_void_p()
    assert lib.acro_robot_create(ctypes.byref(d), ctypes.byref(h)) == 0 and h.value
    lib.acro_robot_destroy(h)
    d.ndof = 9
    assert lib.acro_robot_create(ctypes.byref(d), ctypes.byref(h)) == -2  # ACRO_E_UNSUPPORTED
    assert b"ndof" in lib.acro_last_error()
    d = ab.Kuka()._describe()
    d.joint_type[2] = 7
    assert lib.acro_robot_create(ctypes.byref(d), ctypes.byref(h)) == -1
    d = ab.Kuka()._describe()
    d.self_pairs[0][1] = 99
    assert lib.acro_robot_create(ctypes.byref(d), ctypes.byref(h)) == -1
    # NULL / negative arguments never crash
    assert lib.acro_fk_f64(None, None, 10, 0, None, 0, None) == -1
    assert lib.acro_fk_cc_f64(None, None, None, -1, 0, None, None, 0.0, None) == -1
    assert lib.acro_launch_count() >= 0


def test_no_gpu_means_error_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(device.NoGpuError):
        ab.Kuka().fk([0.0] * 6)
    with pytest.raises(device.NoGpuError):
        ab.Kuka().is_in_collision([0.0] * 6, ab.Scene([ab.Box(1, 1, 1)], [np.eye(4)]))
    with pytest.raises(device.NoGpuError):
        ab.Box(1, 1, 1).is_in_collision(np.eye(4), ab.Box(1, 1, 1), np.eye(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "acrobotics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/acro_b200.h compiles as C99 and examples/c_abi_smoke.c (no Python, no torch)
    builds against the library; it is run on the GPU box by tests/test_gpu_parity.py."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    header = os.path.join(ROOT, "include", "acro_b200.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", header], check=True)
    cuda = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("CUDA toolkit headers not found")
    _native.load()
    exe = tmp_path / "c_abi_smoke"
    subprocess.run(
        [gcc, "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", f"{cuda}/include",
         os.path.join(ROOT, "examples", "c_abi_smoke.c"), "-L", os.path.join(ROOT, "acrobotics_b200"),
         "-lacro_b200", "-L", f"{cuda}/lib64", "-lcudart", "-lm", "-o", str(exe)],
        check=True)
    assert exe.exists()
