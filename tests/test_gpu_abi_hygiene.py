"""ABI hygiene: caller-owned workspace (no allocation in the hot
call), the host pipeline is per device and usable from several threads at once, plans built
on the caller's stream, and the Python handle cache sees every value that feeds a descriptor."""
import ctypes
import threading

import numpy as np
import pytest
import torch

import acrobotics_b200 as ab
from acrobotics_b200 import _native
from acrobotics_b200.synthetic import random_configs, random_configs_device, scene16

pytestmark = pytest.mark.gpu


def test_workspace_entry_point_equals_plain_call(gpu):
    lib = _native.load()
    k, scene = ab.Kuka(), scene16()
    n = 300_000
    q = random_configs_device(n, 6, seed=11)
    want = k.is_in_collision_batch(q, scene, packed=True)
    rh, sh = k._handle(), scene.native_handle()
    nbytes = int(lib.acro_fk_cc_workspace_bytes(n))
    assert nbytes >= 4 * ((n + 31) // 32) + 8 and lib.acro_fk_cc_workspace_bytes(0) == 0
    ws = torch.empty(nbytes // 8 + 1, dtype=torch.int64, device="cuda")
    ws.fill_(-1)  # contents are undefined on entry
    out = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        _native.check(lib.acro_fk_cc_f64_ws(rh, sh, q.data_ptr(), n, 0, out.data_ptr(), None, 0.0,
                                            ws.data_ptr(), nbytes, st))
        assert torch.equal(out, want)
    # too small / misaligned workspaces are refused, a NULL one falls back to the pool
    assert lib.acro_fk_cc_f64_ws(rh, sh, q.data_ptr(), n, 0, out.data_ptr(), None, 0.0,
                                 ws.data_ptr(), 16, st) == -1
    assert lib.acro_fk_cc_f64_ws(rh, sh, q.data_ptr(), n, 0, out.data_ptr(), None, 0.0,
                                 ws.data_ptr() + 4, nbytes, st) == -1
    _native.check(lib.acro_fk_cc_f64_ws(rh, sh, q.data_ptr(), n, 0, out.data_ptr(), None, 0.0,
                                        None, 0, st))
    assert torch.equal(out, want)


def test_first_use_plan_on_side_streams(gpu):
    """A fresh (robot, scene) pair used for the first time from two non-default streams at
    once: the grid is built on the first caller's stream and the second waits on its event."""
    k = ab.Kuka()
    k.tf_base = np.array([[1, 0, 0, 0.013], [0, 1, 0, -0.02], [0, 0, 1, 0.005], [0, 0, 0, 1.0]])
    scene = scene16(seed=5)
    q = random_configs_device(200_000, 6, seed=12)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        a = k.is_in_collision_batch(q, scene, packed=True)
    with torch.cuda.stream(s2):
        b = k.is_in_collision_batch(q, scene, packed=True)
    torch.cuda.synchronize()
    _native.set_option("k2_variant", "simple")
    try:
        want = k.is_in_collision_batch(q, scene, packed=True)
    finally:
        _native.set_option("k2_variant", "filter")
    assert torch.equal(a, want) and torch.equal(b, want)


def test_host_pipeline_from_two_threads(gpu):
    """acro_fk_cc_f64_host from two host threads at once - on two devices when the box has
    them (the pipeline is per device), else on one (calls take turns)."""
    lib = _native.load()
    n = 1_500_000
    devices = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    q_np = [random_configs(n, 6, seed=20 + i) for i in range(2)]
    results, errors = [None, None], []

    def work(i):
        try:
            torch.cuda.set_device(devices[i])
            k, scene = ab.Kuka(), scene16()
            qh = torch.from_numpy(q_np[i]).pin_memory()
            mh = torch.zeros((n + 31) // 32, dtype=torch.int32).pin_memory()
            for _ in range(3):
                _native.check(lib.acro_fk_cc_f64_host(k._handle(), scene.native_handle(),
                                                      qh.data_ptr(), n, mh.data_ptr(), 1 << 18))
            results[i] = mh.numpy().copy()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    torch.cuda.set_device(0)
    k, scene = ab.Kuka(), scene16()
    for i in range(2):
        want = k.is_in_collision_batch(q_np[i], scene, packed=True)
        assert np.array_equal(results[i].view(np.uint32), want)


def test_handle_cache_sees_box_sizes_and_dh_values(gpu):
    k, scene = ab.Kuka(), scene16()
    q = random_configs(20_000, 6, seed=13)
    base = k.is_in_collision_batch(q, scene)
    # grow one link box in place: more collisions, never fewer
    box = k.links[3].geometry.s[0]
    k.links[3].geometry.s[0] = ab.Box(3 * box.dx, 3 * box.dy, 3 * box.dz)
    grown = k.is_in_collision_batch(q, scene)
    assert (grown | ~base).all() and grown.sum() > base.sum()
    k.links[3].geometry.s[0] = box
    assert np.array_equal(k.is_in_collision_batch(q, scene), base)
    # change a DH value: the poses move
    T0 = k.fk_batch(q[:100])
    link = k.links[1]
    k.links[1] = type(link)(link.dh._replace(a=link.dh.a + 0.1), link.joint_type, link.geometry)
    T1 = k.fk_batch(q[:100])
    assert np.abs(T1 - T0).max() > 0.05
    k.links[1] = link
    assert np.array_equal(k.fk_batch(q[:100]), T0)


def test_csr_rejects_misaligned_rows(gpu):
    lib = _native.load()
    k, scene = ab.Kuka(), scene16()
    T = k.fk_batch(random_configs_device(64, 6, seed=14)).contiguous()
    buf = torch.empty(64 * 8 * 6 + 1, dtype=torch.float64, device="cuda")
    q_free = torch.empty((64 * 8, 6), dtype=torch.float64, device="cuda")
    valid = torch.empty(64, dtype=torch.uint8, device="cuda")
    free = torch.empty(64, dtype=torch.uint8, device="cuda")
    off = torch.zeros(3, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.acro_joint_solutions_csr_f64(k._handle(), scene.native_handle(), T.data_ptr(), 2, 32,
                                          buf.data_ptr() + 8, valid.data_ptr(), free.data_ptr(),
                                          q_free.data_ptr(), 64 * 8, off.data_ptr(), st)
    assert rc == -1 and b"16-byte aligned" in lib.acro_last_error()
    _native.check(lib.acro_joint_solutions_csr_f64(k._handle(), scene.native_handle(), T.data_ptr(), 2, 32,
                                                   buf.data_ptr(), valid.data_ptr(), free.data_ptr(),
                                                   q_free.data_ptr(), 64 * 8, off.data_ptr(), st))


def test_host_pipeline_with_fp32_rows(gpu):
    """acro_fk_cc_f32_host: the FP32 rows are widened exactly, so the verdicts are those of the
    FP64 path at the float-valued joint angles (ragged size, several chunks)."""
    lib = _native.load()
    k, scene = ab.Kuka(), scene16()
    n = 777_777
    q32 = random_configs(n, 6, seed=31).astype(np.float32)
    qh = torch.from_numpy(q32).pin_memory()
    mh = torch.zeros((n + 31) // 32, dtype=torch.int32).pin_memory()
    _native.check(lib.acro_fk_cc_f32_host(k._handle(), scene.native_handle(), qh.data_ptr(), n,
                                          mh.data_ptr(), 1 << 17))
    want = k.is_in_collision_batch(q32.astype(np.float64), scene, packed=True)
    assert np.array_equal(mh.numpy().view(np.uint32), want)
    # and the FP64 host path on the same rows
    qh64 = torch.from_numpy(q32.astype(np.float64)).pin_memory()
    mh.zero_()
    _native.check(lib.acro_fk_cc_f64_host(k._handle(), scene.native_handle(), qh64.data_ptr(), n,
                                          mh.data_ptr(), 1 << 17))
    assert np.array_equal(mh.numpy().view(np.uint32), want)
