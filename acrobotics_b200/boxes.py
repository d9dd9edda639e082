"""``Box`` and ``Scene``: the collision geometry of the hot path.

Mirrors the reference interface of ``src/acrobotics/shapes.py`` (``Box``:
``:98-172``, ``Shape.is_in_collision``: ``:50-58``) and
``src/acrobotics/geometry.py`` (``Scene``: ``:7-66``).  The Box-Box predicate
(python-fcl ``fcl.collide`` in the reference) runs on the GPU through
``acro_collide_boxes_f64``; a scene handed to ``Robot.is_in_collision`` is
flattened once into an ``AcroScene`` device handle.

Only boxes are supported (the reference's ``Cylinder`` is used by one tool
model and is out of scope).
"""
from collections import namedtuple

import numpy as np
import torch

from . import _native, device
from .transforms import pack_3x4

Polyhedron = namedtuple("Polyhedron", ["A", "b"])

# corner signs in the reference's vertex order (shapes.py:116-131)
_CORNERS = np.array(
    [
        [-1, 1, 1],
        [-1, 1, -1],
        [-1, -1, 1],
        [-1, -1, -1],
        [1, 1, 1],
        [1, 1, -1],
        [1, -1, 1],
        [1, -1, -1],
    ],
    dtype=np.float64,
)
# vertex index pairs of the 12 edges (shapes.py:133-150)
_EDGES = [(0, 1), (1, 3), (3, 2), (2, 0), (4, 5), (5, 7), (7, 6), (6, 4), (0, 4), (1, 5), (3, 7), (2, 6)]
_NORMALS = np.array(
    [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64
)


def transform_vector(tf, vector):
    return np.asarray(tf)[:3, :3] @ np.asarray(vector, dtype=np.float64) + np.asarray(tf)[:3, 3]


def _pack_boxes(shapes, tfs):
    """-> (n, 15) float64 rows laid out like the ABI's ``AcroBox``."""
    out = np.empty((len(shapes), 15), dtype=np.float64)
    for k, (s, tf) in enumerate(zip(shapes, tfs)):
        if not isinstance(s, Box):
            raise NotImplementedError(f"only Box shapes are supported, got {type(s).__name__}")
        out[k, :3] = s.half_extents
        out[k, 3:] = pack_3x4(tf)
    return out


def collide_box_pairs(boxes_a, boxes_b):
    """Row-wise Box-Box test of two (n, 15) packed arrays -> bool (n,) numpy."""
    n = boxes_a.shape[0]
    if n == 0:
        return np.zeros(0, dtype=bool)
    device.require_cuda()
    lib = _native.load()
    a = torch.from_numpy(np.ascontiguousarray(boxes_a)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(boxes_b)).cuda()
    words = torch.empty(device.mask_words(n), dtype=torch.int32, device="cuda")
    _native.check(
        lib.acro_collide_boxes_f64(
            a.data_ptr(), b.data_ptr(), n, words.data_ptr(), device.current_stream_ptr()
        ),
        "acro_collide_boxes_f64",
    )
    return device.unpack_mask(words, n, host=True)


def sweep_box_pairs(boxes_a_start, boxes_a_goal, boxes_b, n_samples=0):
    """Row-wise swept test (``fcl.continuousCollide`` with CCDM_LINEAR, shapes.py:60-75): box a
    moves from its start to its goal pose, box b is fixed; (n, 15) packed arrays -> bool (n,)."""
    n = boxes_a_start.shape[0]
    if n == 0:
        return np.zeros(0, dtype=bool)
    device.require_cuda()
    lib = _native.load()
    a0 = torch.from_numpy(np.ascontiguousarray(boxes_a_start)).cuda()
    a1 = torch.from_numpy(np.ascontiguousarray(boxes_a_goal)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(boxes_b)).cuda()
    words = torch.empty(device.mask_words(n), dtype=torch.int32, device="cuda")
    _native.check(
        lib.acro_swept_boxes_f64(a0.data_ptr(), a1.data_ptr(), b.data_ptr(), n, int(n_samples),
                                 words.data_ptr(), device.current_stream_ptr()),
        "acro_swept_boxes_f64",
    )
    return device.unpack_mask(words, n, host=True)


class Box:
    """Axis-aligned box in its own frame with side lengths ``dx, dy, dz``; it has
    no pose of its own (a transform is supplied with every query)."""

    num_edges = 12

    def __init__(self, dx, dy, dz):
        self.dx = dx
        self.dy = dy
        self.dz = dz

    @property
    def half_extents(self):
        return 0.5 * np.array([self.dx, self.dy, self.dz], dtype=np.float64)

    def is_in_collision(self, tf, other, tf_other):
        """FCL-equivalent Box-Box test; returns the contact count 0/1 like
        ``fcl.collide`` does (shapes.py:58)."""
        a = _pack_boxes([self], [tf])
        b = _pack_boxes([other], [tf_other])
        return int(collide_box_pairs(a, b)[0])

    def is_path_in_collision(self, tf, tf_target, other, tf_other):
        """``fcl.continuousCollide`` of this box moving from ``tf`` to ``tf_target`` against
        the fixed ``other`` (shapes.py:60-75); stateless (the reference's result object
        accumulates, see ``Robot.is_path_in_collision_batch``)."""
        return bool(sweep_box_pairs(_pack_boxes([self], [tf]), _pack_boxes([self], [tf_target]),
                                    _pack_boxes([other], [tf_other]))[0])

    # -- plain geometry (host side; plotting / optimisation helpers) ----------
    def get_vertices(self, tf):
        tf = np.asarray(tf, dtype=np.float64)
        return (_CORNERS * self.half_extents) @ tf[:3, :3].T + tf[:3, 3]

    def get_edges(self, tf):
        v = self.get_vertices(tf)
        return np.array([np.hstack((v[i], v[j])) for i, j in _EDGES])

    def get_normals(self, tf):
        return _NORMALS @ np.asarray(tf, dtype=np.float64)[:3, :3].T

    def get_polyhedron(self, tf):
        """``A x <= b`` description of the box."""
        A = self.get_normals(tf)
        b = np.repeat(self.half_extents, 2) + A @ np.asarray(tf, dtype=np.float64)[:3, 3]
        return Polyhedron(A, b)


class Scene:
    """A soup of shapes with their transforms in a common frame."""

    def __init__(self, shapes, tf_shapes):
        self.s = shapes
        self.tf_s = tf_shapes
        self._handle_key = None
        self._handle = None

    @property
    def shapes(self):
        return self.s

    def add(self, shapes, tf_shapes):
        self.s.extend(shapes)
        self.tf_s.extend(tf_shapes)

    def translate(self, x, y, z):
        for tf in self.tf_s:
            tf[:3, 3] += np.array([x, y, z])

    def get_polyhedrons(self):
        return [Polyhedron(*s.get_polyhedron(tf)) for s, tf in zip(self.s, self.tf_s)]

    def packed(self, tf=None):
        """(n, 15) packed boxes, optionally moved by ``tf`` (geometry.py:56-59)."""
        tfs = self.tf_s if tf is None else [np.asarray(tf) @ t for t in self.tf_s]
        return _pack_boxes(self.s, tfs)

    def is_in_collision(self, other, tf_self=None, tf_other=None):
        """Any shape of ``self`` against any shape of ``other`` (geometry.py:51-66);
        all k*m pairs go to the GPU in one launch."""
        a = self.packed(tf_self)
        b = other.packed(tf_other)
        if len(a) == 0 or len(b) == 0:
            return False
        ia, ib = np.meshgrid(np.arange(len(a)), np.arange(len(b)), indexing="ij")
        return bool(collide_box_pairs(a[ia.ravel()], b[ib.ravel()]).any())

    def is_path_in_collision(self, tf_self, tf_target, other, tf_other=None):
        """Every shape of ``self`` as its own moving object between ``tf_self @ tf`` and
        ``tf_target @ tf`` against every shape of ``other`` (geometry.py:68-81)."""
        a0 = self.packed(tf_self)
        a1 = self.packed(tf_target)
        b = other.packed(tf_other)
        if len(a0) == 0 or len(b) == 0:
            return False
        ia, ib = np.meshgrid(np.arange(len(a0)), np.arange(len(b)), indexing="ij")
        return bool(sweep_box_pairs(a0[ia.ravel()], a1[ia.ravel()], b[ib.ravel()]).any())

    # -- device handle --------------------------------------------------------
    def native_handle(self):
        """``AcroScene*`` for the current contents on the current device (cached
        until the shapes or transforms change)."""
        device.require_cuda()
        packed = self.packed()
        key = (torch.cuda.current_device(), packed.tobytes())
        if key != self._handle_key:
            self._release()
            lib = _native.load()
            n = packed.shape[0]
            if n > _native.MAX_SCENE_BOXES:
                raise ValueError(f"scene has {n} boxes; at most {_native.MAX_SCENE_BOXES} supported")
            arr = (_native.AcroBox * max(n, 1))()
            for k in range(n):
                arr[k].half[:] = packed[k, :3].tolist()
                arr[k].tf[:] = packed[k, 3:].tolist()
            import ctypes

            h = ctypes.c_void_p()
            _native.check(lib.acro_scene_create(arr, n, ctypes.byref(h)), "acro_scene_create")
            self._handle, self._handle_key = h, key
        return self._handle

    def prepare(self, robot):
        """Build now what the first large ``robot.is_in_collision_batch(q, self)`` would
        build: the candidate grid of the filtered kernel for this (robot, scene) pair
        (``acro_scene_prepare``).  After it, batches of 4096 configurations and more take
        the filtered kernel."""
        _native.check(
            _native.load().acro_scene_prepare(self.native_handle(), robot._handle()),
            "acro_scene_prepare",
        )
        return self

    def _release(self):
        if self._handle is not None:
            try:
                _native.load().acro_scene_destroy(self._handle)
            except Exception:
                pass
        self._handle = None
        self._handle_key = None

    def __del__(self):
        self._release()
