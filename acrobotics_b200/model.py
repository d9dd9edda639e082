"""``Robot``: serial DH chain + box geometry, evaluated on the GPU.

Keeps the reference API of ``src/acrobotics/robot.py`` - ``fk`` (``:59-69``),
``fk_rpy`` (``:71-80``), ``fk_all_links`` (``:82-91``), ``jacobian_position`` /
``jacobian_rpy`` (``:157-171``), ``is_in_self_collision`` (``:239-242``),
``is_in_collision`` (``:244-273``), ``is_path_in_collision_discrete``
(``:275-283``), ``Tool`` (``:23-33``), ``JointLimit`` (``:12``), ``IKResult``
(``:15-20``) - and adds ``*_batch`` entry points over ``(N, ndof)`` arrays.

Scalar calls are N=1 batches of the same CUDA kernels.  ``q`` may be a list,
tuple, numpy array (results come back as numpy) or a CUDA ``torch.Tensor``
(results stay on the device).  The robot is flattened into an ``AcroRobotDesc``
and turned into an immutable native handle, rebuilt only when an attribute that
feeds it (``tf_base``, ``tool``, ``collision_matrix`` ...) changes.
"""
import ctypes
from collections import namedtuple
from typing import List

import numpy as np
import torch

from . import _native, device
from .boxes import Box, Scene, collide_box_pairs
from .chain import JointType, Link
from .transforms import pack_3x4

JointLimit = namedtuple("JointLimit", ["lower", "upper"])

DEFAULT_MARGIN = 1.0e-6  # metres: the reported near-contact bucket


class IKResult:
    def __init__(self, success: bool, solutions: List[np.ndarray] = None):
        self.success = success
        if self.success:
            assert solutions is not None
            self.solutions = solutions


class Tool(Scene):
    """Shapes of a tool plus the tool-tip transform relative to the last link."""

    def __init__(self, shapes, tf_shapes, tf_tool_tip):
        super().__init__(shapes, tf_shapes)
        self.tf_tool_tip = tf_tool_tip


class _Handle:
    """Owner of one ``AcroRobot*``."""

    def __init__(self, desc):
        self.ptr = ctypes.c_void_p()
        _native.check(
            _native.load().acro_robot_create(ctypes.byref(desc), ctypes.byref(self.ptr)),
            "acro_robot_create",
        )

    def __del__(self):
        try:
            if self.ptr:
                _native.load().acro_robot_destroy(self.ptr)
        except Exception:
            pass


class RobotKinematics:
    """Kinematics of a serial chain (no collision geometry needed)."""

    def __init__(self, links, joint_limits: List[JointLimit] = None):
        self.links = links
        self.ndof = len(links)
        if joint_limits is None:
            self.joint_limits = [JointLimit(-np.pi, np.pi)] * self.ndof
        else:
            self.joint_limits = joint_limits
        self.tf_base = np.eye(4)
        self.tf_tool_tip = None
        self._handles = {}

    # ------------------------------------------------------------------
    # flattening
    # ------------------------------------------------------------------
    def _describe(self, kinematics_only=False, as_constructed=False, force_self=False):
        """Build the ABI descriptor.  ``as_constructed``: identity base, no tool
        tip (what the reference's Jacobian functions bake in, robot.py:194-195)."""
        if self.ndof > _native.MAX_DOF:
            raise ValueError(f"at most {_native.MAX_DOF} joints supported, got {self.ndof}")
        d = _native.AcroRobotDesc()
        d.ndof = self.ndof
        for i, link in enumerate(self.links):
            jt = link.joint_type
            d.joint_type[i] = 1 if jt == JointType.prismatic else 0
            d.a[i], d.d[i], d.theta[i] = link.dh.a, link.dh.d, link.dh.theta
            # same calls as the reference (link.py:48-49), so cos(pi/2) = 6.1e-17
            d.cos_alpha[i] = float(np.cos(link.dh.alpha))
            d.sin_alpha[i] = float(np.sin(link.dh.alpha))
        base = np.eye(4) if as_constructed else np.asarray(self.tf_base, dtype=np.float64)
        d.tf_base[:] = pack_3x4(base).tolist()
        tip = None if as_constructed else self.tf_tool_tip
        d.has_tool_tip = int(tip is not None)
        d.tf_tool_tip[:] = pack_3x4(np.eye(4) if tip is None else tip).tolist()
        d.check_self = 0
        d.n_boxes = 0
        d.n_self_pairs = 0
        if not kinematics_only:
            self._describe_geometry(d, force_self)
        return d

    def _describe_geometry(self, d, force_self):  # overridden by Robot
        pass

    def _signature(self):
        """Fingerprint of everything ``_describe`` reads: the descriptor (a ~5 KB ctypes
        struct) is only rebuilt when it changes.  Everything is compared by value - base and
        tool-tip transforms, the DH rows and joint types of every link - so in-place edits and
        swapped links are seen."""
        tip = self.tf_tool_tip
        return (
            np.asarray(self.tf_base, dtype=np.float64).tobytes(),
            None if tip is None else np.asarray(tip, dtype=np.float64).tobytes(),
            tuple(
                (float(l.dh.a), float(l.dh.alpha), float(l.dh.d), float(l.dh.theta),
                 getattr(l.joint_type, "value", l.joint_type))
                for l in self.links
            ),
        )

    def _handle(self, **kw):
        sig = (self._signature(), tuple(sorted(kw.items())))
        h = self._handles.get(sig)
        if h is None:
            if len(self._handles) > 16:
                self._handles.clear()
            h = self._handles[sig] = _Handle(self._describe(**kw))
        return h.ptr

    def _q(self, q, dtype=torch.float64):
        t, host = device.to_device(q, dtype, ndim=2)
        if t.shape[-1] != self.ndof:
            raise ValueError(f"expected {self.ndof} joint values per configuration, got {t.shape[-1]}")
        return t.reshape(-1, self.ndof), host

    # ------------------------------------------------------------------
    # forward kinematics
    # ------------------------------------------------------------------
    def fk_batch(self, q, all_links=False, dtype="f64"):
        """(N, ndof) -> (N, 4, 4) end-effector poses (tool tip applied), or with
        ``all_links`` -> (N, ndof, 4, 4) link frames (base applied, tip not)."""
        tdt = torch.float64 if dtype == "f64" else torch.float32
        qd, host = self._q(q, tdt)
        n = qd.shape[0]
        shape = (n, self.ndof, 4, 4) if all_links else (n, 4, 4)
        out = torch.empty(shape, dtype=tdt, device="cuda")
        lib = _native.load()
        fn = lib.acro_fk_f64 if dtype == "f64" else lib.acro_fk_f32
        _native.check(
            fn(
                self._handle(kinematics_only=True),
                qd.data_ptr(),
                n,
                _native.Q_AOS,
                out.data_ptr(),
                int(all_links),
                device.current_stream_ptr(),
            ),
            "acro_fk",
        )
        return device.to_host(out, host)

    def fk_rpy_batch(self, q):
        qd, host = self._q(q)
        n = qd.shape[0]
        out = torch.empty((n, 6), dtype=torch.float64, device="cuda")
        _native.check(
            _native.load().acro_fk_rpy_f64(
                self._handle(kinematics_only=True),
                qd.data_ptr(),
                n,
                _native.Q_AOS,
                out.data_ptr(),
                device.current_stream_ptr(),
            ),
            "acro_fk_rpy_f64",
        )
        return device.to_host(out, host)

    def fk(self, q) -> np.ndarray:
        """End-effector frame (tool frame when a tool is mounted)."""
        return self.fk_batch(np.asarray(q, dtype=np.float64).reshape(1, -1))[0]

    def fk_rpy(self, q) -> np.ndarray:
        return self.fk_rpy_batch(np.asarray(q, dtype=np.float64).reshape(1, -1))[0]

    def fk_all_links(self, q) -> List[np.ndarray]:
        """Link frames (not base or tool)."""
        T = self.fk_batch(np.asarray(q, dtype=np.float64).reshape(1, -1), all_links=True)[0]
        return [T[i] for i in range(self.ndof)]

    def ik(self, transformation_matrix) -> IKResult:
        raise NotImplementedError

    def estimate_max_extension(self):
        return sum(abs(l.dh.a) + abs(l.dh.d) for l in self.links)

    # ------------------------------------------------------------------
    # Jacobians (analytic; the reference differentiates fk_casadi with casadi)
    # ------------------------------------------------------------------
    def _jac_batch(self, q, rpy):
        qd, host = self._q(q)
        n = qd.shape[0]
        rows = 6 if rpy else 3
        out = torch.empty((n, rows, self.ndof), dtype=torch.float64, device="cuda")
        lib = _native.load()
        fn = lib.acro_jac_rpy_f64 if rpy else lib.acro_jac_pos_f64
        _native.check(
            fn(
                self._handle(kinematics_only=True, as_constructed=True),
                qd.data_ptr(),
                n,
                _native.Q_AOS,
                out.data_ptr(),
                device.current_stream_ptr(),
            ),
            "acro_jac",
        )
        return device.to_host(out, host)

    def jacobian_position_batch(self, q):
        """(N, ndof) -> (N, 3, ndof)."""
        return self._jac_batch(q, False)

    def jacobian_rpy_batch(self, q):
        """(N, ndof) -> (N, 6, ndof): Jacobian of ``fk_rpy``'s 6-vector."""
        return self._jac_batch(q, True)

    def jacobian_position(self, q):
        return self._jac_batch(np.asarray(q, dtype=np.float64).reshape(1, -1), False)[0]

    def jacobian_rpy(self, q):
        return self._jac_batch(np.asarray(q, dtype=np.float64).reshape(1, -1), True)[0]


class Robot(RobotKinematics):
    def __init__(self, links, joint_limits=None):
        super().__init__(links, joint_limits)
        self.geometry_base = None
        self.geometry_tool = None
        self.do_check_self_collision = True
        # neighbours up to two links apart are never checked (band matrix)
        idx = np.arange(self.ndof)
        self.collision_matrix = np.abs(idx[:, None] - idx[None, :]) >= 3
        self.collision_priority = list(range(self.ndof))  # kept for API compatibility
        self.cc_checks = 0

    @property
    def tool(self):
        return self.geometry_tool

    @tool.setter
    def tool(self, new_tool: Tool):
        self.tf_tool_tip = new_tool.tf_tool_tip
        self.geometry_tool = new_tool

    # ------------------------------------------------------------------
    # flattening of the geometry
    # ------------------------------------------------------------------
    def _box_groups(self):
        """[(carrier, Box, tf_offset)] in the order links, tool, base."""
        out = []
        for i, link in enumerate(self.links):
            geom = getattr(link, "geometry", None)
            if geom is not None:
                out += [(i, s, tf) for s, tf in zip(geom.s, geom.tf_s)]
        if self.geometry_tool is not None:
            out += [
                (_native.CARRIER_TOOL, s, tf)
                for s, tf in zip(self.geometry_tool.s, self.geometry_tool.tf_s)
            ]
        if self.geometry_base is not None:
            out += [
                (_native.CARRIER_BASE, s, tf)
                for s, tf in zip(self.geometry_base.s, self.geometry_base.tf_s)
            ]
        return out

    def _signature(self):
        def scene_sig(sc):
            if sc is None:
                return None
            # shapes by value (type + extents), transforms by value
            shapes = tuple(
                (type(sh).__name__,) + tuple(float(v) for v in getattr(sh, "half_extents", ()))
                for sh in sc.s
            )
            return (shapes, b"".join(np.asarray(t, dtype=np.float64).tobytes() for t in sc.tf_s))

        return super()._signature() + (
            bool(self.do_check_self_collision),
            np.asarray(self.collision_matrix, dtype=bool).tobytes(),
            scene_sig(self.geometry_tool),
            scene_sig(self.geometry_base),
            tuple(scene_sig(getattr(l, "geometry", None)) for l in self.links),
        )

    def _describe_geometry(self, d, force_self):
        boxes = self._box_groups()
        if len(boxes) > _native.MAX_ROBOT_BOXES:
            raise ValueError(f"robot has {len(boxes)} boxes; at most {_native.MAX_ROBOT_BOXES}")
        for k, (carrier, shape, tf) in enumerate(boxes):
            if not isinstance(shape, Box):
                raise NotImplementedError("only Box shapes are supported on the GPU path")
            d.boxes[k].carrier = carrier
            d.boxes[k].half[:] = shape.half_extents.tolist()
            d.boxes[k].tf[:] = pack_3x4(tf).tolist()
        d.n_boxes = len(boxes)
        d.check_self = int(bool(self.do_check_self_collision) or force_self)
        # link pairs of the collision matrix (robot.py:206-211; the reference
        # visits (i,j) and (j,i), the predicate is symmetric) + tool against
        # every link but the last (robot.py:214-220)
        cm = np.asarray(self.collision_matrix, dtype=bool)
        by_carrier = {}
        for k, (carrier, _, _) in enumerate(boxes):
            by_carrier.setdefault(carrier, []).append(k)
        group_pairs = [
            (i, j) for i in range(self.ndof) for j in range(i + 1, self.ndof) if cm[i, j] or cm[j, i]
        ]
        if self.geometry_tool is not None:
            group_pairs += [(i, _native.CARRIER_TOOL) for i in range(self.ndof - 1)]
        pairs = [
            (b1, b2)
            for gi, gj in group_pairs
            for b1 in by_carrier.get(gi, [])
            for b2 in by_carrier.get(gj, [])
        ]
        if len(pairs) > _native.MAX_SELF_PAIRS:
            raise ValueError(f"{len(pairs)} self-collision box pairs; at most {_native.MAX_SELF_PAIRS}")
        for k, (b1, b2) in enumerate(pairs):
            d.self_pairs[k][0], d.self_pairs[k][1] = b1, b2
        d.n_self_pairs = len(pairs)

    # ------------------------------------------------------------------
    # collision checking
    # ------------------------------------------------------------------
    def is_in_collision_batch(
        self, q, scene, return_near=False, margin=DEFAULT_MARGIN, packed=False, dtype="f64",
        _force_self=False, out=None,
    ):
        """(N, ndof) x Scene -> bool (N,).  ``return_near``: also a mask of the
        configurations whose decisive separation is within ``margin`` of contact.
        ``packed``: return the raw uint32 bit-mask words instead of booleans.
        ``out``: a CUDA int32 tensor of at least ``ceil(N / 32)`` words that receives the packed
        verdicts in place (implies ``packed``; used by the multi-GPU sharding, whose kernel
        writes straight into its block of the all-gather buffer)."""
        tdt = torch.float64 if dtype == "f64" else torch.float32
        qd, host = self._q(q, tdt)
        n = qd.shape[0]
        self.cc_checks += n
        if out is not None:
            if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.int32
                    and out.is_contiguous() and out.numel() >= device.mask_words(n)):
                raise ValueError("out must be a contiguous CUDA int32 tensor of >= ceil(N/32) words")
            words, packed, host = out, True, False
        else:
            words = torch.zeros(device.mask_words(n), dtype=torch.int32, device="cuda")
        near = torch.zeros_like(words) if return_near else None
        lib = _native.load()
        robot_h = self._handle(force_self=_force_self)
        scene_h = scene.native_handle() if scene is not None else None
        if dtype == "f64":
            # caller-owned scratch: with it the native hot call allocates nothing
            ws_bytes = int(lib.acro_fk_cc_workspace_bytes(n))
            ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.int64, device="cuda") if ws_bytes else None
            rc = lib.acro_fk_cc_f64_ws(
                robot_h, scene_h, qd.data_ptr(), n, _native.Q_AOS, words.data_ptr(),
                near.data_ptr() if near is not None else None, float(margin),
                ws.data_ptr() if ws is not None else None, ws_bytes if ws is not None else 0,
                device.current_stream_ptr(),
            )
        else:
            if return_near:
                raise ValueError("the near-margin mask is only produced by the FP64 kernel")
            rc = lib.acro_fk_cc_f32(
                robot_h, scene_h, qd.data_ptr(), n, _native.Q_AOS, words.data_ptr(),
                device.current_stream_ptr(),
            )
        _native.check(rc, "acro_fk_cc")
        if packed:
            res = words if not host else words.cpu().numpy().view(np.uint32)
            if return_near:
                return res, (near if not host else near.cpu().numpy().view(np.uint32))
            return res
        hit = device.unpack_mask(words, n, host)
        if return_near:
            return hit, device.unpack_mask(near, n, host)
        return hit

    def is_in_collision(self, q, collection):
        """Collision of the robot at ``q`` with the scene, or with itself."""
        if collection is None:
            # the reference crashes here (UnboundLocalError, robot.py:246-271)
            raise ValueError("is_in_collision needs a Scene; use is_in_self_collision(q)")
        q = np.asarray(q, dtype=np.float64).reshape(1, -1)
        return bool(self.is_in_collision_batch(q, collection)[0])

    def is_in_self_collision_batch(self, q):
        """Self pairs only, regardless of ``do_check_self_collision`` (robot.py:239-242)."""
        return self.is_in_collision_batch(q, None, _force_self=True)

    def is_in_self_collision(self, q):
        q = np.asarray(q, dtype=np.float64).reshape(1, -1)
        return bool(self.is_in_self_collision_batch(q)[0])

    def _check_self_collision(self, tf_links, geom_links):
        """Self collision from given link frames (robot.py:206-221)."""
        cm = np.asarray(self.collision_matrix, dtype=bool)
        a, b = [], []
        for i in range(self.ndof):
            for j in range(i + 1, self.ndof):
                if cm[i, j] or cm[j, i]:
                    pi, pj = geom_links[i].packed(tf_links[i]), geom_links[j].packed(tf_links[j])
                    for x in pi:
                        for y in pj:
                            a.append(x)
                            b.append(y)
        if self.geometry_tool is not None:
            pt = self.geometry_tool.packed(tf_links[-1])
            for tf_link, geom in zip(tf_links[:-1], geom_links[:-1]):
                for x in geom.packed(tf_link):
                    for y in pt:
                        a.append(x)
                        b.append(y)
        if not a:
            return False
        return bool(collide_box_pairs(np.array(a), np.array(b)).any())

    def _is_in_limits(self, q):
        return all(l.lower <= qi <= l.upper for qi, l in zip(q, self.joint_limits))

    # ------------------------------------------------------------------
    # discretised path
    # ------------------------------------------------------------------
    @staticmethod
    def _linear_interpolation_path(q_start, q_goal, max_q_step):
        q_start, q_goal = np.array(q_start, dtype=float), np.array(q_goal, dtype=float)
        num_steps = int(np.ceil(np.linalg.norm(q_goal - q_start) / max_q_step))
        return [(1 - s) * q_start + s * q_goal for s in np.linspace(0, 1, num_steps)]

    def is_path_in_collision_discrete_batch(self, q_start, q_goal, scene, max_q_step=0.1,
                                            packed=False, out=None):
        """(E, ndof) segment starts / goals -> bool (E,): some interpolant collides.
        ``packed`` / ``out``: as in ``is_in_collision_batch``."""
        qs, host = self._q(q_start)
        qg, _ = self._q(q_goal)
        if qs.shape != qg.shape:
            raise ValueError("q_start and q_goal must have the same shape")
        e = qs.shape[0]
        if out is not None:
            if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.int32
                    and out.is_contiguous() and out.numel() >= device.mask_words(e)):
                raise ValueError("out must be a contiguous CUDA int32 tensor of >= ceil(E/32) words")
            words, packed, host = out, True, False
        else:
            words = torch.zeros(device.mask_words(e), dtype=torch.int32, device="cuda")
        _native.check(
            _native.load().acro_path_cc_f64(
                self._handle(),
                scene.native_handle() if scene is not None else None,
                qs.data_ptr(), qg.data_ptr(), e, float(max_q_step), words.data_ptr(),
                device.current_stream_ptr(),
            ),
            "acro_path_cc_f64",
        )
        if packed:
            return words if not host else words.cpu().numpy().view(np.uint32)
        return device.unpack_mask(words, e, host)

    # ------------------------------------------------------------------
    # swept (continuous) collision
    # ------------------------------------------------------------------
    def is_path_in_collision_batch(self, q_start, q_goal, scene, return_near=False,
                                   margin=DEFAULT_MARGIN, n_samples=0, packed=False, out=None):
        """(E, ndof) starts / goals -> bool (E,): the reference's ``is_path_in_collision``
        (``robot.py:285-317``) per pair - every tool / link box as a moving object between its
        two world poses under FCL's linear interpolated motion (``shapes.py:60-75``), base
        boxes fixed, no self collision.  ``n_samples`` = 0 takes FCL's default request (10
        samples of the motion, both end poses included).  The verdict is the stateless one:
        python-fcl ORs ``is_collide`` into the result object the reference keeps per shape, so
        the reference itself keeps answering True once a shape has collided.
        ``packed`` / ``out``: as in ``is_in_collision_batch``."""
        qs, host = self._q(q_start)
        qg, _ = self._q(q_goal)
        if qs.shape != qg.shape:
            raise ValueError("q_start and q_goal must have the same shape")
        e = qs.shape[0]
        if out is not None:
            if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.int32
                    and out.is_contiguous() and out.numel() >= device.mask_words(e)):
                raise ValueError("out must be a contiguous CUDA int32 tensor of >= ceil(E/32) words")
            words, packed, host = out, True, False
        else:
            words = torch.zeros(device.mask_words(e), dtype=torch.int32, device="cuda")
        near = torch.zeros(device.mask_words(e), dtype=torch.int32, device="cuda") if return_near else None
        _native.check(
            _native.load().acro_swept_cc_f64(
                self._handle(),
                scene.native_handle() if scene is not None else None,
                qs.data_ptr(), qg.data_ptr(), e, int(n_samples), words.data_ptr(),
                near.data_ptr() if near is not None else None, float(margin),
                device.current_stream_ptr(),
            ),
            "acro_swept_cc_f64",
        )
        if packed:
            res = words if not host else words.cpu().numpy().view(np.uint32)
            if return_near:
                return res, (near if not host else near.cpu().numpy().view(np.uint32))
            return res
        hit = device.unpack_mask(words, e, host)
        if return_near:
            return hit, device.unpack_mask(near, e, host)
        return hit

    def is_path_in_collision(self, q_start, q_goal, collection):
        qs = np.asarray(q_start, dtype=np.float64).reshape(1, -1)
        qg = np.asarray(q_goal, dtype=np.float64).reshape(1, -1)
        return bool(self.is_path_in_collision_batch(qs, qg, collection)[0])

    def is_path_in_collision_discrete(self, q_start, q_goal, collection, max_q_step=0.1):
        qs = np.asarray(q_start, dtype=np.float64).reshape(1, -1)
        qg = np.asarray(q_goal, dtype=np.float64).reshape(1, -1)
        return bool(self.is_path_in_collision_discrete_batch(qs, qg, collection, max_q_step)[0])
