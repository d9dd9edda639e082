"""Per-configuration CPU port with the reference's cost structure (TEST
INFRASTRUCTURE; the timed "reference arm" of bench.py on machines where
``/root/reference`` does not exist).

Every step mirrors what the reference executes for one ``robot.fk(q)`` /
``robot.is_in_collision(q, scene)`` call - python loops over links and box
pairs, numpy scalar trig + 4x4 ``@`` per link (link.py:35-58, robot.py:59-91),
one ``tf_self @ tf`` per shape (geometry.py:56-59) and one native call per box
pair (shapes.py:50-58; the reference builds two fcl.Transform + two
fcl.CollisionObject and calls fcl.collide - here a single ctypes call into
oracle/c/acro_oracle.c, which is if anything cheaper).  Early exits follow the
reference (first colliding pair returns).
"""
import numpy as np

from . import c_oracle
from . import np_oracle as no


class ScalarRobot:
    def __init__(self, spec: no.RobotSpec):
        self.spec = spec
        self.ndof = spec.ndof
        self._T = [np.eye(4) for _ in range(spec.ndof)]  # cached per link like link.py:33
        self._box_box = c_oracle.box_box_fn()
        self.link_boxes = [
            [(np.ascontiguousarray(2.0 * b.half), b.tf) for b in spec.boxes if b.group == i]
            for i in range(spec.ndof)
        ]
        self.tool_boxes = [
            (np.ascontiguousarray(2.0 * b.half), b.tf) for b in spec.boxes if b.group == no.GROUP_TOOL
        ]
        self.base_boxes = [
            (np.ascontiguousarray(2.0 * b.half), b.tf) for b in spec.boxes if b.group == no.GROUP_BASE
        ]
        self.link_pairs = [(i, j) for i, j in spec.self_pairs if i >= 0 and j >= 0]
        self.tool_pairs = [i for i, j in spec.self_pairs if j == no.GROUP_TOOL]

    def link_transform(self, i, qi):
        a, alpha, d, theta = self.spec.dh[i]
        if self.spec.prismatic[i]:
            d = qi
        else:
            theta = qi
        ct, st, ca, sa = np.cos(theta), np.sin(theta), np.cos(alpha), np.sin(alpha)
        T = self._T[i]
        T[0, 0], T[0, 1], T[0, 2], T[0, 3] = ct, -st * ca, st * sa, a * ct
        T[1, 0], T[1, 1], T[1, 2], T[1, 3] = st, ct * ca, -ct * sa, a * st
        T[2, 1], T[2, 2], T[2, 3] = sa, ca, d
        return T

    def fk(self, q):
        T = self.spec.tf_base
        for i in range(self.ndof):
            T = T @ self.link_transform(i, q[i])
        if self.spec.tf_tool_tip is not None:
            T = T @ self.spec.tf_tool_tip
        return T

    def fk_all_links(self, q):
        out = []
        T = self.spec.tf_base
        for i in range(self.ndof):
            T = T @ self.link_transform(i, q[i])
            out.append(T)
        return out

    def _soup_hit(self, boxes1, tf1, boxes2, tf2=None):
        w1 = [(s, np.ascontiguousarray(tf1 @ t)) for s, t in boxes1]
        w2 = boxes2 if tf2 is None else [(s, np.ascontiguousarray(tf2 @ t)) for s, t in boxes2]
        for s1, t1 in w1:
            for s2, t2 in w2:
                if self._box_box(s1, t1, s2, t2):
                    return True
        return False

    def is_in_collision(self, q, scene_boxes):
        """scene_boxes: list of (sides (3,), tf (4,4) contiguous)."""
        tf_links = self.fk_all_links(q)
        if self.tool_boxes and self._soup_hit(self.tool_boxes, tf_links[-1], scene_boxes):
            return True
        if self.base_boxes and self._soup_hit(self.base_boxes, self.spec.tf_base, scene_boxes):
            return True
        for i in range(self.ndof):
            if self._soup_hit(self.link_boxes[i], tf_links[i], scene_boxes):
                return True
        if self.spec.check_self:
            for i, j in self.link_pairs:
                # the reference visits (i, j) and (j, i) (robot.py:207-211)
                if self._soup_hit(self.link_boxes[i], tf_links[i], self.link_boxes[j], tf_links[j]):
                    return True
                if self._soup_hit(self.link_boxes[j], tf_links[j], self.link_boxes[i], tf_links[i]):
                    return True
            for i in self.tool_pairs:
                if self._soup_hit(self.link_boxes[i], tf_links[i], self.tool_boxes, tf_links[-1]):
                    return True
        return False


def scene_list(scene):
    sh, sR, sT = scene if isinstance(scene, tuple) else no.scene_arrays(scene)
    out = []
    for k in range(len(sh)):
        tf = np.eye(4)
        tf[:3, :3], tf[:3, 3] = sR[k], sT[k]
        out.append((np.ascontiguousarray(2.0 * sh[k]), np.ascontiguousarray(tf)))
    return out


def run_chunk(args):
    """Worker for multiprocessing: (spec, scene arrays, q chunk) -> hits."""
    spec, scene, q = args
    bot = ScalarRobot(spec)
    sc = scene_list(scene)
    return np.array([bot.is_in_collision(x, sc) for x in q], dtype=bool)
