"""Generate ``tests/golden/ccd.npz``: swept-collision verdicts of the VERBATIM reference
(``Robot.is_path_in_collision``, robot.py:285-317 -> geometry.py:68-81 -> shapes.py:60-75)
run through ``oracle/ref_shim.py``, whose ``fcl.continuousCollide`` is the restatement in
``oracle/fcl_ccd.py`` (python-fcl is absent: parity with python-fcl itself is unpinned beyond
the three known answers of the reference's ``tests/test_continuous_cc.py:35-83``).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden_ccd

Contents
  known_*          the three set-ups of test_continuous_cc.py; ``known_sticky`` = verdicts of
                   one shared robot over the sequence with python-fcl's accumulating result
                   object (True, True, True = the reference's asserts); ``known_stateless`` =
                   the same with fresh results per query (True, True, False)
  tool_*           the box-only ``torch2`` tool of the reference's test robot
  s3_*  / s16_*    Kuka + torch2, random (q_start, q_goal) pairs vs a three-box scene (table +
                   the plates of the tests) and vs the 16-box benchmark scene, stateless
  full_*           Kuka + torch2 + base geometry + moved base vs the three-box scene
  rail_*           KukaOnRail (prismatic first joint, no tool) vs the 16-box scene
"""
import os

import numpy as np

from oracle import fcl_ccd, ref_shim
from oracle.make_golden import GOLDEN, _q, _ref_scene
from oracle.ref_shim import pose_z, translation
from acrobotics_b200.synthetic import scene16_spec


def main():
    ab = ref_shim.load_reference()
    from acrobotics.tool_examples import torch2  # the reference's own module

    rng = np.random.default_rng(20261020)
    out = {}
    out["tool_sides"] = np.array([[s.dx, s.dy, s.dz] for s in torch2.s])
    out["tool_tfs"] = np.array(torch2.tf_s)
    out["tool_tip"] = np.array(torch2.tf_tool_tip)

    # --- the reference's three known answers (tests/test_continuous_cc.py:35-83) ---
    table = ([2, 2, 0.1], translation(0, 0, -0.2))
    known = [
        (([0.01, 0.01, 1.5], translation(0, 0.5, 0.55)), [1.0, 1.5, -0.3, 0, 0, 0], [2.0, 1.5, 0.3, 0, 0, 0]),
        (([0.2, 0.1, 0.01], translation(0, 0.9, 0.55)), [1.5, 1.5, -0.3, 0, 0.3, 0], [1.5, 1.5, 0.3, 0, -0.3, 0]),
        (([0.01, 0.2, 0.2], translation(0, 1.2, 0)), [1.0, 1.2, -0.5, 0, 0, 0], [2.0, 1.2, -0.5, 0, 0, 0]),
    ]
    out["known_sides"] = np.array([[table[0], k[0][0]] for k in known], float)
    out["known_tfs"] = np.array([[table[1], k[0][1]] for k in known])
    out["known_q_start"] = np.array([k[1] for k in known], float)
    out["known_q_goal"] = np.array([k[2] for k in known], float)
    for name, sticky in (("known_sticky", True), ("known_stateless", False)):
        fcl_ccd.STICKY = sticky
        for s in torch2.s:  # module-level shapes: start from fresh result objects
            s.c_res = fcl_ccd.ContinuousCollisionResult()
        robot = ab.Kuka()
        robot.tool = torch2
        res = []
        for i in range(3):
            scene = _ref_scene(ab, out["known_sides"][i], out["known_tfs"][i])
            res.append(bool(robot.is_path_in_collision(out["known_q_start"][i], out["known_q_goal"][i], scene)))
        out[name] = np.array(res)
    assert out["known_sticky"].tolist() == [True, True, True], out["known_sticky"]

    # --- random pairs, stateless ---
    fcl_ccd.STICKY = False

    def pairs(robot, n, span):
        qs = _q(rng, n, robot)
        qg = qs + rng.uniform(-span, span, qs.shape)
        qg[0] = qs[0]  # no motion at all
        return qs, qg

    def run(robot, scene, qs, qg):
        return np.array([bool(robot.is_path_in_collision(a, b, scene)) for a, b in zip(qs, qg)])

    s3_sides = np.array([[2, 2, 0.1], [0.01, 0.2, 0.2], [0.01, 0.01, 1.5]], float)
    s3_tfs = np.array([translation(0, 0, -0.6), translation(0, 1.2, 0), translation(0, 0.5, 0.55)])
    s3 = _ref_scene(ab, s3_sides, s3_tfs)
    s16_sides, s16_tfs = scene16_spec(0)
    s16 = _ref_scene(ab, s16_sides, s16_tfs)

    kuka = ab.Kuka()
    kuka.tool = torch2
    qs, qg = pairs(kuka, 1500, 1.2)
    out.update(s3_sides=s3_sides, s3_tfs=s3_tfs, s3_q_start=qs, s3_q_goal=qg, s3_hit=run(kuka, s3, qs, qg))
    qs, qg = pairs(kuka, 600, 0.8)
    out.update(s16_q_start=qs, s16_q_goal=qg, s16_hit=run(kuka, s16, qs, qg))

    full = ab.Kuka()
    full.tool = torch2
    full.tf_base = pose_z(0.3, 0.1, 0.05, 0.02)
    base_sides = np.array([[0.4, 0.4, 0.2]])
    base_tfs = np.array([translation(0, 0, -0.1)])
    full.geometry_base = _ref_scene(ab, base_sides, base_tfs)
    qs, qg = pairs(full, 300, 0.8)
    out.update(full_base_tf=full.tf_base, full_base_sides=base_sides, full_base_tfs=base_tfs,
               full_q_start=qs, full_q_goal=qg, full_hit=run(full, s3, qs, qg))

    rail = ab.KukaOnRail()
    qs, qg = pairs(rail, 300, 0.8)
    qs[:, 0] = rng.uniform(-1.0, 1.0, len(qs))
    qg[:, 0] = qs[:, 0] + rng.uniform(-0.5, 0.5, len(qs))
    out.update(rail_q_start=qs, rail_q_goal=qg, rail_hit=run(rail, s16, qs, qg))

    np.savez_compressed(os.path.join(GOLDEN, "ccd.npz"), **out)
    for k in ("known_sticky", "known_stateless"):
        print(k, out[k])
    for k in ("s3_hit", "s16_hit", "full_hit", "rail_hit"):
        print(k, out[k].mean())


if __name__ == "__main__":
    main()
