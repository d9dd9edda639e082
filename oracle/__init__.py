"""CPU oracle for the acrobotics FK + collision hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``acrobotics_b200``) never imports this
package and fails loudly if its CUDA library is missing.

Contents
--------
fcl_boxbox.py   scalar FP64 restatement of FCL's ``boxBox2`` separating-axis
                test (the third-party arithmetic behind ``fcl.collide`` for
                Box-Box, reached from reference ``src/acrobotics/shapes.py:50-58``).
np_oracle.py    vectorised FP64 numpy restatement of FK / fk_all_links / fk_rpy,
                link-box vs scene collision, self collision, Kuka / Arm2 /
                wrist / anthropomorphic IK, geometric + rpy Jacobians and the
                discretised path sweep, over (N, ndof) batches.
scalar_port.py  per-configuration port that keeps the reference's cost
                structure (python loops, 4x4 numpy products, one native call
                per box pair); this is the timed "reference arm" on machines
                where ``/root/reference`` does not exist.
ref_shim.py     imports the *verbatim* reference from ``/root/reference`` with
                stub modules for its absent third-party dependencies (only
                usable in the build container; used to pin the oracle and to
                generate ``tests/golden``).
make_golden.py  regenerates ``tests/golden/*.npz`` from the verbatim reference.
c/              plain-C restatement (fast checker for 10^6..10^7 configs and
                the native box-box call used by scalar_port).

Parity status
-------------
FK, fk_all_links, fk_rpy, IK and the collision *orchestration* are pinned
against the reference's own Python code run verbatim through ``ref_shim`` and
against every golden vector the reference's tests/README hold for this path.
The Box-Box predicate itself lives in python-fcl/FCL (un-vendored, unpinned
version, not installable here): it is restated from FCL's published
``boxBox2`` algorithm and pinned only by the reference's coarse known-answer
cases (tests/test_shapes.py:128-142, tests/test_geometry.py:42-59,
tests/test_robot.py:15-41, README.md:94-96).  Behaviour within ~1e-6 m of
contact is therefore "parity unpinned" and reported in a separate near-margin
bucket.  Jacobians (casadi AD in the reference) are restated analytically and
pinned by the reference test's own method (finite differences of the verbatim
``fk`` / ``fk_rpy``).
"""
