"""Synthetic workloads of the benchmark contract (BASELINE.json / SURVEY.md §8d).

Pure numpy / torch input generators - no reference code, no collision logic.
"""
import numpy as np

from .boxes import Box, Scene
from .transforms import translation


def scene16_spec(seed=0, n_boxes=16):
    """The 16-box benchmark scene as arrays: box 0 is the README table
    ``Box(2, 2, 0.1)`` at (0, 0, -0.2); boxes 1.. have sides U[0.05, 0.4]^3,
    centres U[-1.2, 1.2]^2 x U[0, 1.4] at least 0.45 m from the base axis and
    uniformly random orientations.  Returns (sides (n,3), tfs (n,4,4))."""
    rng = np.random.default_rng(seed)
    sides = [np.array([2.0, 2.0, 0.1])]
    tfs = [translation(0, 0, -0.2)]
    while len(sides) < n_boxes:
        ext = rng.uniform(0.05, 0.4, 3)
        c = np.array([rng.uniform(-1.2, 1.2), rng.uniform(-1.2, 1.2), rng.uniform(0.0, 1.4)])
        quat = rng.normal(size=4)
        if np.hypot(c[0], c[1]) < 0.45:
            continue
        w, x, y, z = quat / np.linalg.norm(quat)
        tf = np.eye(4)
        tf[:3, :3] = [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
        tf[:3, 3] = c
        sides.append(ext)
        tfs.append(tf)
    return np.array(sides), np.array(tfs)


def scene16(seed=0, n_boxes=16):
    sides, tfs = scene16_spec(seed, n_boxes)
    return Scene([Box(*s) for s in sides], [t.copy() for t in tfs])


def readme_scene():
    """Table + pillar of the reference README (README.md:72-79)."""
    return Scene(
        [Box(2, 2, 0.1), Box(0.2, 0.2, 1.5)], [translation(0, 0, -0.2), translation(0, 0.5, 0.55)]
    )


def readme_path():
    """README.md:84-86."""
    return np.linspace([0.5, 1.5, -0.3, 0, 0, 0], [2.5, 1.5, 0.3, 0, 0, 0], 10)


def random_configs(n, ndof=6, seed=0):
    """(n, ndof) float64, i.i.d. U[-pi, pi) (numpy; for parity-size inputs)."""
    return np.random.default_rng(seed).uniform(-np.pi, np.pi, (n, ndof))


def random_configs_device(n, ndof=6, seed=0, device="cuda"):
    """Same distribution generated on the GPU (for benchmark-size inputs)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    q = torch.rand((n, ndof), dtype=torch.float64, device=device, generator=g)
    return q.mul_(2 * np.pi).sub_(np.pi)
