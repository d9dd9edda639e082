"""Search strategies and sampling settings of the sampling-based planner.

``SearchStrategy`` and ``SamplingSetting`` follow the reference
(``src/acrobotics/path/sampling.py:5-62``) field for field.  ``SampleMethod`` and ``Sampler``
come from ``acrolib.sampling`` in the reference - an un-vendored, unpinned dependency - and
are restated here from their call sites (``path/path_pt.py:69-87``: ``Sampler().sample(n, dim,
method)`` returns an ``(n, dim)`` array of numbers in ``[0, 1]``; ``random_uniform`` draws
i.i.d. uniform numbers, ``deterministic_uniform`` continues a Halton sequence, so that
successive calls refine the same low-discrepancy point set).  Parity of the drawn NUMBERS with
acrolib is unpinned; every property the reference's tests assert (ranges, fixed coordinates,
counts: ``tests/test_path.py:88-125,139-156``) holds.
"""
import enum

import numpy as np


class SearchStrategy(enum.Enum):
    GRID = 0
    INCREMENTAL = 1
    MIN_INCREMENTAL = 2


class SampleMethod(enum.Enum):
    random_uniform = 0
    deterministic_uniform = 1


_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def _van_der_corput(index: np.ndarray, base: int) -> np.ndarray:
    out = np.zeros(len(index))
    f = 1.0 / base
    i = index.copy()
    while (i > 0).any():
        out += f * (i % base)
        i //= base
        f /= base
    return out


class Sampler:
    """Incremental source of points in the unit cube."""

    def __init__(self, seed=None):
        self.rng = np.random.default_rng(seed)
        self.count = 0  # Halton points handed out so far

    def sample(self, num_samples: int, dim: int, method: SampleMethod = SampleMethod.random_uniform):
        if dim > len(_PRIMES):
            raise ValueError(f"at most {len(_PRIMES)} sampled dimensions")
        if method == SampleMethod.random_uniform:
            return self.rng.random((num_samples, dim))
        if method == SampleMethod.deterministic_uniform:
            idx = np.arange(self.count + 1, self.count + 1 + num_samples)
            self.count += num_samples
            return np.stack([_van_der_corput(idx, _PRIMES[d]) for d in range(dim)], axis=1).reshape(
                num_samples, dim)
        raise NotImplementedError(method)


class SamplingSetting:
    """Settings of the sampling-based planner (reference path/sampling.py:16-62)."""

    def __init__(
        self,
        search_strategy: SearchStrategy,
        iterations: int = None,
        sample_method: SampleMethod = None,
        num_samples: int = None,
        desired_num_samples: int = None,
        max_search_iters: int = None,
        tolerance_reduction_factor: float = None,
        cost_function=None,
        weights=None,
        use_state_cost=False,
        state_cost_weight=1.0,
    ):
        self.cost_function = cost_function
        self.weights = weights
        self.use_state_cost = use_state_cost
        self.state_cost_weight = state_cost_weight

        if iterations is None or iterations == 1:
            self.iterations = 1
        else:
            self.iterations = iterations
            assert tolerance_reduction_factor is not None
            self.tolerance_reduction_factor = tolerance_reduction_factor

        self.search_strategy = search_strategy
        if self.search_strategy == SearchStrategy.GRID:
            pass
        elif self.search_strategy == SearchStrategy.INCREMENTAL:
            assert sample_method is not None
            assert num_samples is not None
            self.sample_method = sample_method
            self.num_samples = num_samples
        elif self.search_strategy == SearchStrategy.MIN_INCREMENTAL:
            assert sample_method is not None
            assert desired_num_samples is not None
            assert max_search_iters is not None
            self.sample_method = sample_method
            self.desired_num_samples = desired_num_samples
            self.max_search_iters = max_search_iters
            self.step_size = 1
