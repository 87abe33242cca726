"""Host-side mirror of scorus::linear_space::utils (src/linear_space/utils.rs) over the C ABI.

    mean(x)   # :4-15   ensemble mean
    cov(x)    # :17-41  biased (/N) covariance about the mean; the reference reports it through a
              #         callback cov_result(i, j, value), here it is returned as a [dim, dim] array
"""
from __future__ import annotations

import numpy as np

from .ensemble_sample import Gaussian, Sampler, _check_arrays


def mean(x: np.ndarray) -> np.ndarray:
    _check_arrays(x, None)
    assert x.shape[0] > 0  # assert!(!x.is_empty()), utils.rs:12
    with Sampler(Gaussian(), x.shape[0] + (x.shape[0] & 1), x.shape[1]) as s:
        s.upload(x if x.shape[0] % 2 == 0 else np.ascontiguousarray(np.vstack([x, x[-1:]])), np.zeros(s.n))
        return s.mean(0, x.shape[0])


def cov(x: np.ndarray) -> np.ndarray:
    _check_arrays(x, None)
    assert x.shape[0] > 0
    with Sampler(Gaussian(), x.shape[0] + (x.shape[0] & 1), x.shape[1]) as s:
        s.upload(x if x.shape[0] % 2 == 0 else np.ascontiguousarray(np.vstack([x, x[-1:]])), np.zeros(s.n))
        return s.cov(0, x.shape[0])
