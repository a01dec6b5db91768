"""Host-side polynomial-chaos set-up of the Stochastic NMPC controller, a mirror of
``Model_Predictive_Controller/Stochastic_NMPC/stochastic_mpc_utils.py:17-91`` without chaospy.

The reference draws its collocation samples with ``chaospy.J(Normal(0,1),...).sample(n, rule='hammersley')``
(chaospy 4.3.13).  chaospy is not installable here; its Hammersley rule is restated from the committed
``acados_ocp_SNMPC.json``: dimensions 0..d-2 are van-der-Corput sequences in the bases 2, 3, 5, ...
started at index ``burnin + 1`` with ``burnin = max(primes used)``, the last dimension is the equispaced grid
``(i+1)/(n+1)``; the uniform points are mapped through the normal quantile.  ``tests/test_pce_host.py`` checks
the result against the sample matrix stored in that JSON."""
from __future__ import annotations

import itertools
import math
from typing import Tuple

import numpy as np
from scipy.special import ndtri

_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19]


def van_der_corput(index: int, base: int) -> float:
    out, f = 0.0, 1.0 / base
    while index > 0:
        out += (index % base) * f
        index //= base
        f /= base
    return out


def hammersley_normal(n_samples: int, n_vars: int) -> np.ndarray:
    """[n_vars, n_samples] standard-normal Hammersley samples as chaospy 4.3.13 produces them."""
    u = np.zeros((n_vars, n_samples))
    if n_vars > 1:
        primes = _PRIMES[: n_vars - 1]
        burnin = max(primes)
        for d, p in enumerate(primes):
            u[d] = [van_der_corput(i + burnin + 1, p) for i in range(n_samples)]
    u[n_vars - 1] = (np.arange(n_samples) + 1.0) / (n_samples + 1.0)
    return ndtri(u)


def hermite(x, n: int):
    """Normalised Hermite recursion of ``hermiteGeneration`` (:17-25), reproduced literally (it divides by
    sqrt(n!) at every level, which is only the orthonormal polynomial up to degree 2 -- all that is used)."""
    if n == 0:
        return (1.0 + 0.0 * x) / math.sqrt(math.factorial(n))
    if n == 1:
        return x / math.sqrt(math.factorial(n))
    return (x * hermite(x, n - 1) - (n - 1) * hermite(x, n - 2)) / math.sqrt(math.factorial(n))


def alpha_generation(n_rand: int, n: int) -> np.ndarray:
    """Multi-indices of total degree <= n in the reference's order (``alphaGeneration`` :27-39)."""
    alphas = np.array([i for i in itertools.product(range(0, n + 1), repeat=n_rand)])
    alphas = alphas[np.sum(alphas, axis=1) <= n]
    alphas = alphas[np.argsort(alphas.sum(axis=1))[::-1], :]
    return alphas[::-1]


def samples_and_A(n_samples: int, x0_std, degree: int = 2) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(w_samples [n_vars, n_samples], A [L, n_samples], alphas): ``computeSamplesAndAmatrix`` (:56-76)."""
    std = np.asarray(x0_std, float)
    active = np.nonzero(std)[0]
    n_vars = len(active)
    alphas = alpha_generation(n_vars, degree)
    w = hammersley_normal(n_samples, n_vars)
    Phi = np.zeros((n_samples, len(alphas)))
    for i in range(n_samples):
        for j, al in enumerate(alphas):
            t = 1.0
            for d in range(n_vars):
                t *= hermite(w[d, i], int(al[d]))
            Phi[i, j] = t
    A = np.linalg.inv(Phi.T @ Phi) @ Phi.T
    return w, A, alphas


def sample_offsets(w_samples: np.ndarray, x0_std) -> np.ndarray:
    """[n_samples, 8] offsets so that ``x0_i = x0 + offsets[i]`` (``compute_x0dist`` :78-91)."""
    std = np.asarray(x0_std, float)
    active = np.nonzero(std)[0]
    off = np.zeros((w_samples.shape[1], 8))
    for i in range(w_samples.shape[1]):
        off[i, active] = std[active] * w_samples[:, i]
    return off
