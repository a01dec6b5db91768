"""Host-side polynomial-chaos set-up of the Stochastic NMPC controller, a mirror of
``Model_Predictive_Controller/Stochastic_NMPC/stochastic_mpc_utils.py:17-91`` without chaospy.

The reference draws its collocation samples with ``chaospy.J(Normal(0,1),...).sample(n, rule='hammersley')``
(chaospy 4.3.13).  chaospy is not installable here; its Hammersley rule is restated from the committed
``acados_ocp_SNMPC.json``: dimensions 0..d-2 are van-der-Corput sequences in the bases 2, 3, 5, ...
started at index ``burnin + 1`` with ``burnin = max(primes used)``, the last dimension is the equispaced grid
``(i+1)/(n+1)``; the uniform points are mapped through the normal quantile.  ``tests/test_oracle_snmpc.py`` checks
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
    """Basis polynomial of degree ``n`` in one standard-normal variable as the reference's PCE uses it
    (``hermiteGeneration``, stochastic_mpc_utils.py:17-25): He_0 = 1, He_1 = x, He_2 = (x^2 - 1) / sqrt(2) -- the
    orthonormal probabilists' Hermite polynomials, all that degree <= 2 needs.  Beyond degree 2 the reference's
    three-term recurrence divides by sqrt(k!) at EVERY level (not orthonormal any more); that behaviour is kept, as a
    forward loop over the degrees, so that a higher ``expansion_degree`` still builds the reference's matrix."""
    x = np.asarray(x, float)
    lower, cur = np.ones_like(x), x.copy()             # degrees 0 and 1
    if n == 0:
        return lower
    for k in range(2, n + 1):
        lower, cur = cur, (x * cur - (k - 1) * lower) / math.sqrt(math.factorial(k))
    return cur


def alpha_generation(n_rand: int, n: int) -> np.ndarray:
    """Multi-indices of total degree <= n, graded: the constant term first, then by total degree and lexicographic inside
    a degree -- [000, 001, 010, 100, 002, 011, 020, 101, 110, 200] for (3, 2), the order ``alphaGeneration``
    (stochastic_mpc_utils.py:27-39) produces for the committed configuration.  Only "constant term first"
    matters downstream: row 0 of A is the mean, the other rows enter the variance as a sum of squares."""
    out = []
    for total in range(n + 1):
        out.extend(sorted(a for a in itertools.product(range(total + 1), repeat=n_rand) if sum(a) == total))
    return np.array(out, dtype=int).reshape(len(out), n_rand)


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
                t *= float(hermite(w[d, i], int(al[d])))
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
