"""Seeded synthetic polarization stacks (generators G1-G3) for tests and bench.

Host-side NumPy generators only (`numpy.random.default_rng(seed)`); they produce the uint16 [N,H,W]
angle-major stacks the reference reads from TIFF (PyPOLAR.py:1923-1933). No reference code involved.
"""
from __future__ import annotations

import numpy as np

K_TH_3D = np.array([[0.3530, 0.0030, 0.0750, 0.0], [0.0030, 0.3530, 0.0750, 0.0],
                    [0.3220, 0.3220, 0.3070, 0.5840], [0.3220, 0.3220, 0.3070, -0.5840]])
K_TH_2D = np.array([[0.4770, 0.0230, 0.0], [0.0230, 0.4770, 0.0], [0.2500, 0.2500, 0.4540], [0.2500, 0.2500, -0.4540]])


def _grid(H, W, rho_lo: float = 0.0, rho_span: float = 180.0):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    rho0 = np.deg2rad((rho_lo + rho_span * xx / W) % 180.0)
    return yy, xx, rho0


def stack_1pf(H: int, W: int, N: int = 18, seed: int = 0, dark: int = 480, dtype=np.uint16,
              level: float = 2000.0, amp: float = 1500.0, rho_lo: float = 0.0, rho_span: float = 180.0) -> np.ndarray:
    """G1: one-photon fluorescence stack, orientation sweeping along x (over [rho_lo, rho_lo + rho_span) degrees: the default
    covers every orientation; a span of a few degrees is an aligned fibre, all of whose pixels fall into two or three
    histogram bins), modulation depth along y."""
    rng = np.random.default_rng(seed)
    yy, xx, rho0 = _grid(H, W, rho_lo, rho_span)
    m = 0.05 + 0.9 * yy / H
    I0 = level + amp * np.sin(xx / 37.0) * np.cos(yy / 53.0)
    alpha = (-np.pi * np.arange(N) / N)[:, None, None]
    lam = I0 * (1.0 + m * np.cos(2.0 * (alpha - rho0)))
    v = rng.poisson(lam) + dark
    return np.clip(v, 0, np.iinfo(dtype).max).astype(dtype)


def stack_4harm(H: int, W: int, N: int = 36, seed: int = 0, dark: int = 0, dtype=np.uint16) -> np.ndarray:
    """G2: second+fourth harmonic response (SHG / CARS / SRS / 2PF)."""
    rng = np.random.default_rng(seed)
    _, _, rho0 = _grid(H, W)
    alpha = (-np.pi * np.arange(N) / N)[:, None, None]
    lam = 1500.0 * (1.0 + 0.5 * np.cos(2.0 * (alpha - rho0)) + 0.2 * np.cos(4.0 * (alpha - rho0)))
    v = rng.poisson(lam) + dark
    return np.clip(v, 0, np.iinfo(dtype).max).astype(dtype)


def stack_4polar(H: int, W: int, seed: int = 0, three_d: bool = True, dtype=np.uint16) -> np.ndarray:
    """G3: 4 polarization channels in acquisition order (0, 45, 90, 135 deg), i.e. *before* the
    reference's plane swap at PyPOLAR.py:2957."""
    rng = np.random.default_rng(seed)
    yy, _, rho0 = _grid(H, W)
    p = 0.1 + 0.8 * yy / H
    if three_d:
        pzz = 0.2
        s = 1.0 - pzz
        # (Pxx, Pyy, Pzz, Puv): Pxx+Pyy = s, Pxx-Pyy = s*p*cos(2rho), Puv = s*p*sin(2rho)
        P = np.stack([s * (1 + p * np.cos(2 * rho0)) / 2, s * (1 - p * np.cos(2 * rho0)) / 2,
                      np.full_like(p, pzz), s * p * np.sin(2 * rho0) / 2])
        K = K_TH_3D
    else:
        P = np.stack([(1 + p * np.cos(2 * rho0)) / 2, (1 - p * np.cos(2 * rho0)) / 2, p * np.sin(2 * rho0) / 2])
        K = K_TH_2D
    chan = np.einsum('ij,j...->i...', K, P) * 3000.0          # rows of K: (0, 90, 45, 135)
    chan = chan[[0, 2, 1, 3]]                                  # acquisition order (0, 45, 90, 135)
    v = rng.poisson(np.maximum(chan, 0))
    return np.clip(v, 0, np.iinfo(dtype).max).astype(dtype)


def adversarial_1pf(H: int = 32, W: int = 48, N: int = 18, seed: int = 0) -> np.ndarray:
    """Edge-case pixels embedded in a G1 stack: all-zero, constant, saturated,
    single hot angle, below-dark, pure cosine (chi2 -> 0)."""
    v = stack_1pf(H, W, N, seed)
    v[:, 0, :8] = 0
    v[:, 1, :8] = 1000
    v[:, 2, :8] = 65535
    v[:, 3, :8] = 480
    v[3, 3, :8] = 60000
    v[:, 4, :8] = 100
    alpha = -np.pi * np.arange(N) / N
    v[:, 5, :8] = np.round(2000 + 1000 * np.cos(2 * alpha))[:, None] + 480
    v[:, 6, :8] = (np.arange(N) % 2 * 4000 + 480)[:, None]
    return v
