"""TEST INFRASTRUCTURE: numpy helpers for the reference's even-odd layout
(/root/reference/lattice/lattice.h:75-81) and the U(1) fixtures under tests/golden/."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def site_index(x, y, X, Y):
    """coord_to_index (lattice.h:75-81): all even sites, then all odd; x fastest inside a parity."""
    x = np.asarray(x)
    y = np.asarray(y)
    par = (x + y) & 1
    return (y + par * Y) * (X // 2) + x // 2


def site_coords(X, Y):
    """index_to_coord for every site index (lattice.h:199-205): returns x[i], y[i]."""
    i = np.arange(X * Y)
    half = X * Y // 2
    par = (i >= half).astype(np.int64)
    h = i - par * half
    y = h // (X // 2)
    k = h % (X // 2)
    x = 2 * k + ((y + par) & 1)
    return x, y


def phases_to_gauge(phases, X, Y):
    """File order (x outer, y, mu inner; u1/u1_utils.h:53-63) -> eo gauge array [mu*V + site] of exp(i phase)."""
    ph = np.asarray(phases, dtype=np.float64).reshape(X, Y, 2)
    xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    idx = site_index(xs, ys, X, Y)
    V = X * Y
    g = np.zeros(2 * V, np.complex128)
    for mu in range(2):
        g[mu * V + idx.ravel()] = np.exp(1j * ph[:, :, mu].ravel())
    return g


def load_gauge(L, beta=60):
    """A thermalised reference config (tests/common_cfgs_u1/l{L}t{L}b{beta}_heatbath.dat) from its committed .npy copy."""
    ph = np.load(os.path.join(GOLDEN, "l%dt%db%d_phases.npy" % (L, L, beta)))
    return phases_to_gauge(ph, L, L)


def synthetic_gauge(X, Y, beta=6.0, seed=1337):
    """Gaussian non-compact phases of width 1/sqrt(beta): the large-lattice stand-in for a heatbath config."""
    rng = np.random.default_rng(seed)
    return phases_to_gauge(rng.normal(0.0, 1.0 / np.sqrt(beta), size=X * Y * 2), X, Y)


def gaussian_cv(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex128)


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
