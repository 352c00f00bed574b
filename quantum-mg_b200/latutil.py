"""numpy helpers for the reference's even-odd layout (/root/reference/lattice/lattice.h:75-81), the reference's
thermalised U(1) configs (re-encoded under tests/golden/) and the synthetic large-lattice gauge fields bench.py runs on."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


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


def synthetic_phases(X, Y, beta=6.0, seed=1337, slab=False):
    """Link phases with the plaquette statistics of the 2D non-compact U(1) theory at coupling beta -- the large-lattice
    stand-in for the reference's serial heatbath (u1/u1_utils.h:607-667), which needs hours beyond 1024^2.
    In 2D the plaquette angles are independent gaussians of variance 1/beta (up to the torus constraints): draw
    F(x,y) ~ N(0, 1/beta) with zero total flux per column, integrate it into theta_x in the gauge theta_y = 0, then apply a
    random gauge transformation.  <cos plaquette> = exp(-1/(2 beta)) = 0.920 at beta = 6, as in tests/common_cfgs_u1.
    slab=True: the gauge transformation is trivial on row y = 0.  Such fields can be stacked in y into one valid
    periodic configuration (theta_x vanishes on row 0 of every slab because the flux per column is zero slab by slab), so N
    ranks can each draw their own (X, Y) slab of an (X, N Y) lattice without seeing the others.
    Returns phases in the reference's file order (x outer, y, mu inner)."""
    rng = np.random.default_rng(seed)
    F = rng.normal(0.0, 1.0 / np.sqrt(beta), size=(X, Y))
    F -= F.mean(axis=1, keepdims=True)
    thx = -(np.cumsum(F, axis=1) - F)
    del F
    a = rng.uniform(-np.pi, np.pi, size=(X, Y))
    if slab:
        a[:, 0] = 0.0
    thx += np.roll(a, -1, axis=0) - a
    thy = np.roll(a, -1, axis=1) - a
    return np.stack([thx, thy], axis=2).ravel()


def synthetic_gauge(X, Y, beta=6.0, seed=1337, slab=False):
    return phases_to_gauge(synthetic_phases(X, Y, beta, seed, slab), X, Y)


def synthetic_gauge_units(X, Y, unit_rows, first_unit, beta=6.0, seed=1337):
    """Rows [first_unit * unit_rows, ... + Y) of the periodic field that is the stack of stackable (X, unit_rows) slabs, slab u drawn
    with seed + u.  Every rank of an N-way y-slab decomposition, for every N that divides the number of units, gets its part of the
    SAME global field this way (strong scaling on one lattice), and a taller lattice is the same field continued (weak scaling)."""
    if Y % unit_rows:
        raise ValueError("synthetic_gauge_units: Y must be a multiple of unit_rows")
    parts = [synthetic_phases(X, unit_rows, beta, seed + first_unit + u, slab=True).reshape(X, unit_rows, 2) for u in range(Y // unit_rows)]
    return phases_to_gauge(np.concatenate(parts, axis=1).ravel(), X, Y)


def average_plaquette(gauge, X, Y):
    """<Re U_x(x) U_y(x+x^) U_x*(x+y^) U_y*(x)> (u1/u1_utils.h:424-460)."""
    xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    idx = site_index(xs, ys, X, Y)
    ux, uy = gauge[idx], gauge[X * Y + idx]
    return float(np.mean((ux * np.roll(uy, -1, axis=0) * np.conj(np.roll(ux, -1, axis=1)) * np.conj(uy)).real))


def gaussian_cv(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex128)


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
