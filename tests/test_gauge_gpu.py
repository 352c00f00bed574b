"""U(1) gauge-field side (/root/reference/u1/u1_utils.h) on the GPU against the oracle: observables,
gauge transformation, APE smearing, instantons, file format -- same driver, two back ends -- and the subset heatbath,
which is a different sweep order of the reference's conditional update and is therefore checked statistically (and, on
y-slabs, bit for bit against the periodic run: its random counter is keyed by global coordinates)."""
import numpy as np
import pytest

import capi
import latutil

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not capi.have_ref():
        pytest.skip("oracle/_ref/libqmg_ref.so not built")
    return capi.Backend("ref")


@pytest.fixture(scope="module")
def gpu(qmg_gpu):
    return capi.Backend("gpu")


def test_observables_transform_smear(ref, gpu):
    """n01_u1_test quantities on a shipped thermalised field and on rough gaussian fields."""
    for L, g in ((64, latutil.load_gauge(64)), (16, None), (24, None)):
        lr, lg = ref.lattice(L, L, 1), gpu.lattice(L, L, 1)
        if g is None:
            g = lr.u1_create(2, beta=1.0, seed=L)
            assert np.max(np.abs(g - lg.u1_create(2, beta=1.0, seed=L))) < 1e-15     # host draws: same field on both back ends
        (pa, qa), (pb, qb) = lr.u1_observables(g), lg.u1_observables(g)
        assert abs(pa - pb) < 1e-13 and abs(qa - qb) < 1e-10
        tr = lr.u1_create(3, seed=7)
        ta, tb = lr.u1_gauge_trans(g, tr), lg.u1_gauge_trans(g, tr)
        assert latutil.rel_l2(tb, ta) < 1e-14
        for alpha, n_iter in ((0.5, 1), (0.5, 4), (0.3, 3)):
            sa, sb = lr.u1_ape_smear(g, alpha, n_iter), lg.u1_ape_smear(g, alpha, n_iter)
            assert latutil.rel_l2(sb, sa) < 1e-12, (L, alpha, n_iter)
        ia, ib = lr.u1_instanton(g, 1.0, L // 2, L // 4), lg.u1_instanton(g, 1.0, L // 2, L // 4)
        assert latutil.rel_l2(ib, ia) < 1e-14


def test_ape_smear_textbook(qmg_gpu):
    """textbook = 1: every link smeared with its own two staples (what u1_utils.h:276-383 describes; the reference itself
    accumulates the y staples on the x links, which is the default and is checked against the oracle above)."""
    qmg = qmg_gpu
    L, alpha = 16, 0.4
    g = latutil.synthetic_gauge(L, L, 2.0, 9)
    xs, ys = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    idx = latutil.site_index(xs, ys, L, L)
    ux, uy = g[idx], g[L * L + idx]

    def sh(a, dx, dy):
        return np.roll(np.roll(a, -dx, axis=0), -dy, axis=1)
    for _ in range(2):
        nx = ux + alpha * (uy * sh(ux, 0, 1) * np.conj(sh(uy, 1, 0)) + np.conj(sh(uy, 0, -1)) * sh(ux, 0, -1) * sh(uy, 1, -1))
        ny = uy + alpha * (ux * sh(uy, 1, 0) * np.conj(sh(ux, 0, 1)) + np.conj(sh(ux, -1, 0)) * sh(uy, -1, 0) * sh(ux, -1, 1))
        ux, uy = nx / abs(nx), ny / abs(ny)
    want = np.zeros(2 * L * L, np.complex128)
    want[idx.ravel()] = ux.ravel()
    want[L * L + idx.ravel()] = uy.ravel()
    got = qmg.u1_ape_smear(qmg.to_device(g), L, L, alpha, 2, textbook=True).cpu().numpy()
    assert latutil.rel_l2(got, want) < 1e-13


def test_phase_fields_and_files(ref, gpu, tmp_path):
    L = 16
    lr, lg = ref.lattice(L, L, 1), gpu.lattice(L, L, 1)
    ph = np.random.default_rng(1).normal(0, 0.5, size=2 * L * L)
    assert abs(lg.u1_action(ph, 6.0) - lr.u1_action(ph, 6.0)) < 1e-10 * lr.u1_action(ph, 6.0)
    assert latutil.rel_l2(lg.u1_polar(ph), lr.u1_polar(ph)) < 1e-15
    assert np.allclose(lg.u1_noncompact_instanton(ph, 2.0), lr.u1_noncompact_instanton(ph, 2.0), rtol=0, atol=1e-15)
    g = lr.u1_polar(ph)
    # written by one back end, read by the other, both directions, both file flavours
    pa, pb = str(tmp_path / "a.dat"), str(tmp_path / "b.dat")
    lg.u1_file(1, pa, gauge=g)
    lr.u1_file(1, pb, gauge=g)
    assert open(pa).read() == open(pb).read()
    assert np.max(np.abs(lr.u1_file(0, pa) - lg.u1_file(0, pb))) < 1e-15
    lg.u1_file(3, pa, phases=ph)
    lr.u1_file(3, pb, phases=ph)
    assert open(pa).read() == open(pb).read()
    assert np.array_equal(lg.u1_file(2, pa), lr.u1_file(2, pb))


def test_heatbath_statistics(ref, gpu):
    """Cold start, beta = 6: after thermalisation the subset heatbath and the reference's serial sweep agree on
    <cos plaq> = exp(-1/(2 beta)) and on the action per plaquette 1/2 (equipartition of the free non-compact theory)."""
    L, beta = 64, 6.0
    lr, lg = ref.lattice(L, L, 1), gpu.lattice(L, L, 1)
    cold = np.zeros(2 * L * L)
    want = np.exp(-0.5 / beta)
    pg = lg.u1_heatbath(cold, beta, 200, 11)
    pr = lr.u1_heatbath(cold, beta, 200, 11)
    plg = lg.u1_observables(lg.u1_polar(pg))[0].real
    plr = lr.u1_observables(lr.u1_polar(pr))[0].real
    assert abs(plg - want) < 0.01 and abs(plr - want) < 0.01
    # 200 sweeps from a cold start leave the longest wavelengths slightly cool on BOTH sweeps (action approaches 1/2 per
    # plaquette from below); the two chains must agree with each other more tightly than with the asymptote
    ag, ar = lg.u1_action(pg, beta) / (L * L), lr.u1_action(pr, beta) / (L * L)
    assert abs(ag - 0.5) < 0.06 and abs(ar - 0.5) < 0.06 and abs(ag - ar) < 0.04
    # continuing the chain keeps it there, and different seeds decorrelate
    pg2 = lg.u1_heatbath(pg, beta, 50, 12)
    assert abs(lg.u1_observables(lg.u1_polar(pg2))[0].real - want) < 0.01
    assert not np.allclose(pg2, lg.u1_heatbath(pg, beta, 50, 13))
    # a single link's conditional variance is 1/(2 beta): update once from a FIXED background, many seeds
    samples = np.array([lg.u1_heatbath(cold, beta, 1, s)[5] for s in range(200)])
    assert abs(samples.std() - np.sqrt(0.5 / beta)) < 0.05


def test_gauge_side_on_slabs_loopback(qmg_gpu, gpu):
    """Loopback (every access across y = 0 / Y-1 through the halo-row path): identical results, including the heatbath."""
    qmg = qmg_gpu
    if qmg.comm_counters()["active"]:
        pytest.skip("QMG_LOOPBACK already forced for the whole session")
    X, Y = 32, 16
    lat = gpu.lattice(X, Y, 1)
    g = latutil.synthetic_gauge(X, Y, 6.0, 3)
    ph = np.random.default_rng(4).normal(0, 0.4, size=2 * X * Y)
    tr = lat.u1_create(3, seed=2)

    def run():
        return (lat.u1_observables(g), lat.u1_action(ph, 6.0), lat.u1_gauge_trans(g, tr), lat.u1_ape_smear(g, 0.5, 3),
                lat.u1_heatbath(ph, 6.0, 5, 9))
    a = run()
    qmg.comm_set_loopback(True)
    try:
        b = run()
    finally:
        qmg.comm_set_loopback(False)
    assert a[0] == b[0] and a[1] == b[1]
    for u, v in zip(a[2:], b[2:]):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("X,Y,nc", [(16, 8, 2), (32, 32, 8), (6, 4, 1)])
def test_timeslice_reductions_and_wall_source(ref, gpu, X, Y, nc):
    """reductions/reductions.h: per-row norm2sq / re_dot / dot and the gaussian wall source (the correlator measurement of
    tests/n16_wilson_kcycle_heatbath), against the oracle and against a direct numpy sum."""
    lr, lg = ref.lattice(X, Y, nc), gpu.lattice(X, Y, nc)
    a, b = latutil.gaussian_cv(X * Y * nc, 1), latutil.gaussian_cv(X * Y * nc, 2)
    for op in (0, 1, 2):
        ra, rb = lr.timeslice(op, a, b), lg.timeslice(op, a, b)
        assert np.allclose(rb, ra, rtol=1e-13, atol=1e-13), op
    xs, ys = latutil.site_coords(X, Y)
    want = np.zeros(Y)
    np.add.at(want, np.repeat(ys, nc), np.abs(a) ** 2)
    assert np.allclose(lg.timeslice(0, a), want, rtol=1e-13)
    wa, wb = lr.wall_source(Y // 2, nc - 1, 5, 0.7, 0.1), lg.wall_source(Y // 2, nc - 1, 5, 0.7, 0.1)
    assert np.array_equal(wa, wb)
    assert np.count_nonzero(wb) == X
