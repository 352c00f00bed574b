"""The algebra behind the fused solver kernels, checked in numpy on small dense problems (no GPU, no library):
what quantum-mg_b200/csrc/qmg_blas.cu computes in fewer passes is what the step-by-step recurrences compute."""
import numpy as np


def _rand(rng, *shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


def dot(x, y):
    return np.vdot(x, y)          # sum conj(x) y, quantum-linalg's dot


def test_two_pass_mr2_equals_two_mr_steps():
    """qmg_mr2_gram / qmg_mr2_update (inverters/generic_minres.h, SOLVE_TWO_STEP_MR): A applied to q1 = A r0 instead of to r1,
    both step lengths from the six Gram entries of (r0, q1, p2), x and r2 as combinations of the three vectors."""
    rng = np.random.default_rng(3)
    n, omega = 60, 0.85
    A = _rand(rng, n, n) + 4.0 * np.eye(n)
    r0 = _rand(rng, n)
    # step by step (oracle/qlinalg_shim/inverters/generic_minres.h from a zero start)
    x, r = np.zeros(n, complex), r0.copy()
    for _ in range(2):
        q = A @ r
        alpha = omega * dot(q, r) / dot(q, q).real
        x = x + alpha * r
        r = r - alpha * q
    # two passes
    q1 = A @ r0
    p2 = A @ q1
    a, b, c, d, e = dot(q1, r0), dot(q1, q1).real, dot(p2, r0), dot(p2, q1), dot(p2, p2).real
    a1 = omega * a / b
    q2r1 = a - a1 * b - np.conj(a1) * c + abs(a1) ** 2 * d
    q2q2 = b - 2.0 * (np.conj(a1) * d).real + abs(a1) ** 2 * e
    a2 = omega * q2r1 / q2q2
    x2 = (a1 + a2) * r0 - a1 * a2 * q1
    r2 = r0 - (a1 + a2) * q1 + a1 * a2 * p2
    assert np.linalg.norm(x2 - x) < 1e-13 * np.linalg.norm(x)
    assert np.linalg.norm(r2 - r) < 1e-13 * np.linalg.norm(r0)
    assert np.linalg.norm(r0 - A @ x2 - r2) < 1e-13 * np.linalg.norm(r0)      # the residual the K-cycle takes over IS b - A x
    f = dot(r0, r0).real
    r1sq = f - 2.0 * (np.conj(a1) * a).real + abs(a1) ** 2 * b
    assert abs(r1sq - np.linalg.norm(r0 - a1 * q1) ** 2) < 1e-12 * f


def test_preconditioner_can_hand_over_A_times_its_output():
    """PrecondAzRequest (inverters/generic_gcr.h): a smoother-corrected iterate z = z_mid + z3, z3 = MR steps on r2 = rhs - A z_mid,
    satisfies A z = rhs - r', r' the smoother's recurrence residual -- what the K-cycle answers the flexible solver with."""
    rng = np.random.default_rng(5)
    n, omega = 50, 0.85
    A = _rand(rng, n, n) + 4.0 * np.eye(n)
    rhs, z_mid = _rand(rng, n), _rand(rng, n)
    r2 = rhs - A @ z_mid
    z3, r = np.zeros(n, complex), r2.copy()
    for _ in range(2):
        q = A @ r
        alpha = omega * dot(q, r) / dot(q, q).real
        z3 += alpha * r
        r -= alpha * q
    z = z_mid + z3
    assert np.linalg.norm(A @ z - (rhs - r)) < 1e-13 * np.linalg.norm(rhs)


def test_bicgstab_lower_vector_replay():
    """qmg_bicgstab_replay: in the BiCG part of a BiCGstab(L) sweep only the top vectors r_j, u_j feed the applies and the dot
    products of step j, so the updates of the lower ones (i < j) and of x can be replayed after the last step from the top
    values the vectors held when they were on top."""
    rng = np.random.default_rng(7)
    n, L = 40, 6
    A = _rand(rng, n, n) / np.sqrt(n) + 2.0 * np.eye(n)
    rt = _rand(rng, n)
    r0, u0, x0 = _rand(rng, n), _rand(rng, n), _rand(rng, n)
    rho0, alpha = 1.3 - 0.2j, 0.4 + 0.1j

    def sweep(deferred):
        r = [r0.copy()] + [None] * L
        u = [u0.copy()] + [np.zeros(n, complex) for _ in range(L)]
        x = x0.copy()
        rho, al = rho0, alpha
        als, bes = [], []
        for j in range(L):
            rho1 = dot(rt, r[j])
            beta = al * rho1 / rho
            rho = rho1
            for i in range(j if deferred else 0, j + 1):
                u[i] = r[i] - beta * u[i]
            u[j + 1] = A @ u[j]
            al = rho / dot(rt, u[j + 1])
            for i in range(j if deferred else 0, j + 1):
                r[i] = r[i] - al * u[i + 1]
            r[j + 1] = A @ r[j]
            if not deferred:
                x = x + al * u[0]
            als.append(al); bes.append(beta)
        if deferred:
            for j in range(L):              # the replay: step j's updates of the vectors below j, then x
                for i in range(j):
                    u[i] = r[i] - bes[j] * u[i]
                for i in range(j):
                    r[i] = r[i] - als[j] * u[i + 1]
                x = x + als[j] * u[0]
        return r, u, x

    (ra, ua, xa), (rb, ub, xb) = sweep(False), sweep(True)
    for i in range(L + 1):
        assert np.allclose(ra[i], rb[i], rtol=0, atol=1e-12), i
        assert np.allclose(ua[i], ub[i], rtol=0, atol=1e-12), i
    assert np.allclose(xa, xb, rtol=0, atol=1e-12)


def test_right_looking_gram_schmidt_is_the_left_looking_one():
    """qmg_bicgstab_mgs: pass i subtracts tau_ij r_i from ALL later r_j, tau_ij = <r_i|r_j> / sigma_i taken of the r_j already
    updated against r_1 .. r_{i-1} -- the updates every r_j receives, in the order the left-looking loop applies them."""
    rng = np.random.default_rng(9)
    n, L = 50, 6
    R0 = [_rand(rng, n) for _ in range(L + 1)]
    left = [v.copy() for v in R0]
    sigma, tau, gp = {}, {}, {}
    for j in range(1, L + 1):
        for i in range(1, j):
            tau[(i, j)] = dot(left[i], left[j]) / sigma[i]
            left[j] = left[j] - tau[(i, j)] * left[i]
        sigma[j] = dot(left[j], left[j]).real
        gp[j] = dot(left[j], left[0]) / sigma[j]
    right = [v.copy() for v in R0]
    sums = []
    for i in range(L):                  # pass i: r_i (i >= 1) out of r_{i+1..L}, then the sums of the now final r_{i+1}
        if i >= 1:
            prev = sums[-1]
            for m, j in enumerate(range(i + 1, L + 1)):
                right[j] = right[j] - (prev[2 + m] / prev[0]) * right[i]
        y = right[i + 1]
        sums.append([dot(y, y).real, dot(y, right[0])] + [dot(y, right[j]) for j in range(i + 2, L + 1)])
    for j in range(1, L + 1):
        assert np.allclose(right[j], left[j], rtol=0, atol=1e-12)
        row = sums[j - 1]
        assert abs(row[0] - sigma[j]) < 1e-12 * sigma[j]
        assert abs(row[1] / row[0] - gp[j]) < 1e-12
        for m, k in enumerate(range(j + 1, L + 1)):
            assert abs(row[2 + m] / row[0] - tau[(j, k)]) < 1e-12


def test_chirality_packing_drops_only_zeros():
    """qmg_transfer_pack_chiral: with null vector j = upper and j + n/2 = lower chirality projection of one solve, the packed
    array holds everything a prolongation / restriction reads."""
    rng = np.random.default_rng(11)
    ncf, nvec, sites = 8, 8, 32
    nf, nvh = sites * ncf, nvec // 2
    comp = np.arange(nf) % ncf
    nv = np.zeros((nvec, nf), complex)
    for v in range(nvec):
        mask = (comp >= ncf // 2) == (v >= nvh)
        nv[v, mask] = _rand(rng, int(mask.sum()))
    h = (comp >= ncf // 2).astype(int)
    packed = np.stack([nv[h * nvh + i, np.arange(nf)] for i in range(nvh)], axis=1)
    coarse, fine = _rand(rng, nvec), _rand(rng, nf)
    prolong = (nv * coarse[:, None]).sum(axis=0)
    prolong_packed = np.array([sum(packed[e, i] * coarse[h[e] * nvh + i] for i in range(nvh)) for e in range(nf)])
    assert np.allclose(prolong_packed, prolong, rtol=0, atol=1e-13)
    restrict = np.array([np.vdot(nv[v], fine) for v in range(nvec)])
    restrict_packed = np.zeros(nvec, complex)
    for e in range(nf):
        for i in range(nvh):
            restrict_packed[h[e] * nvh + i] += np.conj(packed[e, i]) * fine[e]
    assert np.allclose(restrict_packed, restrict, rtol=0, atol=1e-12)
