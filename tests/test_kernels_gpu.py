"""Parity of the sm_100a kernels, called through the C ABI (include/qmg_b200.h), against the
oracle: the reference's own unmodified headers (oracle/_ref/libqmg_ref.so) and the committed
golden outputs.  Tolerance: 1e-12 relative L2 in fp64 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import capi
import latutil

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ref():
    if not capi.have_ref():
        pytest.skip("oracle/_ref/libqmg_ref.so not built")
    return capi.Backend("ref")


def dev(qmg, a):
    return qmg.to_device(a)


def host(t):
    return t.cpu().numpy()


def make_op(ref, qmg, kind, L):
    """Build the same operator through the oracle (class API) and through the fill kernels."""
    if kind == "wilson":
        lat = ref.lattice(L, L, 2)
        g = latutil.load_gauge(L)
        op = lat.wilson(-0.055, g)
        cl, hp = qmg.fill_wilson(L, L, dev(qmg, g))
        d = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055)
    elif kind == "staggered":
        lat = ref.lattice(L, L, 1)
        g = latutil.load_gauge(L)
        op = lat.staggered(0.1, g)
        cl, hp = None, qmg.fill_staggered(L, L, dev(qmg, g))
        d = qmg.stencil_desc(L, L, 1, None, hp, shift=0.1)
    elif kind == "laplace":
        lat = ref.lattice(L, L, 1)
        g = latutil.load_gauge(L)
        op = lat.laplace(0.01, g)
        cl, hp = qmg.fill_laplace(L, L, dev(qmg, g))
        d = qmg.stencil_desc(L, L, 1, cl, hp, shift=0.01)
    elif kind.startswith("dwf"):
        Ls = int(kind[3:])
        lat = ref.lattice(L, L, 2 * Ls)
        g = latutil.load_gauge(L)
        op = lat.dwf(0.05, g, Ls, -1.0)
        cl, hp = qmg.fill_dwf(L, L, Ls, dev(qmg, g), 0.05)
        d = qmg.stencil_desc(L, L, 2 * Ls, cl, hp, shift=op.shifts()[0])
    else:
        raise ValueError(kind)
    return lat, op, cl, hp, d


@pytest.mark.parametrize("kind,L", [("wilson", 64), ("wilson", 32), ("staggered", 32), ("laplace", 32), ("dwf2", 32), ("dwf4", 32), ("dwf6", 32)])
def test_fill_and_apply(ref, qmg_gpu, kind, L):
    qmg = qmg_gpu
    lat, op, cl, hp, d = make_op(ref, qmg, kind, L)
    if cl is not None:
        assert latutil.rel_l2(host(cl), op.get("clover")) < TOL
    assert latutil.rel_l2(host(hp), op.get("hopping")) < TOL
    rhs = latutil.gaussian_cv(lat.size_cv, 7)
    want = op.apply(rhs, 0)
    out = qmg.cvec(lat.size_cv)
    qmg.stencil_apply(d, out, dev(qmg, rhs))
    assert latutil.rel_l2(host(out), want) < TOL
    op.free()


def test_golden_n11(qmg_gpu):
    """Committed outputs of the reference on l64t64b60, mass -0.055 (tests/n11_wilson_test/wilson_test.cpp:37-43)."""
    qmg = qmg_gpu
    gold = np.load(os.path.join(latutil.GOLDEN, "golden_outputs.npz"))
    L = 64
    g = latutil.load_gauge(L)
    cl, hp = qmg.fill_wilson(L, L, dev(qmg, g))
    d = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055)
    n = L * L * 2
    rhs = latutil.gaussian_cv(n, int(gold["n11_wilson64_gauss_rhs_seed"][0]))
    out = qmg.cvec(n)
    qmg.stencil_apply(d, out, dev(qmg, rhs))
    assert latutil.rel_l2(host(out), gold["n11_wilson64_gauss_out"]) < TOL
    pt = np.zeros(n, np.complex128)
    pt[int(latutil.site_index(32, 32, L, L)) * 2] = 1.0
    qmg.stencil_apply(d, out, dev(qmg, pt))
    assert latutil.rel_l2(host(out), gold["n11_wilson64_point_out"]) < TOL


def random_stencil(L, nc, seed):
    rng = np.random.default_rng(seed)
    V = L * L
    cl = (rng.normal(size=V * nc * nc) + 1j * rng.normal(size=V * nc * nc))
    hp = (rng.normal(size=4 * V * nc * nc) + 1j * rng.normal(size=4 * V * nc * nc))
    # keep the site blocks well conditioned for the rbjacobi inverse
    cl = cl.reshape(V, nc, nc) + 4.0 * nc * np.eye(nc)[None]
    return cl.ravel().astype(np.complex128), hp.astype(np.complex128)


@pytest.mark.parametrize("nc", [1, 2, 4, 8, 6, 16])
def test_generic_stencil_pieces(ref, qmg_gpu, nc):
    """Coarse-operator-shaped dense blocks with all three shifts; every apply_M piece (stencil_2d.h:666-936)."""
    qmg = qmg_gpu
    L = 16
    lat = ref.lattice(L, L, nc)
    cl, hp = random_stencil(L, nc, 100 + nc)
    sh, eo, df = 0.3 - 0.1j, 0.05 + 0.02j, (0.07 - 0.03j if nc % 2 == 0 else 0.0)
    op = lat.generic(cl, hp, sh, eo, df)
    dcl, dhp = dev(qmg, cl), dev(qmg, hp)
    d = qmg.stencil_desc(L, L, nc, dcl, dhp, sh, eo, df)
    rhs = latutil.gaussian_cv(lat.size_cv, 5)
    drhs = dev(qmg, rhs)
    acc0 = latutil.gaussian_cv(lat.size_cv, 6)
    # full apply through the zeroing wrapper
    out = qmg.cvec(lat.size_cv)
    qmg.stencil_apply(d, out, drhs)
    assert latutil.rel_l2(host(out), op.apply(rhs, 0)) < TOL
    # accumulate flavour == Stencil2D::apply_M
    out = dev(qmg, acc0)
    qmg.stencil_apply(d, out, drhs, qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE)
    assert latutil.rel_l2(host(out), op.apply_piece(0, rhs, lhs=acc0)) < TOL
    # pieces: clover, eo, oe, hopping, shift
    for piece, flags in ((1, qmg.APPLY_CLOVER), (2, qmg.APPLY_HOP_TO_EVEN), (3, qmg.APPLY_HOP_TO_ODD),
                         (4, qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD), (6, qmg.APPLY_SHIFT)):
        out = dev(qmg, acc0)
        qmg.stencil_apply(d, out, drhs, flags | qmg.APPLY_ACCUMULATE)
        assert latutil.rel_l2(host(out), op.apply_piece(piece, rhs, lhs=acc0)) < TOL, piece
    # single directions (coarse.h probes)
    for mu in range(4):
        out = dev(qmg, acc0)
        qmg.stencil_apply(d, out, drhs, qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD | qmg.APPLY_ACCUMULATE, 1 << mu)
        assert latutil.rel_l2(host(out), op.apply_piece(5, rhs, dir=mu, lhs=acc0)) < TOL, mu
    op.free()


@pytest.mark.parametrize("nc", [2, 8])
def test_variant_builders(ref, qmg_gpu, nc):
    """build_dagger_stencil (stencil_2d.h:1080) and build_rbjacobi_stencil (:1452) against the reference."""
    import ctypes as C
    qmg = qmg_gpu
    L = 16
    lat = ref.lattice(L, L, nc)
    cl, hp = random_stencil(L, nc, 200 + nc)
    sh, eo, df = 0.3 - 0.1j, 0.05 + 0.02j, 0.07 - 0.03j
    op = lat.generic(cl, hp, sh, eo, df)
    op.build(dagger=True, rbjacobi=True)
    dcl, dhp = dev(qmg, cl), dev(qmg, hp)
    d = qmg.stencil_desc(L, L, nc, dcl, dhp, sh, eo, df)
    dag_cl, dag_hp = qmg.cvec(cl.size), qmg.cvec(hp.size)
    qmg.check(qmg.lib().qmg_build_dagger(L, L, nc, qmg.ptr(dcl), qmg.ptr(dhp), qmg.ptr(dag_cl), qmg.ptr(dag_hp)))
    assert latutil.rel_l2(host(dag_cl), op.get("dagger_clover")) < TOL
    assert latutil.rel_l2(host(dag_hp), op.get("dagger_hopping")) < TOL
    cinv, rcl, rhp = qmg.cvec(cl.size), qmg.cvec(cl.size), qmg.cvec(hp.size)
    qmg.check(qmg.lib().qmg_build_rbjacobi(C.byref(d), qmg.ptr(cinv), qmg.ptr(rcl), qmg.ptr(rhp)))
    assert latutil.rel_l2(host(cinv), op.get("rbjacobi_cinv")) < 1e-11
    assert latutil.rel_l2(host(rcl), op.get("rbjacobi_clover")) < TOL
    assert latutil.rel_l2(host(rhp), op.get("rbjacobi_hopping")) < 1e-11
    # rbjacobi apply: identity clover is not read (stencil_2d.h:1685)
    rhs = latutil.gaussian_cv(lat.size_cv, 9)
    d2 = qmg.stencil_desc(L, L, nc, None, rhp)
    out = qmg.cvec(lat.size_cv)
    qmg.stencil_apply(d2, out, dev(qmg, rhs), qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD | qmg.APPLY_IDENTITY_CLOVER)
    assert latutil.rel_l2(host(out), op.apply(rhs, 2)) < 1e-11
    op.free()


def test_cshift_known_answer(ref, qmg_gpu):
    """tests/n00_cshift: values = site number on a 6x4 lattice, all four directions, nc 1 and 2."""
    qmg = qmg_gpu
    X, Y = 6, 4
    for nc in (1, 2):
        lat = ref.lattice(X, Y, nc)
        xs, ys = latutil.site_coords(X, Y)
        v = np.repeat((ys * X + xs).astype(np.complex128), nc) + 1j * np.tile(np.arange(nc), X * Y)
        for cdir, (dx, dy) in ((2, (1, 0)), (3, (0, 1)), (4, (-1, 0)), (5, (0, -1))):
            out = qmg.cvec(v.size)
            qmg.check(qmg.lib().qmg_cshift(qmg.ptr(out), qmg.ptr(dev(qmg, v)), cdir, 3, nc, X, Y))
            got = host(out)
            want = np.repeat(((((ys + dy) % Y) * X + (xs + dx) % X)).astype(np.complex128), nc) + 1j * np.tile(np.arange(nc), X * Y)
            assert np.array_equal(got, want)
            assert np.array_equal(lat.cshift(v, cdir, 3, nc), want)   # the oracle agrees with the known answer


def test_halo_slabs_equal_periodic(qmg_gpu):
    """y-slab sharding: two half-height slabs fed with each other's boundary rows
    reproduce the single-lattice periodic apply bit for bit."""
    qmg = qmg_gpu
    L = 32
    g = latutil.load_gauge(L)
    cl, hp = qmg.fill_wilson(L, L, dev(qmg, g))
    d = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055)
    n = L * L * 2
    rhs = latutil.gaussian_cv(n, 3)
    full = qmg.cvec(n)
    qmg.stencil_apply(d, full, dev(qmg, rhs))
    full = host(full)
    import shard
    nr = 2
    pieces = []
    for r in range(nr):
        sl = shard.Slab(L, L, nr, r)
        lcl = dev(qmg, sl.take(host(cl), 4))
        lhp = np.concatenate([sl.take(host(hp)[mu * L * L * 4:(mu + 1) * L * L * 4], 4) for mu in range(4)])
        lrhs = sl.take(rhs, 2)
        ym = sl.halo_row(rhs, 2, -1)
        yp = sl.halo_row(rhs, 2, sl.Yl)
        dl = qmg.stencil_desc(L, sl.Yl, 2, lcl, dev(qmg, lhp), shift=-0.055, halo_ym=dev(qmg, ym), halo_yp=dev(qmg, yp))
        out = qmg.cvec(lrhs.size)
        qmg.stencil_apply(dl, out, dev(qmg, lrhs))
        pieces.append((sl, host(out)))
    got = np.zeros(n, np.complex128)
    for sl, o in pieces:
        sl.put(got, o, 2)
    assert np.array_equal(got, full)


@pytest.mark.parametrize("nc,chunk", [(2, 0), (2, 2), (2, 6), (2, 5), (8, 4), (1, 64)])
def test_host_vector_apply_pipelined(qmg_gpu, nc, chunk):
    """qmg_stencil_apply_host (upload / compute / download pipelined over row chunks) == the device-resident apply, bit for
    bit, for chunk sizes that do and do not divide Y, with and without accumulation."""
    qmg = qmg_gpu
    X, Y = 16, 24
    V = X * Y
    cl = latutil.gaussian_cv(V * nc * nc, 1)
    hp = latutil.gaussian_cv(4 * V * nc * nc, 2)
    d = qmg.stencil_desc(X, Y, nc, dev(qmg, cl), dev(qmg, hp), shift=0.2)
    rhs = latutil.gaussian_cv(V * nc, 3)
    old = latutil.gaussian_cv(V * nc, 4)
    for pieces in (qmg.APPLY_ALL, qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE):
        want = dev(qmg, old)
        qmg.stencil_apply(d, want, dev(qmg, rhs), pieces)
        got = old.copy()
        qmg.stencil_apply_host(d, got, rhs, pieces, rows_per_chunk=chunk)
        assert np.array_equal(got, host(want)), (nc, chunk, pieces)


def test_blas(qmg_gpu):
    import ctypes as C
    qmg = qmg_gpu
    n = 100003
    x, y = latutil.gaussian_cv(n, 1), latutil.gaussian_cv(n, 2)
    dx, dy = dev(qmg, x), dev(qmg, y)
    assert abs(qmg.dot(dx, dy) - np.vdot(x, y)) < 1e-12 * n
    assert abs(qmg.norm2sq(dx) - np.vdot(x, x).real) < 1e-12 * n
    a, b = 0.3 - 0.7j, -1.1 + 0.2j
    lib = qmg.lib()
    cd = C.c_double
    qmg.check(lib.qmg_caxpy(cd(a.real), cd(a.imag), qmg.ptr(dx), qmg.ptr(dy), C.c_long(n)))
    y = y + a * x
    assert latutil.rel_l2(host(dy), y) < 1e-15
    qmg.check(lib.qmg_caxpby(cd(a.real), cd(a.imag), qmg.ptr(dx), cd(b.real), cd(b.imag), qmg.ptr(dy), C.c_long(n)))
    y = a * x + b * y
    assert latutil.rel_l2(host(dy), y) < 1e-15
    # fused Krylov update: x += a p ; r -= a q ; |r|^2
    p_, q_, xx, rr = (latutil.gaussian_cv(n, s) for s in (3, 4, 5, 6))
    dp, dq, dxx, drr = (dev(qmg, v) for v in (p_, q_, xx, rr))
    out = C.c_double()
    qmg.check(lib.qmg_update_xr_norm(cd(a.real), cd(a.imag), qmg.ptr(dp), qmg.ptr(dq), qmg.ptr(dxx), qmg.ptr(drr), C.c_long(n), C.byref(out)))
    assert latutil.rel_l2(host(dxx), xx + a * p_) < 1e-15
    assert latutil.rel_l2(host(drr), rr - a * q_) < 1e-15
    assert abs(out.value - np.vdot(rr - a * q_, rr - a * q_).real) < 1e-12 * n
    # empty input: reductions return 0 (the reference's loops simply do not execute)
    qmg.check(lib.qmg_norm2sq(qmg.ptr(dx), C.c_long(0), C.byref(out)))
    assert out.value == 0.0
    # multi-dot / multi-axpy (GCR orthogonalisation)
    vs = [latutil.gaussian_cv(n, 10 + j) for j in range(11)]
    dvs = [dev(qmg, v) for v in vs]
    arr = (C.c_void_p * len(dvs))(*[t.data_ptr() for t in dvs])
    res = (C.c_double * (2 * len(dvs)))()
    qmg.check(lib.qmg_multi_dot(arr, len(dvs), qmg.ptr(dx), C.c_long(n), res))
    for j, v in enumerate(vs):
        assert abs(complex(res[2 * j], res[2 * j + 1]) - np.vdot(v, x)) < 1e-12 * n
    coef = latutil.gaussian_cv(len(vs), 77)
    ca = (C.c_double * (2 * len(vs)))(*np.stack([coef.real, coef.imag], 1).ravel())
    yy = latutil.gaussian_cv(n, 99)
    dyy = dev(qmg, yy)
    qmg.check(lib.qmg_multi_axpy(ca, arr, len(dvs), qmg.ptr(dyy), C.c_long(n)))
    assert latutil.rel_l2(host(dyy), yy + sum(c * v for c, v in zip(coef, vs))) < 1e-14


def test_gamma5_hermitian_apply(ref, qmg_gpu):
    """Link-compressed apply (B200 extension): for Wilson, and for its Galerkin coarsening with chirally doubled null
    vectors, the backward blocks equal gamma5 (forward block of the neighbour)^dag gamma5, the deviation check says so, and
    the apply that reads only clover / +x / +y blocks reproduces the stored-block apply (and the oracle) to rounding.
    A generic stencil fails the check."""
    qmg = qmg_gpu
    if qmg.comm_counters()["active"]:
        pytest.skip("kernel-level descriptors here carry no hop_halo_ym; the slab flavour is covered by test_shard_gpu.py")
    L = 32
    lat, op, cl, hp, d = make_op(ref, qmg, "wilson", L)
    assert qmg.stencil_gamma5_deviation(d) < 1e-15
    dh = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055, gamma5_hermitian=True)
    rhs = latutil.gaussian_cv(lat.size_cv, 7)
    want = op.apply(rhs, 0)
    out, outh = qmg.cvec(lat.size_cv), qmg.cvec(lat.size_cv)
    qmg.stencil_apply(d, out, dev(qmg, rhs))
    qmg.stencil_apply(dh, outh, dev(qmg, rhs))
    assert latutil.rel_l2(host(outh), want) < TOL and latutil.rel_l2(host(outh), host(out)) < 1e-14
    # pieces and directions go through the same code
    for pieces, dm in ((qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_EVEN_ROWS_ONLY, 15), (qmg.APPLY_HOP_TO_ODD | qmg.APPLY_ODD_ROWS_ONLY, 4),
                       (qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD, 8), (qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE, 15)):
        a, b = qmg.cvec(lat.size_cv) + 1.0, qmg.cvec(lat.size_cv) + 1.0
        qmg.stencil_apply(d, a, dev(qmg, rhs), pieces, dm)
        qmg.stencil_apply(dh, b, dev(qmg, rhs), pieces, dm)
        assert latutil.rel_l2(host(b), host(a)) < 1e-14, (pieces, dm)
    # Galerkin coarsening with projection-doubled null vectors (what the K-cycle builds): nc = 8
    Lc, ncc = 8, 8
    half = [latutil.gaussian_cv(lat.size_cv, 40 + v) for v in range(ncc // 2)]
    up = [v.copy() for v in half]
    dn = [v.copy() for v in half]
    for u_, d_ in zip(up, dn):
        u_[1::2] = 0.0          # chirality "up" = spin component 0
        d_[0::2] = 0.0
    nv = [dev(qmg, v) for v in up + dn]
    td = qmg.transfer_desc(L, L, 2, Lc, Lc, ncc)
    qmg.block_orthonormalize(td, nv)
    ccl, chp = qmg.coarse_build(td, d, nv)
    dc = qmg.stencil_desc(Lc, Lc, ncc, ccl, chp, shift=-0.055)
    assert qmg.stencil_gamma5_deviation(dc) < 1e-13
    dch = qmg.stencil_desc(Lc, Lc, ncc, ccl, chp, shift=-0.055, gamma5_hermitian=True)
    x = dev(qmg, latutil.gaussian_cv(Lc * Lc * ncc, 3))
    y0, y1 = qmg.cvec(Lc * Lc * ncc), qmg.cvec(Lc * Lc * ncc)
    qmg.stencil_apply(dc, y0, x)
    qmg.stencil_apply(dch, y1, x)
    assert latutil.rel_l2(host(y1), host(y0)) < 1e-13
    # a generic stencil does not obey the relation
    gcl, ghp = random_stencil(16, 4, 3)
    assert qmg.stencil_gamma5_deviation(qmg.stencil_desc(16, 16, 4, dev(qmg, gcl), dev(qmg, ghp))) > 0.5
    op.free()


@pytest.mark.parametrize("X,Y", [(64, 64), (32, 96), (16, 8), (4, 2)])
def test_wilson_matrix_free_apply_has_the_bits_of_the_stored_blocks(qmg_gpu, X, Y):
    """B200 extension: the whole-operator Wilson apply from the gauge links (qmg_stencil_desc.wilson_gauge: 96 instead of 384
    bytes per site) rebuilds the block elements with the fill's arithmetic and sums them in the element kernel's order -- plain,
    accumulating, with shifts on parity and chirality, with the residual epilogue: np.array_equal with the stored-block apply.
    Piece applies of the same descriptor keep reading the stored blocks; qmg_wilson_mf_deviation is 0 for untouched blocks and
    positive as soon as one element is edited."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    w = 0.9
    g = qmg.to_device(latutil.synthetic_gauge(X, Y, 6.0, 11))
    cl, hp = qmg.fill_wilson(X, Y, g, w)
    n = X * Y * 2
    x, b = dev(qmg, latutil.gaussian_cv(n, 1)), dev(qmg, latutil.gaussian_cv(n, 2))
    kw = dict(shift=-0.05 + 0.01j, eo_shift=0.02, dof_shift=0.003j)
    stored = qmg.stencil_desc(X, Y, 2, cl, hp, **kw)
    free = qmg.stencil_desc(X, Y, 2, cl, hp, wilson_gauge=g, wilson_w=w, **kw)
    assert qmg.wilson_mf_deviation(free) == 0.0
    launches = qmg.kernel_launches()
    for pieces in (qmg.APPLY_ALL, qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE, qmg.APPLY_ALL | qmg.APPLY_EVEN_ROWS_ONLY, qmg.APPLY_ALL | qmg.APPLY_ODD_ROWS_ONLY,
                   qmg.APPLY_CLOVER | qmg.APPLY_SHIFT, qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD):
        want, got = dev(qmg, latutil.gaussian_cv(n, 3)), dev(qmg, latutil.gaussian_cv(n, 3))
        qmg.stencil_apply(stored, want, x, pieces)
        qmg.stencil_apply(free, got, x, pieces)
        assert np.array_equal(host(got), host(want)), pieces
    want, got = qmg.cvec(n), qmg.cvec(n)
    qmg.check(lib.qmg_stencil_apply_residual(C.byref(stored), C.c_int(15), C.c_int(15), qmg.ptr(want), qmg.ptr(x), qmg.ptr(b)))
    qmg.check(lib.qmg_stencil_apply_residual(C.byref(free), C.c_int(15), C.c_int(15), qmg.ptr(got), qmg.ptr(x), qmg.ptr(b)))
    assert np.array_equal(host(got), host(want))
    assert qmg.kernel_launches() > launches
    # a single direction is a piece: stored blocks
    want, got = qmg.cvec(n), qmg.cvec(n)
    qmg.stencil_apply(stored, want, x, qmg.APPLY_ALL, 5)
    qmg.stencil_apply(free, got, x, qmg.APPLY_ALL, 5)
    assert np.array_equal(host(got), host(want))
    # the licence is withdrawn by any edit of the blocks (tests/n18_rbjacobi_stencil_test mutates the clover), or by the wrong w
    cl2 = cl.clone()
    cl2[3] += 1e-13
    assert qmg.wilson_mf_deviation(qmg.stencil_desc(X, Y, 2, cl2, hp, wilson_gauge=g, wilson_w=w)) > 0.0
    assert qmg.wilson_mf_deviation(qmg.stencil_desc(X, Y, 2, cl, hp, wilson_gauge=g, wilson_w=1.0)) > 0.0


@pytest.mark.parametrize("case", [(16, 16, 2, 4, 4, 8), (16, 16, 8, 4, 4, 8), (8, 8, 4, 2, 2, 4), (32, 16, 2, 8, 4, 2), (16, 32, 2, 8, 8, 8)])
def test_chirality_packed_transfer(qmg_gpu, case):
    """B200 extension: with chirality-doubled null vectors (vector j upper, j + ncc/2 lower components) the packed copy keeps the
    non-zero half of every fine element; restrict / prolong from it against the unpacked kernels (the restriction adds its lanes
    up in another order: 1e-14; the prolongation drops exact zeros only), the accumulate / overwrite / base flavours, and the
    refusal: vectors that are not chirally split report what packing would drop."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    Xf, Yf, ncf, Xc, Yc, ncc = case
    nf, ncv, nvh = Xf * Yf * ncf, Xc * Yc * ncc, ncc // 2
    t = qmg.TransferDesc(Xf, Yf, ncf, Xc, Yc, ncc)
    assert lib.qmg_transfer_packed_supported(C.byref(t)) == 1
    comp = np.arange(nf) % ncf
    vecs = []
    for v in range(ncc):
        a = latutil.gaussian_cv(nf, 300 + v)
        a[(comp >= ncf // 2) != (v >= nvh)] = 0.0       # vector v lives on the chirality half v // nvh
        vecs.append(dev(qmg, a))
    PP = C.c_void_p * ncc
    ptrs = PP(*[qmg.ptr(v) for v in vecs])
    packed = qmg.cvec(nf * nvh)
    dropped = C.c_double(-1.0)
    qmg.check(lib.qmg_transfer_pack_chiral(C.byref(t), ptrs, C.c_int(ncc), qmg.ptr(packed), C.byref(dropped)))
    assert dropped.value == 0.0
    fine, coarse, base = dev(qmg, latutil.gaussian_cv(nf, 1)), dev(qmg, latutil.gaussian_cv(ncv, 2)), dev(qmg, latutil.gaussian_cv(nf, 3))
    # restrict: overwrite and accumulate
    want, got = qmg.cvec(ncv), dev(qmg, latutil.gaussian_cv(ncv, 9))
    qmg.check(lib.qmg_restrict_overwrite(C.byref(t), ptrs, C.c_int(ncc), qmg.ptr(fine), qmg.ptr(want)))
    qmg.check(lib.qmg_restrict_packed(C.byref(t), qmg.ptr(packed), qmg.ptr(fine), qmg.ptr(got), C.c_int(1)))
    assert latutil.rel_l2(host(got), host(want)) < 1e-14
    want, got = dev(qmg, latutil.gaussian_cv(ncv, 9)), dev(qmg, latutil.gaussian_cv(ncv, 9))
    qmg.check(lib.qmg_restrict(C.byref(t), ptrs, C.c_int(ncc), qmg.ptr(fine), qmg.ptr(want)))
    qmg.check(lib.qmg_restrict_packed(C.byref(t), qmg.ptr(packed), qmg.ptr(fine), qmg.ptr(got), C.c_int(0)))
    assert latutil.rel_l2(host(got), host(want)) < 1e-14
    # zero + accumulate == overwrite, bit for bit (the fused K-cycle relies on it)
    z = qmg.cvec(ncv)
    qmg.check(lib.qmg_restrict_packed(C.byref(t), qmg.ptr(packed), qmg.ptr(fine), qmg.ptr(z), C.c_int(0)))
    o = dev(qmg, latutil.gaussian_cv(ncv, 9))
    qmg.check(lib.qmg_restrict_packed(C.byref(t), qmg.ptr(packed), qmg.ptr(fine), qmg.ptr(o), C.c_int(1)))
    assert np.array_equal(host(z), host(o))
    # prolong: accumulate, base, no base
    want, got = dev(qmg, latutil.gaussian_cv(nf, 4)), dev(qmg, latutil.gaussian_cv(nf, 4))
    qmg.check(lib.qmg_prolong(C.byref(t), ptrs, C.c_int(ncc), qmg.ptr(coarse), qmg.ptr(want)))
    qmg.check(lib.qmg_prolong_packed(C.byref(t), qmg.ptr(packed), qmg.ptr(coarse), None, qmg.ptr(got), C.c_int(0)))
    assert latutil.rel_l2(host(got), host(want)) < 1e-15
    for b in (base, None):
        want, got = qmg.cvec(nf), qmg.cvec(nf)
        qmg.check(lib.qmg_prolong_add(C.byref(t), ptrs, C.c_int(ncc), qmg.ptr(coarse), qmg.ptr(b) if b is not None else None, qmg.ptr(want)))
        qmg.check(lib.qmg_prolong_packed(C.byref(t), qmg.ptr(packed), qmg.ptr(coarse), qmg.ptr(b) if b is not None else None, qmg.ptr(got), C.c_int(1)))
        assert latutil.rel_l2(host(got), host(want)) < 1e-15
    # not chirally split: packing says how much it would drop
    vecs[0][1 if ncf == 2 else ncf // 2] = 1e-9
    qmg.check(lib.qmg_transfer_pack_chiral(C.byref(t), ptrs, C.c_int(ncc), qmg.ptr(packed), C.byref(dropped)))
    assert dropped.value > 0.0
    # odd blocks are not covered
    t_odd = qmg.TransferDesc(12, 12, 2, 4, 4, 8)
    assert lib.qmg_transfer_packed_supported(C.byref(t_odd)) == 0


def test_fused_krylov_step(qmg_gpu):
    """qmg_step_xr_norm (alpha formed on the device between two kernels, one host wait) == qmg_dot_norm + host alpha +
    qmg_update_xr_norm, bit for bit, including the MR aliasing p == r."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    for n in (5, 4096, 100003):
        for alias in (False, True):
            p0, q0, x0, r0 = (latutil.gaussian_cv(n, s) for s in (1, 2, 3, 4))
            omega = 0.85
            # reference sequence
            q, x, r = dev(qmg, q0), dev(qmg, x0), dev(qmg, r0)
            p = r if alias else dev(qmg, p0)
            d = (C.c_double * 3)()
            qmg.check(lib.qmg_dot_norm(qmg.ptr(q), qmg.ptr(r), C.c_long(n), d))
            alpha = omega * complex(d[0], d[1]) / d[2]
            rsq = C.c_double()
            qmg.check(lib.qmg_update_xr_norm(C.c_double(alpha.real), C.c_double(alpha.imag), qmg.ptr(p), qmg.ptr(q), qmg.ptr(x), qmg.ptr(r), C.c_long(n), C.byref(rsq)))
            want = (host(x), host(r), rsq.value, tuple(d))
            # fused step
            q, x, r = dev(qmg, q0), dev(qmg, x0), dev(qmg, r0)
            p = r if alias else dev(qmg, p0)
            out = (C.c_double * 4)()
            qmg.check(lib.qmg_step_xr_norm(C.c_double(omega), qmg.ptr(p), qmg.ptr(q), qmg.ptr(x), qmg.ptr(r), C.c_long(n), out))
            assert np.array_equal(host(x), want[0]) and np.array_equal(host(r), want[1]), (n, alias)
            assert out[0] == want[2] and (out[1], out[2], out[3]) == want[3], (n, alias)


@pytest.mark.parametrize("L", [1, 2, 3, 6, 7])
def test_bicgstab_fused_kernels(qmg_gpu, L):
    """The three kernels behind the fused BiCGstab(L) sweep against the BLAS calls they replace.
    qmg_bicgstab_replay: the lower-vector updates of the L BiCG steps (u_i = r_i - beta_j u_i, r_i -= alpha_j u_{i+1}, i < j,
    x += alpha_j u_0) == the same updates issued step by step with qmg_caxpby / qmg_caxpy, BIT FOR BIT;
    qmg_bicgstab_finish: x, r_0 == two qmg_multi_axpy calls bit for bit, |r_0|^2 to rounding;
    qmg_bicgstab_mgs: the right-looking Gram-Schmidt == the left-looking loop of dot / caxpy / dot_norm calls to rounding."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    lib.qmg_bicgstab_max_l.restype = C.c_int
    assert lib.qmg_bicgstab_max_l() >= 7
    PP = C.c_void_p * (L + 1)

    def ptrs(vs):
        return PP(*[qmg.ptr(v) for v in vs])

    def coef(cs):
        return (C.c_double * (2 * len(cs)))(*[t for c in cs for t in (c.real, c.imag)])

    for n in (7, 100003):
        rng = np.random.default_rng(17 * L + n)
        r0 = [latutil.gaussian_cv(n, 10 + i) for i in range(L + 1)]
        u0 = [latutil.gaussian_cv(n, 40 + i) for i in range(L + 1)]
        x0 = latutil.gaussian_cv(n, 99)
        al = [complex(a, b) for a, b in rng.normal(size=(L, 2))]
        be = [complex(a, b) for a, b in rng.normal(size=(L, 2))]
        # ---- replay.  Step by step: only the lower vectors (i < j) move here -- the solver has done the top ones itself
        r, u, x = [dev(qmg, v) for v in r0], [dev(qmg, v) for v in u0], dev(qmg, x0)
        for j in range(L):
            for i in range(j):
                qmg.check(lib.qmg_caxpby(C.c_double(1.0), C.c_double(0.0), qmg.ptr(r[i]), C.c_double(-be[j].real), C.c_double(-be[j].imag), qmg.ptr(u[i]), C.c_long(n)))
            for i in range(j):
                qmg.check(lib.qmg_caxpy(C.c_double(-al[j].real), C.c_double(-al[j].imag), qmg.ptr(u[i + 1]), qmg.ptr(r[i]), C.c_long(n)))
            qmg.check(lib.qmg_caxpy(C.c_double(al[j].real), C.c_double(al[j].imag), qmg.ptr(u[0]), qmg.ptr(x), C.c_long(n)))
        want = [host(v) for v in r] + [host(v) for v in u] + [host(x)]
        r, u, x = [dev(qmg, v) for v in r0], [dev(qmg, v) for v in u0], dev(qmg, x0)
        qmg.check(lib.qmg_bicgstab_replay(C.c_int(L), ptrs(r), ptrs(u), qmg.ptr(x), coef(al), coef(be), C.c_long(n)))
        got = [host(v) for v in r] + [host(v) for v in u] + [host(x)]
        for k, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a, b), (L, n, k)
        # ---- finish
        cx = [complex(a, b) for a, b in rng.normal(size=(L, 2))]
        cr = [complex(a, b) for a, b in rng.normal(size=(L, 2))]
        r, x = [dev(qmg, v) for v in r0], dev(qmg, x0)
        PL = C.c_void_p * L
        qmg.check(lib.qmg_multi_axpy(coef(cx), PL(*[qmg.ptr(v) for v in r[:L]]), C.c_int(L), qmg.ptr(x), C.c_long(n)))
        qmg.check(lib.qmg_multi_axpy(coef(cr), PL(*[qmg.ptr(v) for v in r[1:]]), C.c_int(L), qmg.ptr(r[0]), C.c_long(n)))
        wn = C.c_double()
        qmg.check(lib.qmg_norm2sq(qmg.ptr(r[0]), C.c_long(n), C.byref(wn)))
        wx, wr = host(x), host(r[0])
        r, x = [dev(qmg, v) for v in r0], dev(qmg, x0)
        gn = C.c_double()
        qmg.check(lib.qmg_bicgstab_finish(C.c_int(L), ptrs(r), qmg.ptr(x), coef(cx), coef(cr), C.c_long(n), C.byref(gn)))
        assert np.array_equal(host(x), wx) and np.array_equal(host(r[0]), wr), (L, n)
        assert abs(gn.value - wn.value) <= 1e-13 * wn.value
        # ---- Gram-Schmidt: left-looking loop of the call-by-call solver
        r = [dev(qmg, v) for v in r0]
        sigma, gp, tau = {}, {}, {}
        d2, d3 = (C.c_double * 2)(), (C.c_double * 3)()
        for j in range(1, L + 1):
            for i in range(1, j):
                qmg.check(lib.qmg_dot(qmg.ptr(r[i]), qmg.ptr(r[j]), C.c_long(n), d2))
                tau[(i, j)] = complex(d2[0], d2[1]) / sigma[i]
                qmg.check(lib.qmg_caxpy(C.c_double(-tau[(i, j)].real), C.c_double(-tau[(i, j)].imag), qmg.ptr(r[i]), qmg.ptr(r[j]), C.c_long(n)))
            qmg.check(lib.qmg_dot_norm(qmg.ptr(r[j]), qmg.ptr(r[0]), C.c_long(n), d3))
            sigma[j] = d3[2]
            gp[j] = complex(d3[0], d3[1]) / sigma[j]
        want = [host(v) for v in r]
        r = [dev(qmg, v) for v in r0]
        stride = 2 * lib.qmg_bicgstab_max_l() + 2
        sums = (C.c_double * (L * stride))()
        qmg.check(lib.qmg_bicgstab_mgs(C.c_int(L), ptrs(r), C.c_long(n), sums))
        for j in range(L + 1):
            assert latutil.rel_l2(host(r[j]), want[j]) < 1e-12, (L, n, j)
        for j in range(1, L + 1):
            row = sums[(j - 1) * stride:j * stride]
            assert abs(row[0] - sigma[j]) <= 1e-12 * sigma[j]
            assert abs(complex(row[1], row[2]) / row[0] - gp[j]) <= 1e-11 * (1 + abs(gp[j]))
            for m in range(j + 1, L + 1):
                t = complex(row[3 + 2 * (m - j - 1)], row[4 + 2 * (m - j - 1)]) / row[0]
                assert abs(t - tau[(j, m)]) <= 1e-11 * (1 + abs(tau[(j, m)])), (L, n, j, m)


def test_krylov_step_flavours(qmg_gpu):
    """qmg_krylov_step: without flags bit-identical to qmg_step_xr_norm; the zero-start first step (x_in NULL, r_in = b != r_out,
    |b|^2 on the side), the x-only last step and the accumulate-into flavour reproduce the separate sweeps bit for bit."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    NULL = C.c_void_p(0)
    for n in (5, 4096, 100003):
        p0, q0, x0, r0, a0 = (latutil.gaussian_cv(n, s) for s in (1, 2, 3, 4, 5))
        omega = 0.85
        # (a) no flags, MR aliasing p == r
        q, x, r = dev(qmg, q0), dev(qmg, x0), dev(qmg, r0)
        want4 = (C.c_double * 4)()
        qmg.check(lib.qmg_step_xr_norm(C.c_double(omega), qmg.ptr(r), qmg.ptr(q), qmg.ptr(x), qmg.ptr(r), C.c_long(n), want4))
        wx, wr = host(x), host(r)
        q, x, r = dev(qmg, q0), dev(qmg, x0), dev(qmg, r0)
        out = (C.c_double * 5)()
        qmg.check(lib.qmg_krylov_step(C.c_double(omega), qmg.ptr(r), qmg.ptr(q), qmg.ptr(x), qmg.ptr(x), qmg.ptr(r), qmg.ptr(r), NULL, C.c_long(n), 0, out, NULL))
        assert np.array_equal(host(x), wx) and np.array_equal(host(r), wr) and tuple(out)[:4] == tuple(want4), n
        # (b) first step from a zero start: x written, b read in place of r, |b|^2 returned == qmg_norm2sq(b)
        q, b = dev(qmg, q0), dev(qmg, r0)
        x, r = qmg.cvec(n), dev(qmg, r0)
        qmg.check(lib.qmg_step_xr_norm(C.c_double(omega), qmg.ptr(r), qmg.ptr(q), qmg.ptr(x), qmg.ptr(r), C.c_long(n), want4))
        wx, wr = host(x), host(r)
        x2 = dev(qmg, x0)          # garbage on entry: must not be read
        r2 = dev(qmg, a0)
        qmg.check(lib.qmg_krylov_step(C.c_double(omega), qmg.ptr(b), qmg.ptr(q), NULL, qmg.ptr(x2), qmg.ptr(b), qmg.ptr(r2), NULL, C.c_long(n), 1, out, NULL))
        assert np.array_equal(host(x2), wx) and np.array_equal(host(r2), wr) and tuple(out)[:4] == tuple(want4), n
        assert out[4] == qmg.norm2sq(b) and np.array_equal(host(b), r0), n
        # (c) x-only last step with lhs += z folded in: lhs + (x + alpha p)
        q, x, r, lhs = dev(qmg, q0), dev(qmg, x0), dev(qmg, r0), dev(qmg, a0)
        qmg.check(lib.qmg_step_xr_norm(C.c_double(omega), qmg.ptr(r), qmg.ptr(q), qmg.ptr(x), qmg.ptr(r), C.c_long(n), want4))
        qmg.check(lib.qmg_caxpy(C.c_double(1.0), C.c_double(0.0), qmg.ptr(x), qmg.ptr(lhs), C.c_long(n)))
        wl = host(lhs)
        q, x, r, lhs = dev(qmg, q0), dev(qmg, x0), dev(qmg, r0), dev(qmg, a0)
        qmg.check(lib.qmg_krylov_step(C.c_double(omega), qmg.ptr(r), qmg.ptr(q), qmg.ptr(x), qmg.ptr(lhs), qmg.ptr(r), qmg.ptr(r), qmg.ptr(lhs), C.c_long(n), 2, out, NULL))
        assert np.array_equal(host(lhs), wl) and np.array_equal(host(r), r0), n


@pytest.mark.parametrize("k", [1, 3, 8, 11])
def test_gcr_orthogonalize_on_device(qmg_gpu, k):
    """qmg_gcr_orthogonalize + qmg_krylov_step(DOTS_READY) == qmg_multi_dot + host division + 2 x qmg_multi_axpyz +
    qmg_krylov_step (coefficients formed on the device, both basis updates and the step's dot products in one launch;
    k = 11 takes two passes of the stored set)."""
    import ctypes as C
    import torch
    qmg = qmg_gpu
    lib = qmg.lib()
    NULL = C.c_void_p(0)
    n = 70001
    Ap0 = [latutil.gaussian_cv(n, 100 + j) for j in range(k)]
    p0 = [latutil.gaussian_cv(n, 200 + j) for j in range(k)]
    apk0, dir0, r0, x0 = (latutil.gaussian_cv(n, sd) for sd in (1, 2, 3, 4))
    apn = np.array([np.vdot(a, a).real for a in Ap0])

    def fresh():
        return [dev(qmg, a) for a in Ap0], [dev(qmg, a) for a in p0], dev(qmg, apk0), dev(qmg, dir0), qmg.cvec(n), dev(qmg, r0), dev(qmg, x0)
    # reference sequence (what gcr_core did before)
    Ap, P, apk, dirv, pk, r, x = fresh()
    arrA, arrP = (C.c_void_p * k)(*[t.data_ptr() for t in Ap]), (C.c_void_p * k)(*[t.data_ptr() for t in P])
    dots = (C.c_double * (2 * k))()
    qmg.check(lib.qmg_multi_dot(arrA, k, qmg.ptr(apk), C.c_long(n), dots))
    beta = (C.c_double * (2 * k))(*[-dots[i] / apn[i // 2] for i in range(2 * k)])
    qmg.check(lib.qmg_multi_axpyz(beta, arrA, k, qmg.ptr(apk), qmg.ptr(apk), C.c_long(n)))
    qmg.check(lib.qmg_multi_axpyz(beta, arrP, k, qmg.ptr(dirv), qmg.ptr(pk), C.c_long(n)))
    want5 = (C.c_double * 5)()
    qmg.check(lib.qmg_krylov_step(C.c_double(1.0), qmg.ptr(pk), qmg.ptr(apk), qmg.ptr(x), qmg.ptr(x), qmg.ptr(r), qmg.ptr(r), NULL, C.c_long(n), 0, want5, NULL))
    want = [host(t) for t in (apk, pk, x, r)]
    # device-side orthogonalisation
    Ap, P, apk, dirv, pk, r, x = fresh()
    arrA, arrP = (C.c_void_p * k)(*[t.data_ptr() for t in Ap]), (C.c_void_p * k)(*[t.data_ptr() for t in P])
    d_dots = torch.zeros(2 * k, dtype=torch.float64, device="cuda")
    d_apn = torch.tensor(np.concatenate([apn, [0.0]]), dtype=torch.float64, device="cuda")
    qmg.check(lib.qmg_gcr_orthogonalize(arrA, arrP, k, qmg.ptr(apk), qmg.ptr(dirv), qmg.ptr(pk), qmg.ptr(r), C.c_long(n), qmg.ptr(d_dots), qmg.ptr(d_apn)))
    got5 = (C.c_double * 5)()
    qq = C.c_void_p(d_apn.data_ptr() + 8 * k)
    qmg.check(lib.qmg_krylov_step(C.c_double(1.0), qmg.ptr(pk), qmg.ptr(apk), qmg.ptr(x), qmg.ptr(x), qmg.ptr(r), qmg.ptr(r), NULL, C.c_long(n), 4, got5, qq))
    got = [host(t) for t in (apk, pk, x, r)]
    # the basis updates are element-wise: same bits (coefficients formed on the device by the same IEEE division)
    assert np.allclose(d_dots.cpu().numpy(), np.array(list(dots)), rtol=1e-13, atol=1e-13 * n)
    if np.array_equal(d_dots.cpu().numpy(), np.array(list(dots))):
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    # the reductions are sized to the blocks resident per SM, which differs between the kernels: same sums, different trees
    for u, v in zip(got, want):
        assert latutil.rel_l2(u, v) < 1e-13
    assert np.allclose(np.array(tuple(got5)[:4]), np.array(tuple(want5)[:4]), rtol=1e-12, atol=1e-12 * n)
    assert abs(float(d_apn[k].item()) - want5[3]) <= 1e-12 * want5[3]


@pytest.mark.parametrize("nc,herm", [(2, False), (8, False), (8, True), (1, False), (6, False)])
def test_apply_residual_epilogue(qmg_gpu, nc, herm):
    """qmg_stencil_apply_residual == qmg_stencil_apply + qmg_caxpbyz(1, b, -1, A x), bit for bit, on the streaming, tile
    (gamma5-hermitian nc = 8) and generic kernels; lhs may alias b."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    L = 32
    n = L * L * nc
    if herm:
        cl, hop = _herm_stencil(L, nc, 5)
        clover, hopping = dev(qmg, cl), dev(qmg, hop)
        d = qmg.stencil_desc(L, L, nc, clover, hopping, shift=0.3, gamma5_hermitian=True)
        assert qmg.stencil_gamma5_deviation(d) < 1e-14
    else:
        cl, hp = random_stencil(L, nc, 3)
        clover, hopping = dev(qmg, cl), dev(qmg, hp)
        d = qmg.stencil_desc(L, L, nc, clover, hopping, shift=0.3 + 0.1j, eo_shift=0.05, dof_shift=(0.02 if nc % 2 == 0 else 0.0))
    x, b = dev(qmg, latutil.gaussian_cv(n, 8)), dev(qmg, latutil.gaussian_cv(n, 9))
    Ax, want = qmg.cvec(n), qmg.cvec(n)
    qmg.stencil_apply(d, Ax, x)
    qmg.check(lib.qmg_caxpbyz(C.c_double(1.0), C.c_double(0.0), qmg.ptr(b), C.c_double(-1.0), C.c_double(0.0), qmg.ptr(Ax), qmg.ptr(want), C.c_long(n)))
    got = qmg.cvec(n)
    qmg.check(lib.qmg_stencil_apply_residual(C.byref(d), C.c_int(15), C.c_int(15), qmg.ptr(got), qmg.ptr(x), qmg.ptr(b)))
    assert np.array_equal(host(got), host(want))
    b2 = b.clone()
    qmg.check(lib.qmg_stencil_apply_residual(C.byref(d), C.c_int(15), C.c_int(15), qmg.ptr(b2), qmg.ptr(x), qmg.ptr(b2)))
    assert np.array_equal(host(b2), host(want))
    assert lib.qmg_stencil_apply_residual(C.byref(d), C.c_int(15), C.c_int(15), qmg.ptr(x), qmg.ptr(x), qmg.ptr(b)) != 0   # lhs == rhs refused


def _herm_stencil(L, nc, seed):
    """A random gamma5-hermitian link set: forward blocks random, backward blocks s s conj(forward of the neighbour)^T."""
    rng = np.random.default_rng(seed)
    V = L * L
    hop = rng.normal(size=(4, V, nc, nc)) + 1j * rng.normal(size=(4, V, nc, nc))
    s = np.where(np.arange(nc) < nc // 2, 1.0, -1.0)
    x_, y_ = latutil.site_coords(L, L)
    for mu, (dx, dy) in ((2, (-1, 0)), (3, (0, -1))):
        nb = latutil.site_index((x_ + dx) % L, (y_ + dy) % L, L, L)
        hop[mu] = (s[None, :, None] * s[None, None, :]) * np.conj(np.swapaxes(hop[mu - 2][nb], 1, 2))
    cl = rng.normal(size=(V, nc, nc)) + 1j * rng.normal(size=(V, nc, nc))
    return cl.reshape(-1), hop.reshape(-1)


@pytest.mark.parametrize("L", [16, 64, 128])
def test_tile_kernel_flavours_agree(qmg_gpu, L):
    """nc = 8 link-compressed apply: streaming (0), cp.async patch kernel with two / one thread per column (1, 2) and the
    TMA-staged patch kernel (4: cp.async.bulk + mbarrier), the 16-site patch experiments (5, 6) and the persistent
    warp-specialised ring kernel (9: producer warp + mbarrier ring, loop-invariant consumers) against the stored-block apply
    -- plain, accumulating and with the residual epilogue; L = 16 makes every patch touch the periodic wrap in x, L = 128 makes
    every CTA of the ring kernel go round its ring."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    nc = 8
    n = L * L * nc
    cl, hop = _herm_stencil(L, nc, 31)
    clover, hopping = dev(qmg, cl), dev(qmg, hop)
    stored = qmg.stencil_desc(L, L, nc, clover, hopping, shift=0.2, dof_shift=0.03)
    herm = qmg.stencil_desc(L, L, nc, clover, hopping, shift=0.2, dof_shift=0.03, gamma5_hermitian=True)
    assert qmg.stencil_gamma5_deviation(stored) < 1e-14
    x, b, acc0 = (dev(qmg, latutil.gaussian_cv(n, sd)) for sd in (1, 2, 3))
    want = qmg.cvec(n)
    qmg.stencil_apply(stored, want, x)
    want_acc = acc0.clone()
    qmg.stencil_apply(stored, want_acc, x, qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE)
    want_res = qmg.cvec(n)
    qmg.check(lib.qmg_stencil_apply_residual(C.byref(stored), C.c_int(15), C.c_int(15), qmg.ptr(want_res), qmg.ptr(x), qmg.ptr(b)))
    old = lib.qmg_get_tile_kernel()
    try:
        for mode in (0, 1, 2, 3, 4, 5, 6, 9):
            qmg.check(lib.qmg_set_tile_kernel(mode))
            got = qmg.cvec(n)
            qmg.stencil_apply(herm, got, x)
            assert latutil.rel_l2(host(got), host(want)) < 1e-14, mode
            got = acc0.clone()
            qmg.stencil_apply(herm, got, x, qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE)
            assert latutil.rel_l2(host(got), host(want_acc)) < 1e-14, mode
            got = qmg.cvec(n)
            qmg.check(lib.qmg_stencil_apply_residual(C.byref(herm), C.c_int(15), C.c_int(15), qmg.ptr(got), qmg.ptr(x), qmg.ptr(b)))
            assert latutil.rel_l2(host(got), host(want_res)) < 1e-14, mode
    finally:
        qmg.check(lib.qmg_set_tile_kernel(old))


@pytest.mark.parametrize("nc", [1, 2, 8, 12])
def test_batched_qr_pair(qmg_gpu, nc):
    """quantum-linalg's cMATx_do_qr_square / cMATqr_do_xinv_square pair (stencil/stencil_2d.h:1536-1537): M = Q R with Q unitary
    and R upper triangular with a positive real diagonal (the unique factorisation, so numpy's is the reference), and
    Minv = R^-1 Q^dag."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    ns = 257
    rng = np.random.default_rng(nc)
    M = rng.normal(size=(ns, nc, nc)) + 1j * rng.normal(size=(ns, nc, nc)) + 2.0 * nc * np.eye(nc)[None]
    dM = dev(qmg, M.reshape(-1))
    dQ, dR, dI = qmg.cvec(ns * nc * nc), qmg.cvec(ns * nc * nc), qmg.cvec(ns * nc * nc)
    qmg.check(lib.qmg_cmat_qr(qmg.ptr(dM), qmg.ptr(dQ), qmg.ptr(dR), C.c_long(ns), C.c_int(nc)))
    qmg.check(lib.qmg_cmat_qr_inverse(qmg.ptr(dQ), qmg.ptr(dR), qmg.ptr(dI), C.c_long(ns), C.c_int(nc)))
    Q, R, Minv = (host(t).reshape(ns, nc, nc) for t in (dQ, dR, dI))
    eye = np.eye(nc)[None]
    assert np.abs(np.conj(np.swapaxes(Q, 1, 2)) @ Q - eye).max() < 1e-13
    assert np.abs(np.tril(R, -1)).max() == 0.0 and (np.abs(np.imag(np.diagonal(R, axis1=1, axis2=2))).max() == 0.0) and (np.real(np.diagonal(R, axis1=1, axis2=2)) > 0).all()
    assert np.abs(Q @ R - M).max() < 1e-12 * np.abs(M).max()
    assert np.abs(Minv @ M - eye).max() < 1e-12
    Qn, Rn = np.linalg.qr(M)
    ph = np.sign(np.real(np.diagonal(Rn, axis1=1, axis2=2))) if nc == 0 else (np.diagonal(Rn, axis1=1, axis2=2) / np.abs(np.diagonal(Rn, axis1=1, axis2=2)))
    assert np.abs(Qn * ph[:, None, :] - Q).max() < 1e-12
    assert np.abs(np.conj(ph)[:, :, None] * Rn - R).max() < 1e-12 * np.abs(M).max()


@pytest.mark.parametrize("L", [32, 64])
def test_fine_level_link_compressed_apply(ref, qmg_gpu, L):
    """nc = 2 (Wilson fine level): the link-compressed apply (clover, +x, +y blocks only; backward hops from the neighbours'
    forward blocks) against the stored-block apply and the oracle, on the reference's own configs -- plain, accumulating, with
    the residual epilogue and with shifts."""
    import ctypes as C
    qmg = qmg_gpu
    lib = qmg.lib()
    g = latutil.load_gauge(L)
    cl, hp = qmg.fill_wilson(L, L, dev(qmg, g))
    n = L * L * 2
    stored = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055, dof_shift=0.01)
    herm = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055, dof_shift=0.01, gamma5_hermitian=True)
    assert qmg.stencil_gamma5_deviation(stored) < 1e-14
    x0, b0, a0 = (latutil.gaussian_cv(n, sd) for sd in (1, 2, 3))
    x, b = dev(qmg, x0), dev(qmg, b0)
    op = ref.lattice(L, L, 2).wilson(-0.055, g)
    want_ref = op.apply(x0, 0)
    op.free()
    want = qmg.cvec(n)
    qmg.stencil_apply(stored, want, x)
    old = lib.qmg_get_tile_kernel()
    try:
        for mode in (1, 0):
            qmg.check(lib.qmg_set_tile_kernel(mode))
            got = qmg.cvec(n)
            qmg.stencil_apply(herm, got, x)
            assert latutil.rel_l2(host(got), host(want)) < 1e-14, mode
            acc = dev(qmg, a0)
            qmg.stencil_apply(herm, acc, x, qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE)
            assert latutil.rel_l2(host(acc), host(want) + a0) < 1e-14, mode
            res = qmg.cvec(n)
            qmg.check(lib.qmg_stencil_apply_residual(C.byref(herm), C.c_int(15), C.c_int(15), qmg.ptr(res), qmg.ptr(x), qmg.ptr(b)))
            assert latutil.rel_l2(host(res), b0 - host(want)) < 1e-14, mode
    finally:
        qmg.check(lib.qmg_set_tile_kernel(old))
    # against the oracle: the descriptor's dof_shift is not part of Wilson2D, so compare without it
    herm0 = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055, gamma5_hermitian=True)
    got = qmg.cvec(n)
    qmg.stencil_apply(herm0, got, x)
    assert latutil.rel_l2(host(got), want_ref) < TOL


def test_in_place_hopping_is_sequential(ref, qmg_gpu):
    """apply_M_hopping(x, x): the reference runs apply_M_eo, then apply_M_oe on the UPDATED even rows
    (stencil/stencil_2d.h:843-850); the in-place both-parity request is two ordered launches with that meaning."""
    qmg = qmg_gpu
    L, nc = 16, 2
    cl, hp = random_stencil(L, nc, 11)
    lat = ref.lattice(L, L, nc)
    op = lat.generic(cl, hp)
    x0 = latutil.gaussian_cv(L * L * nc, 12)
    y = op.apply_piece(2, x0, lhs=x0)             # apply_M_eo: even rows += H x_odd
    want = op.apply_piece(3, y, lhs=y)            # apply_M_oe on the updated vector: odd rows += H y_even
    op.free()
    d = qmg.stencil_desc(L, L, nc, dev(qmg, cl), dev(qmg, hp))
    x = dev(qmg, x0)
    qmg.stencil_apply(d, x, x, pieces=qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_HOP_TO_ODD | qmg.APPLY_ACCUMULATE)
    assert latutil.rel_l2(host(x), want) < TOL


def test_fused_apply_dot(ref, qmg_gpu):
    qmg = qmg_gpu
    L = 64
    g = latutil.load_gauge(L)
    cl, hp = qmg.fill_wilson(L, L, dev(qmg, g))
    d = qmg.stencil_desc(L, L, 2, cl, hp, shift=-0.055)
    n = L * L * 2
    rhs = latutil.gaussian_cv(n, 21)
    drhs = dev(qmg, rhs)
    out = qmg.cvec(n)
    dotv, nrm = qmg.stencil_apply_dot(d, out, drhs, drhs)
    o = host(out)
    assert abs(nrm - np.vdot(o, o).real) < 1e-12 * nrm
    assert abs(dotv - np.vdot(o, rhs)) < 1e-12 * abs(np.vdot(o, rhs))


# ------------------------------------------------------------------ transfer / coarse operator --
TRANSFER_CASES = [
    # (fine X, fine Y, ncf, coarse X, coarse Y, ncc)
    (16, 16, 2, 4, 4, 8),     # Wilson level 0 -> 1: 4x4 blocks, 32 fine dof per aggregate
    (8, 8, 8, 2, 2, 8),       # level 1 -> 2: 128 fine dof per aggregate
    (8, 8, 1, 4, 4, 2),       # 2x2 blocks of a scalar field (n07-style), 4 dof per aggregate
    (4, 4, 2, 1, 1, 6),       # n05: 4x4 -> a single coarse site, 6 null vectors
    (12, 6, 1, 4, 2, 3),      # odd block size (3x3), nvec not a power of two
    (16, 8, 2, 4, 2, 12),     # more than 8 null vectors: two passes
]


def _transfer_pair(ref, qmg, case, seed=0, block_ortho=True):
    Xf, Yf, ncf, Xc, Yc, ncc = case
    fl, cl = ref.lattice(Xf, Yf, ncf), ref.lattice(Xc, Yc, ncc)
    nv = np.stack([latutil.gaussian_cv(fl.size_cv, seed + 50 + v) for v in range(ncc)])
    tr = capi.Transfer(fl, cl, nv, block_ortho=block_ortho, save_decomp=block_ortho)
    td = qmg.transfer_desc(Xf, Yf, ncf, Xc, Yc, ncc)
    dnv = [dev(qmg, nv[v]) for v in range(ncc)]
    return fl, cl, nv, tr, td, dnv


@pytest.mark.parametrize("case", TRANSFER_CASES)
def test_block_ortho_prolong_restrict(ref, qmg_gpu, case):
    """transfer.h:514-607 (run twice, :160-173), :455-511; identities of tests/n05_prolong_restrict_test (:85-103)."""
    qmg = qmg_gpu
    fl, cl, nv, tr, td, dnv = _transfer_pair(ref, qmg, case)
    chol = qmg.cvec(cl.size_cm)
    qmg.block_orthonormalize(td, dnv, chol)
    qmg.block_orthonormalize(td, dnv, None)
    want = tr.nullvecs()
    for v in range(cl.nc):
        assert latutil.rel_l2(host(dnv[v]), want[v]) < 1e-11, v
    assert latutil.rel_l2(host(chol), tr.cholesky()) < 1e-11
    # P and R = P^dagger, accumulating into non-zero destinations
    cvec, fvec = latutil.gaussian_cv(cl.size_cv, 1), latutil.gaussian_cv(fl.size_cv, 2)
    f0, c0 = latutil.gaussian_cv(fl.size_cv, 3), latutil.gaussian_cv(cl.size_cv, 4)
    df = dev(qmg, f0)
    qmg.prolong(td, dnv, dev(qmg, cvec), df)
    assert latutil.rel_l2(host(df), tr.prolong(cvec, f0)) < 1e-11
    dc = dev(qmg, c0)
    qmg.restrict(td, dnv, dev(qmg, fvec), dc)
    assert latutil.rel_l2(host(dc), tr.restrict(fvec, c0)) < 1e-11
    # n05 identities: (1 - P^dag P) v_c = 0 and (1 - P P^dag) on the span of the null vectors
    df = qmg.cvec(fl.size_cv)
    qmg.prolong(td, dnv, dev(qmg, cvec), df)
    dc = qmg.cvec(cl.size_cv)
    qmg.restrict(td, dnv, df, dc)
    assert latutil.rel_l2(host(dc), cvec) < 1e-12
    tr.free()


@pytest.mark.parametrize("case,kind", [(TRANSFER_CASES[0], "wilson"), (TRANSFER_CASES[1], "generic"), (TRANSFER_CASES[2], "laplace"),
                                        (TRANSFER_CASES[3], "wilson"), ((8, 8, 1, 2, 2, 4), "staggered"), ((8, 4, 2, 2, 2, 4), "generic")])
def test_coarse_build(ref, qmg_gpu, case, kind):
    """CoarseOperator2D build (coarse.h:90-471) == direct Galerkin R A P; tests/n08_distance1_build_test."""
    qmg = qmg_gpu
    Xf, Yf, ncf, Xc, Yc, ncc = case
    fl, cl, nv, tr, td, dnv = _transfer_pair(ref, qmg, case, seed=7)
    rng = np.random.default_rng(5)
    ph = rng.normal(0, 0.4, size=Xf * Yf * 2)
    g = latutil.phases_to_gauge(ph, Xf, Yf)
    if kind == "wilson":
        op = fl.wilson(-0.05, g)
    elif kind == "laplace":
        op = fl.laplace(0.1, g)
    elif kind == "staggered":
        op = fl.staggered(0.1, g)
    else:
        c_, h_ = random_stencil(Xf, ncf, 31) if Xf == Yf else (None, None)
        if c_ is None:
            V = Xf * Yf
            c_ = latutil.gaussian_cv(V * ncf * ncf, 41)
            h_ = latutil.gaussian_cv(4 * V * ncf * ncf, 42)
        op = fl.generic(c_, h_, 0.2, 0.0, 0.0)
    fcl, fhp = op.get("clover"), op.get("hopping")
    dcl = None if fcl is None else dev(qmg, fcl)
    dhp = None if fhp is None else dev(qmg, fhp)
    fd = qmg.stencil_desc(Xf, Yf, ncf, dcl, dhp)
    want = tr.coarse_operator(op, is_chiral=False)
    onv = [dev(qmg, v) for v in tr.nullvecs()]
    ccl, chp = qmg.coarse_build(td, fd, onv)
    assert latutil.rel_l2(host(ccl), want.get("clover")) < 1e-11
    wh = want.get("hopping")
    assert np.linalg.norm(host(chp) - wh) < 1e-11 * max(1.0, np.linalg.norm(wh))
    want.free(); tr.free(); op.free()
