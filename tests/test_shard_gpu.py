"""y-slab sharding on the GPU.

* loopback: one rank exchanging with itself, so every access across y = 0 / Y-1 -- the rhs rows of the stencil apply,
  the U_y / hopping / B^-1 rows of the link fills and variant builders, the prolong-vector rows of the Galerkin build --
  runs through the pack / exchange / halo-row path, and must reproduce the periodic single-GPU results BIT FOR BIT.
  (The whole `-m gpu` suite can also be run that way: QMG_LOOPBACK=1 python -m pytest tests -m gpu.)
* two ranks (needs 2 GPUs, else skipped): tests/shard_worker.py under torchrun: a 2-slab K-cycle solve against the
  single-GPU solve of the same lattice -- same null vectors (the device RNG is indexed by the global element), iteration
  counts +-1, solution to 1e-8 relative (two inexact inner solves in between)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import capi
import latutil

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host(t):
    return t.cpu().numpy()


@pytest.fixture()
def loop(qmg_gpu):
    """Callable running fn() once periodic and once in loopback; returns both results."""
    qmg = qmg_gpu

    def run(fn):
        if qmg.comm_counters()["active"]:
            pytest.skip("QMG_LOOPBACK already forced for the whole session")
        a = fn()
        qmg.comm_set_loopback(True)
        try:
            before = qmg.comm_counters()["halo_exchanges"]
            b = fn()
            assert qmg.comm_counters()["halo_exchanges"] > before, "loopback run did not exchange any halo row"
        finally:
            qmg.comm_set_loopback(False)
        return a, b
    return run


@pytest.mark.parametrize("kind", ["wilson", "staggered", "laplace", "dwf4"])
def test_loopback_fill_apply(qmg_gpu, loop, kind):
    qmg = qmg_gpu
    X, Y = 32, 16
    g = qmg.to_device(latutil.synthetic_gauge(X, Y, 6.0, 3))
    nc = {"wilson": 2, "staggered": 1, "laplace": 1, "dwf4": 8}[kind]
    rhs = qmg.to_device(latutil.gaussian_cv(X * Y * nc, 5))

    def fn():
        if kind == "wilson":
            cl, hp = qmg.fill_wilson(X, Y, g)
        elif kind == "staggered":
            cl, hp = None, qmg.fill_staggered(X, Y, g)
        elif kind == "laplace":
            cl, hp = qmg.fill_laplace(X, Y, g)
        else:
            cl, hp = qmg.fill_dwf(X, Y, 4, g, 0.05)
        d = qmg.stencil_desc(X, Y, nc, cl, hp, shift=0.1)
        out = [host(hp)]
        for pieces, dm in ((qmg.APPLY_ALL, 15), (qmg.APPLY_HOP_TO_EVEN | qmg.APPLY_EVEN_ROWS_ONLY, 15),
                           (qmg.APPLY_HOP_TO_ODD | qmg.APPLY_ODD_ROWS_ONLY, 10), (qmg.APPLY_ALL | qmg.APPLY_ACCUMULATE, 2)):
            lhs = qmg.cvec(X * Y * nc)
            lhs += 1.0
            qmg.stencil_apply(d, lhs, rhs, pieces, dm)
            out.append(host(lhs))
        hout = np.zeros(X * Y * nc, np.complex128)
        qmg.stencil_apply_host(d, hout, host(rhs), rows_per_chunk=5)      # pipelined host-vector entry, halo exchanged mid-upload
        out.append(hout)
        dot, nrm = qmg.stencil_apply_dot(d, qmg.cvec(X * Y * nc), rhs, rhs)
        out.append(np.array([dot.real, dot.imag, nrm]))
        return out

    a, b = loop(fn)
    for i, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(u, v), (kind, i)


def test_loopback_classes_and_kcycle(qmg_gpu, loop):
    """Host classes end to end: variant builders, Galerkin coarse operator from both link sets, 3-level K-cycle."""
    be = capi.Backend("gpu")
    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 9)

    def fn():
        lat = be.lattice(L, L, 2)
        op = lat.wilson(-0.03, g)
        op.build(dagger=True, rbjacobi=True, rbj_dagger=True)
        arrays = [op.get(n) for n in ("dagger_hopping", "rbjacobi_hopping", "rbjacobi_cinv", "rbj_dagger_hopping")]
        applies = [op.apply(b, t) for t in range(9)]
        op.free()
        kc = capi.KCycle(be, L, -0.03, g, n_refine=2, block=4, coarse_dof=8, seed=5)
        x, info = kc.solve(b, tol=1e-10, want_x=True)
        coarse = kc.tracker(1), kc.tracker(2)
        kc.free()
        return arrays, applies, x, info, coarse

    (arr0, app0, x0, i0, c0), (arr1, app1, x1, i1, c1) = loop(fn)
    for u, v in zip(arr0 + app0, arr1 + app1):
        assert np.array_equal(u, v)
    assert i0["iter"] == i1["iter"] and c0 == c1
    assert np.array_equal(x0, x1)


def test_loopback_link_compressed_kcycle(qmg_gpu, loop):
    """gamma5-hermitian applies on a slab: row -1 of the +y blocks comes from the lower rank (here: itself)."""
    be = capi.Backend("gpu")
    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 9)

    def fn():
        kc = capi.KCycle(be, L, -0.03, g, n_refine=2, seed=5)
        n = kc.gamma5_hermitian(True)
        x, info = kc.solve(b, tol=1e-10, want_x=True)
        kc.free()
        return n, x, info["iter"]
    (n0, x0, it0), (n1, x1, it1) = loop(fn)
    # (periodic: shared-memory tile kernel; slab: streaming kernel with the halo row -- different summation trees)
    assert n0 == n1 == 3 and it0 == it1 and latutil.rel_l2(x1, x0) < 1e-9


def test_loopback_matrix_free_wilson(qmg_gpu, loop):
    """Matrix-free Wilson apply on a slab: row -1 of U_y comes from the lower rank (here: itself) -- the applies and a K-cycle
    whose fine operator applies matrix-free have the bits of the periodic run."""
    be = capi.Backend("gpu")
    L = 64
    g = latutil.load_gauge(L)
    b = latutil.gaussian_cv(L * L * 2, 9)

    def fn():
        lat = be.lattice(L, L, 2)
        op = lat.wilson(-0.03, g)
        stored = op.apply(b, 0)
        on = op.matrix_free(g)
        free = op.apply(b, 0)
        op.free()
        was = be.fn("kcycle_setup_matrix_free")(1)
        try:
            kc = capi.KCycle(be, L, -0.03, g, n_refine=2, seed=5)
        finally:
            be.fn("kcycle_setup_matrix_free")(was)
        active = kc.matrix_free(True)
        x, info = kc.solve(b, tol=1e-10, want_x=True)
        kc.free()
        return on, active, stored, free, x, info["iter"]
    (on0, ac0, s0, f0, x0, it0), (on1, ac1, s1, f1, x1, it1) = loop(fn)
    assert on0 == on1 == 1 and ac0 == ac1 == 1
    assert np.array_equal(f0, s0) and np.array_equal(f1, s1) and np.array_equal(f1, f0)
    assert it0 == it1 and np.array_equal(x1, x0)


def test_two_rank_kcycle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, QMG_DEVICE_RNG="1")
    env.pop("QMG_LOOPBACK", None)          # the workers build a real two-rank communicator
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "shard_worker.py"), "--L", "256", "--levels", "3"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "SHARD-OK" in r.stdout, r.stdout[-4000:]
