"""CPU suite, part 3: the y-slab sharding logic of the multi-GPU path, without a GPU.
A numpy restatement of the stencil on ONE slab with explicit halo rows (the contract of qmg_stencil_desc.halo_ym/yp)
is checked against the oracle's periodic apply on the global lattice, single-process and with two gloo ranks that
exchange their boundary rows the way bench.py does over NCCL."""
import os

import numpy as np
import pytest

import capi
import latutil
import shard

pytestmark = pytest.mark.skipif(not capi.have_ref(), reason="oracle/_ref not built")


def slab_apply_numpy(X, Yl, nc, clover, hopping, shift, rhs, halo_ym, halo_yp):
    """out(x) = (clover + shift) in(x) + sum_mu hopping_mu(x) in(x+mu) on an X x Yl slab in its local eo layout;
    rows y=-1 / y=Yl come from halo_ym / halo_yp laid out (parity, x/2, c)."""
    xh, half, V = X // 2, X // 2 * Yl, X * Yl
    out = np.zeros(V * nc, np.complex128)
    cl = clover.reshape(V, nc, nc)
    hp = hopping.reshape(4, V, nc, nc)
    v = rhs.reshape(V, nc)
    hm, hpl = halo_ym.reshape(2, xh, nc), halo_yp.reshape(2, xh, nc)
    o = out.reshape(V, nc)
    for p in (0, 1):
        q = 1 - p
        for y in range(Yl):
            sft = (y + p) & 1
            for k in range(xh):
                s = p * half + y * xh + k
                acc = cl[s] @ v[s] + shift * v[s]
                kk = (k + sft) % xh
                acc += hp[0, s] @ v[q * half + y * xh + kk]
                kk = (k - 1 + sft) % xh
                acc += hp[2, s] @ v[q * half + y * xh + kk]
                acc += hp[1, s] @ (hpl[q, k] if y == Yl - 1 else v[q * half + (y + 1) * xh + k])
                acc += hp[3, s] @ (hm[q, k] if y == 0 else v[q * half + (y - 1) * xh + k])
                o[s] = acc
    return out


def _global_problem(L):
    ref = capi.Backend("ref")
    g = latutil.phases_to_gauge(np.random.default_rng(4).normal(0, 0.5, size=2 * L * L), L, L)
    op = ref.lattice(L, L, 2).wilson(-0.05, g)
    rhs = latutil.gaussian_cv(2 * L * L, 8)
    return op.get("clover"), op.get("hopping"), rhs, op.apply(rhs, 0)


def test_slab_reindex_roundtrip():
    L = 16
    f = latutil.gaussian_cv(L * L * 3, 1)
    back = np.zeros_like(f)
    for r in range(4):
        sl = shard.Slab(L, L, 4, r)
        back_local = sl.take(f, 3)
        assert back_local.size == f.size // 4
        sl.put(back, back_local, 3)
    assert np.array_equal(back, f)


@pytest.mark.parametrize("nranks", [1, 2, 4])
def test_slabs_with_halos_equal_periodic(nranks):
    L = 16
    cl, hp, rhs, want = _global_problem(L)
    got = np.zeros_like(want)
    V = L * L
    for r in range(nranks):
        sl = shard.Slab(L, L, nranks, r)
        lhp = np.concatenate([sl.take(hp[mu * V * 4:(mu + 1) * V * 4], 4) for mu in range(4)])
        out = slab_apply_numpy(L, sl.Yl, 2, sl.take(cl, 4), lhp, -0.05, sl.take(rhs, 2), sl.halo_row(rhs, 2, -1), sl.halo_row(rhs, 2, sl.Yl))
        sl.put(got, out, 2)
    assert latutil.rel_l2(got, want) < 1e-14


def _gloo_worker(rank, world, port, L, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cl, hp, rhs, want = _global_problem(L)
    sl = shard.Slab(L, L, world, rank)
    V, xh = L * L, L // 2
    local = sl.take(rhs, 2)
    half = xh * sl.Yl * 2
    # boundary rows of the LOCAL vector, packed (parity, x/2, c) exactly as bench.py's exchange() packs them
    lo = torch.from_numpy(np.concatenate([local[0:xh * 2], local[half:half + xh * 2]]))
    hi = torch.from_numpy(np.concatenate([local[(sl.Yl - 1) * xh * 2:sl.Yl * xh * 2], local[half + (sl.Yl - 1) * xh * 2:half + sl.Yl * xh * 2]]))
    ym, yp = torch.zeros_like(lo), torch.zeros_like(lo)
    up, down = (rank + 1) % world, (rank - 1) % world
    reqs = [dist.isend(hi, up), dist.irecv(ym, down), dist.isend(lo, down), dist.irecv(yp, up)]
    for r_ in reqs:
        r_.wait()
    assert np.array_equal(ym.numpy(), sl.halo_row(rhs, 2, -1)) and np.array_equal(yp.numpy(), sl.halo_row(rhs, 2, sl.Yl))
    lhp = np.concatenate([sl.take(hp[mu * V * 4:(mu + 1) * V * 4], 4) for mu in range(4)])
    out = slab_apply_numpy(L, sl.Yl, 2, sl.take(cl, 4), lhp, -0.05, local, ym.numpy(), yp.numpy())
    err = latutil.rel_l2(out, sl.take(want, 2))
    # a global reduction the way the sharded solvers do it: local partial + all-reduce
    part = torch.tensor([np.vdot(out, out).real])
    dist.all_reduce(part)
    q.put((rank, err, float(part.item()), float(np.vdot(want, want).real)))
    dist.destroy_process_group()


def test_two_rank_gloo_halo_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nrm, want in res:
        assert err < 1e-14
        assert abs(nrm - want) < 1e-10 * want


def test_stackable_synthetic_slabs_form_one_configuration():
    """bench.py's sharded runs let every rank draw its own slab (latutil.synthetic_phases(slab=True)); stacked in y the slabs
    must be ONE smooth periodic configuration: the plaquette across the seams is as gaussian as inside a slab."""
    import latutil
    X, Yl, N, beta = 32, 16, 4, 6.0
    parts = [latutil.synthetic_phases(X, Yl, beta, 100 + r, slab=True).reshape(X, Yl, 2) for r in range(N)]
    ph = np.concatenate(parts, axis=1)
    thx, thy = ph[:, :, 0], ph[:, :, 1]
    plaq = thx + np.roll(thy, -1, axis=0) - np.roll(thx, -1, axis=1) - thy
    plaq = (plaq + np.pi) % (2 * np.pi) - np.pi
    seams = plaq[:, Yl - 1::Yl]
    assert abs(plaq.std() - 1 / np.sqrt(beta)) < 0.03
    assert abs(seams.std() - 1 / np.sqrt(beta)) < 0.08
    g = latutil.phases_to_gauge(ph.ravel(), X, Yl * N)
    assert abs(latutil.average_plaquette(g, X, Yl * N) - np.exp(-0.5 / beta)) < 0.01


def test_unit_built_fields_are_one_field_for_every_slab_count():
    """bench.py builds its K-cycle fields from fixed units (latutil.synthetic_gauge_units): the same eight units cut 1, 2, 4
    or 8 ways are the same global field (strong scaling compares like with like), and rank 0 of a weak-scaling run holds the
    one-GPU field."""
    import latutil
    X, Y, unit = 16, 64, 8
    whole = latutil.synthetic_gauge_units(X, Y, unit, 0, seed=7)
    assert abs(latutil.average_plaquette(whole, X, Y) - np.exp(-0.5 / 6.0)) < 0.03
    for n in (2, 4, 8):
        Yl = Y // n
        rebuilt = np.zeros_like(whole)
        for r in range(n):
            part = latutil.synthetic_gauge_units(X, Yl, unit, r * (Yl // unit), seed=7)
            for mu in range(2):
                # a slab's rows y0 .. y0 + Yl of both parity halves of the global eo array (quantum-mg_b200/shard.py)
                for par in range(2):
                    loc = part[mu * X * Yl + par * (X // 2) * Yl: mu * X * Yl + (par + 1) * (X // 2) * Yl]
                    g0 = mu * X * Y + par * (X // 2) * Y + r * Yl * (X // 2)
                    rebuilt[g0:g0 + Yl * (X // 2)] = loc
        assert np.array_equal(rebuilt, whole), n
    # weak scaling: rank 0's slab of a taller lattice is the one-rank field
    assert np.array_equal(latutil.synthetic_gauge_units(X, Y, unit, 0, seed=7), whole)
    assert not np.array_equal(latutil.synthetic_gauge_units(X, Y, unit, Y // unit, seed=7), whole)


def test_global_index_rng_mapping_matches_slab_layout():
    """qmg_gaussian on a y-slab keys its counter by  p * (N half) + rank * half + i'  (csrc/qmg_blas.cu): that must be the
    position of local element i = p * half + i' in the global even-odd field, i.e. exactly what shard.Slab.take selects --
    so N slabs together draw the vector one GPU draws for the whole lattice."""
    X, Y, N = 8, 16, 4
    for dof in (1, 2, 8):
        glob = np.arange(X * Y * dof)
        for rank in range(N):
            sl = shard.Slab(X, Y, N, rank)
            local = sl.take(glob, dof)
            n = local.size
            half = n // 2
            i = np.arange(n)
            p = (i >= half).astype(np.int64)
            gi = p * N * half + rank * half + (i - p * half)
            assert np.array_equal(local, gi)
