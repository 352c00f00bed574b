"""TEST DRIVER (run under torchrun, one process per GPU): the K-cycle on N y-slabs against the same solve on one GPU.

Phase A: every rank solves the WHOLE lattice on its own GPU (no communicator).  Phase B: qmg.comm_init, every rank
builds the hierarchy on its (X, Y/N) slab of the same gauge field and solves the same right-hand side.  Checks: outer
iteration count +-1, per-level operator counts within 5 %, and the slab of the phase-A solution to 1e-8 relative."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "quantum-mg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=256)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--mass", type=float, default=-0.03)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--hermitian", action="store_true", help="link-compressed (gamma5-hermitian) applies on every level")
    a = ap.parse_args()
    os.environ["QMG_DEVICE_RNG"] = "1"
    import torch
    import torch.distributed as dist
    import capi
    import latutil
    import qmg
    import shard
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    qmg.init(local)
    be = capi.Backend("gpu")
    L = a.L
    g = latutil.synthetic_gauge(L, L, 6.0, 11)
    b = latutil.gaussian_cv(L * L * 2, 21)
    kw = dict(n_refine=a.levels - 1, block=4, coarse_dof=8, seed=5)

    kc = capi.KCycle(be, L, a.mass, g, **kw)
    if a.hermitian:
        assert kc.gamma5_hermitian(True) == a.levels
    x_one, info_one = kc.solve(b, tol=a.tol, want_x=True)
    ops_one = [kc.tracker(l)["total"] for l in range(a.levels)]
    warm_one = kc.solve(b, tol=a.tol)["seconds"]          # second solve: warm allocator
    prec_one = kc.time_precond(1, 3)
    kc.free()

    qmg.comm_init()
    sl = shard.Slab(L, L, world, rank)
    V = L * L
    g_loc = np.concatenate([sl.take(g[:V], 1), sl.take(g[V:], 1)])
    kc = capi.KCycle(be, L, a.mass, g_loc, Y=sl.Yl, **kw)
    if a.hermitian:
        assert kc.gamma5_hermitian(True) == a.levels
    x_loc, info = kc.solve(sl.take(b, 2), tol=a.tol, want_x=True)
    ops = [kc.tracker(l)["total"] for l in range(a.levels)]
    warm = kc.solve(sl.take(b, 2), tol=a.tol)["seconds"]
    prec = kc.time_precond(1, 3)
    kc.free()
    cnt = qmg.comm_counters()
    qmg.comm_finalize()

    err = latutil.rel_l2(x_loc, sl.take(x_one, 2))
    ok = abs(info["iter"] - info_one["iter"]) <= 1 and err < 1e-8 and info["success"]
    ok = ok and all(abs(p - q) <= max(2, 0.05 * q) for p, q in zip(ops, ops_one))
    print("[rank %d] one GPU: iter %d ops %s relres %.2e | %d slabs: iter %d ops %s relres %.2e | slab error %.2e | halo exchanges %d nccl all-reduces %d p2p %s | warm solve %.4f s, K-cycle apply %.5f s (one GPU %.4f s, %.5f s)"
          % (rank, info_one["iter"], ops_one, info_one["check_relres"], world, info["iter"], ops, info["check_relres"], err,
             cnt["halo_exchanges"], cnt["allreduces"], cnt["p2p"], warm, prec, warm_one, prec_one), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARD-OK" if int(flag.item()) == 1 else "SHARD-MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
