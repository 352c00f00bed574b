#!/usr/bin/env python
"""bench.py -- the hot path of quantum-mg on B200: Wilson stencil apply (Stencil2D::apply_M through
apply_stencil_2D_M, /root/reference/stencil/stencil_2d.h:912,2571) on a synthetic U(1) lattice.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--L 8192] [--impl reference]

A "step" is one out-of-place apply lhs = M rhs over the whole (per-rank) lattice.  `value` is the
whole-job algorithmic GB/s (384 B/site for the stored-block Wilson operator) with all
operands resident in HBM; `e2e` is the same apply through the C ABI with HOST source/result vectors
(pinned host -> device copy of rhs and device -> host copy of lhs inside the timed region).
N > 1: y-slab weak scaling, every rank owns an L x L slab of an L x (N L) lattice; the library
(qmg_comm_init) exchanges one boundary row with the two ring neighbours per apply, overlapped with
the interior rows.  The second half of the metric, the 3-level Wilson K-cycle solve, runs on the
same slabs (`kcycle` in the JSON line).  N > 1 adds two records the driver's SCALE file then carries: `shard_parity`
(a 512 x 512 K-cycle on N slabs against the same solve on one GPU: iteration counts, per-level operator counts, slab error)
and `kcycle_strong` (the SAME 8192 x 8192 lattice cut into N slabs: strong scaling of the solve).

--impl reference: the reference's own CPU implementation (oracle/_ref: the unmodified reference
headers, single thread -- the reference has no threading) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "quantum-mg_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BYTES_PER_SITE = 384.0       # 16 * (nc^2 * 5 + 2 nc), nc = 2
METRIC = "wilson_stencil_GBps"


def ncu_traffic(X, Y):
    """DRAM bytes per launch of the stencil kernel from the committed ncu --set full capture of this workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get("stencil_kernel<2>@%dx%d" % (X, Y))
        return None if e is None else {"bytes": e["traffic_GB"] * 1e9, "algorithmic_bytes": e["algorithmic_GB"] * 1e9, "source": e["source"]}
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem,power.draw,power.limit"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = max([float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()] or [0.0])
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(self.rows)}

        def median(col):
            v = []
            for r in self.rows:
                try:
                    v.append(float(r[col]))
                except (IndexError, ValueError):
                    pass
            v.sort()
            return v[len(v) // 2] if v else None
        # context for sw_power_cap: memory clock and board power against its limit while the kernel runs
        out["mem_mhz"], out["power_w"], out["power_limit_w"] = median(6), median(7), median(8)
        return out


def cpu_reference_apply(L, reps, warm=1):
    """Time the reference's CPU apply (oracle/_ref) on an L x L Wilson lattice; returns (GB/s, seconds per apply)."""
    import capi          # tests/capi.py: the oracle binding (CPU baseline legs only)
    import latutil
    if not capi.have_ref():
        raise RuntimeError("oracle/_ref/libqmg_ref.so missing: run __graft_entry__.build() in the build container")
    be = capi.Backend("ref")
    lat = be.lattice(L, L, 2)
    op = lat.wilson(-0.075, latutil.synthetic_gauge(L, L))
    rhs = latutil.gaussian_cv(lat.size_cv, 1)
    sec = op.time_apply(rhs, 0, warm, reps) / reps
    op.free()
    return BYTES_PER_SITE * L * L / sec / 1e9, sec


def kcycle_run(backend, L, seed=1337, n_refine=2, tol=1e-10, restart=32, gauge=None, world=1, rank=0, Yl=None, link_compressed=False):
    """3-level (n_refine=2) Wilson K-cycle as tests/n13_wilson_kcycle sets it up (4x4 blocks, 8 coarse dof, BiCGstab-6 null
    vectors, MR(2,2) smoothing, inner tol 0.2), gaussian right-hand side; returns the solve record.
    world > 1 (after qmg.comm_init): this rank owns an L x L y-slab of the L x (world L) lattice (weak scaling); the
    hierarchy, the solvers and the K-cycle are the same host code, the library exchanges halo rows and all-reduces dots."""
    import latutil
    if backend == "gpu":
        import driver                     # quantum-mg_b200/driver.py: the product's own driver API (libqmg_host.so)
        be = driver.Backend("gpu")
        KC = driver.KCycle
    else:
        import capi                       # tests/capi.py: the oracle binding (CPU baseline legs only)
        be = capi.Backend(backend)
        KC = capi.KCycle
    # Mass -0.05 everywhere: n13's usage string suggests -0.075, but that is beyond the critical mass of some of the shipped
    # configs (l128t128b60 and l256t256b60 stall there on the CPU reference too: "eigenvalues go negative around -0.075",
    # tests/n13_wilson_kcycle/wilson_kcycle.cpp:81; the oracle's own run is committed as
    # profiles/r03_oracle_mass_m0075_nonconvergence.log), and the chance of an exceptional mode grows with the volume.
    # Iteration caps (inner 100, coarsest 400, outer 100; n13 uses 1000) only bind if a solve stalls.
    mass = -0.05
    if gauge is None:
        try:
            if world > 1:
                raise ValueError
            gauge = latutil.load_gauge(L)          # the reference's own thermalised config where one exists (32, 64, 128, 256)
            cfg = "tests/common_cfgs_u1 l%dt%db60" % (L, L)
        except Exception:
            # The global field is a stack of (L x L/8) stackable slabs, slab u drawn with seed + u; this rank's rows are units
            # [u0, u0 + n_u).  One rank: units 0..7.  Strong scaling (Yl = L / world): the SAME eight units cut N ways -- the N = 1
            # leg of this bench is its reference point.  Weak scaling: eight more units per rank, rank 0's slab = the N = 1 field.
            Ys = Yl or L
            unit = L // 8 if (L % 8 == 0 and (L // 8) % 32 == 0 and Ys % (L // 8) == 0) else Ys
            u0 = rank * (Ys // unit)
            gauge = latutil.synthetic_gauge_units(L, Ys, unit, u0, beta=6.0, seed=seed)
            cfg = ("synthetic non-compact U(1), beta 6.0: stack of stackable (%d x %d) slabs, slab u drawn with seed %d + u; this rank: units %d..%d "
                   "(quantum-mg_b200/latutil.py synthetic_gauge_units)" % (L, unit, seed, u0, u0 + Ys // unit - 1))
    else:
        cfg = "caller-supplied"
    if world > 1:
        # the ranks finish drawing their slabs at different times: meet on the host before the first device-side collective
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    if backend == "gpu" and link_compressed:
        # B200 extension: every coarse operator switches to its link-compressed apply (shared-memory tile kernel) the moment it
        # is built, so the BiCGstab-L null-vector solves of the level below already use it
        be.fn("kcycle_setup_link_compressed")(1)
        # ... and the Wilson fine operator applies matrix-free (gauge links instead of stored blocks: 96 instead of 384 bytes per
        # site, the same bits; Wilson2D::enable_matrix_free_apply checks the stored blocks first) from its first null-vector solve on
        be.fn("kcycle_setup_matrix_free")(1)
    kc = KC(be, L, mass, gauge, n_refine=n_refine, seed=seed, inner_iters=100, coarsest_iters=400, Y=Yl)
    del gauge
    mf_active = 0
    if backend == "gpu" and link_compressed:
        be.fn("kcycle_setup_link_compressed")(0)
        be.fn("kcycle_setup_matrix_free")(0)
        kc.gamma5_hermitian(False)          # the first two solves are the stored-block reference point
        mf_active = kc.matrix_free(True)
        kc.matrix_free(False)
    out = kc.solve(tol=tol, restart=restart, max_iter=100)
    if backend == "gpu":
        # warm-up rule: the first solve also pays the cudaMalloc of every work vector (the block cache is empty); the
        # timed solve is the second one, a fresh gaussian right-hand side on the warm allocator
        first = out
        out = kc.solve(tol=tol, restart=restart, max_iter=100)
        out["first_solve_seconds"], out["first_solve_iter"] = first["seconds"], first["iter"]
        if link_compressed:
            # B200 extension: every level whose stored blocks pass the gamma5-hermiticity check applies its
            # operator from clover / +x / +y blocks only (coarse levels: shared-memory tile kernel).  Same iteration counts.
            stored = out
            n_sw = kc.gamma5_hermitian(True, tile_levels_only=True)      # the nc = 8 levels; the fine level keeps its stored blocks ...
            if mf_active:
                kc.matrix_free(True)                                      # ... or reads its gauge links
            out = kc.solve(tol=tol, restart=restart, max_iter=100)
            if os.environ.get("QMG_BENCH_PROFILE") == "1":
                # one more solve with the per-entry-point profile on (every call bracketed by stream synchronisations: the
                # shares are what to read, the total is slower than the timed solve above); rank 0 prints the table to stderr
                import ctypes as C
                import qmg
                lib = qmg.lib()
                lib.qmg_profile_reset(); lib.qmg_profile_enable(1)
                prof = kc.solve(tol=tol, restart=restart, max_iter=100)
                lib.qmg_profile_enable(0)
                if rank == 0:
                    sys.stdout.flush()
                    lib.qmg_profile_report.restype = C.c_double
                    lib.qmg_profile_report()
                    sys.stderr.write("[bench] profiled solve: %.3f s, %d iterations (timed solve %.3f s)\n" % (prof["seconds"], prof["iter"], out["seconds"]))
            out["first_solve_seconds"], out["first_solve_iter"] = first["seconds"], first["iter"]
            out["link_compressed_levels"] = n_sw
            out["matrix_free_fine_level"] = bool(mf_active)
            out["seconds_stored_blocks"], out["iter_stored_blocks"] = stored["seconds"], stored["iter"]
    out["mass"] = mass
    out["levels"] = n_refine + 1
    out["L"] = L
    out["lattice"] = [L, (Yl or L) * world]
    out["outer_restart"] = restart
    if backend == "gpu":
        import torch
        free, total = torch.cuda.mem_get_info()
        out["hbm_in_use_gb"] = (total - free) / 1e9
    out["config"] = cfg
    out["per_level_ops"] = [kc.tracker(l)["total"] for l in range(n_refine + 1)]
    if backend == "gpu":
        # the trackers count what the reference counts; the fused K-cycle launches fewer (no A.0, no unread true residuals)
        out["per_level_ops_executed"] = [kc.executed(l) for l in range(n_refine + 1)]
    out["per_level_iters"] = [kc.tracker(l)["iters"] for l in range(n_refine + 1)]
    out["precond_apply_s"] = kc.time_precond(1, 2)
    out["wall_s_incl_setup"] = time.perf_counter() - t0
    kc.free()
    return out


def shard_parity_one_gpu(L, levels=3, mass=-0.03, tol=1e-10):
    """Phase A of the multi-rank equivalence check (tests/shard_worker.py in bench form), called BEFORE qmg.comm_init: this
    rank solves the whole L x L lattice alone on its GPU.  Returns what phase B compares against."""
    import driver
    import latutil
    be = driver.Backend("gpu")
    g = latutil.synthetic_gauge(L, L, 6.0, 11)
    b = latutil.gaussian_cv(L * L * 2, 21)
    kc = driver.KCycle(be, L, mass, g, n_refine=levels - 1, block=4, coarse_dof=8, seed=5)
    x, info = kc.solve(b, tol=tol, want_x=True)
    ops = [kc.tracker(l)["total"] for l in range(levels)]
    kc.free()
    return dict(L=L, levels=levels, mass=mass, tol=tol, g=g, b=b, x=x, info=info, ops=ops)


def shard_parity_slabs(one, world, rank):
    """Phase B, after qmg.comm_init: the same gauge field and right-hand side cut into `world` y-slabs, same host code.
    Gates (tests/shard_worker.py): outer iterations +-1, per-level operator counts within 5 %, this rank's slab of the
    one-GPU solution to 1e-8.  Returns the record for the JSON line (max / min over ranks)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import driver
    import latutil
    import shard
    L, levels = one["L"], one["levels"]
    be = driver.Backend("gpu")
    sl = shard.Slab(L, L, world, rank)
    V = L * L
    g_loc = np.concatenate([sl.take(one["g"][:V], 1), sl.take(one["g"][V:], 1)])
    kc = driver.KCycle(be, L, one["mass"], g_loc, Y=sl.Yl, n_refine=levels - 1, block=4, coarse_dof=8, seed=5)
    x_loc, info = kc.solve(sl.take(one["b"], 2), tol=one["tol"], want_x=True)
    ops = [kc.tracker(l)["total"] for l in range(levels)]
    kc.free()
    err = float(latutil.rel_l2(x_loc, sl.take(one["x"], 2)))
    ok = bool(abs(info["iter"] - one["info"]["iter"]) <= 1 and err < 1e-8 and info["success"]
              and all(abs(p - q) <= max(2, 0.05 * q) for p, q in zip(ops, one["ops"])))
    t = torch.tensor([err, 0.0 if ok else 1.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"lattice": [L, L], "slabs": world, "levels": levels, "mass": one["mass"],
            "one_gpu": {"iter": one["info"]["iter"], "per_level_ops": one["ops"], "check_relres": one["info"]["check_relres"]},
            "slabs_result": {"iter": info["iter"], "per_level_ops": ops, "check_relres": info["check_relres"]},
            "max_slab_rel_error": float(t[0].item()), "ok_on_every_rank": bool(t[1].item() == 0.0),
            "gates": "outer iterations +-1, per-level operator counts within 5 %, slab of the one-GPU solution to 1e-8 (tests/shard_worker.py)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = args.cpu_L
    times = []
    for i in range(args.warmup + args.steps):
        gbs, sec = cpu_reference_apply(L, args.cpu_reps, warm=1 if i == 0 else 0)
        if i >= args.warmup:
            times.append(sec)
    sec = sum(times) / len(times)
    val = BYTES_PER_SITE * L * L / sec / 1e9
    sample = "Wilson apply on %dx%d (of the %dx%d workload), %d applies per step, g++ -O2 single thread" % (L, L, args.L, args.L, args.cpu_reps)
    kc = None
    if args.cpu_kcycle_L > 0:
        kc = kcycle_run("ref", args.cpu_kcycle_L)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3 * args.cpu_reps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "wilson_stencil_apply_%dx%d_u1" % (args.L, args.L), "sample_L": L},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kcycle": kc,
    }))


def run_gpu(args):
    import ctypes as C
    import torch
    import qmg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        os.environ["QMG_DEVICE_RNG"] = "1"      # gaussian fills keyed by the GLOBAL element index: N slabs draw what one GPU draws
    qmg.init(local)
    shard_parity = None
    if world > 1:
        one = shard_parity_one_gpu(args.parity_L) if args.parity_L > 0 else None
        qmg.comm_init()          # from here on every lattice handed to the library is this rank's y-slab
        if one is not None:
            shard_parity = shard_parity_slabs(one, world, rank)
            del one
    lib = qmg.lib()
    L = args.L
    strong = args.scaling == "strong"
    if strong and (L % (32 * world) != 0):
        raise SystemExit("--scaling strong needs L divisible by 32 x the number of GPUs (4x4 blocks twice, even slabs)")
    X, Y = L, (L // world if strong else L)
    V = X * Y
    n = 2 * V
    beta = 6.0
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1337 + rank)
    phases = torch.randn(2 * V, generator=gen, device="cuda", dtype=torch.float64) / beta ** 0.5
    gauge = torch.polar(torch.ones_like(phases), phases)
    del phases
    clover, hopping = qmg.fill_wilson(X, Y, gauge)
    rhs = qmg.cvec(n)
    qmg.check(lib.qmg_gaussian(qmg.ptr(rhs), C.c_long(n), C.c_uint64(7), C.c_uint64(rank), C.c_double(1.0)))
    lhs = qmg.cvec(n)
    desc = qmg.stencil_desc(X, Y, 2, clover, hopping, shift=-0.075)

    def step():
        # N > 1: the library sends the two boundary rows of rhs round the ring on its exchange stream while the interior
        # rows are computed, then finishes rows 0 and Y-1 (qmg_stencil_apply, csrc/qmg_stencil.cu apply_sharded)
        qmg.stencil_apply(desc, lhs, rhs)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = qmg.kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = qmg.kernel_launches() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    # the stencil kernel alone (events bracket exchange + kernel when N > 1; at N = 1 a step IS the kernel)
    kern_ms = total_ms / args.steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = BYTES_PER_SITE * V * world / (ms_per_step * 1e-3) / 1e9

    # beside the contract kernel: the opt-in matrix-free apply of the same operator (gauge links instead of stored blocks, 96 B per
    # site, the same bits -- DESIGN.md K1); one GPU only (a slab needs row -1 of U_y from its neighbour, which the host classes fetch)
    matrix_free = None
    if world == 1:
        try:
            dmf = qmg.stencil_desc(X, Y, 2, clover, hopping, shift=-0.075, wilson_gauge=gauge, wilson_w=1.0)
            if qmg.wilson_mf_deviation(dmf) == 0.0:
                chk = qmg.cvec(n)
                qmg.stencil_apply(dmf, chk, rhs)
                same = bool(torch.equal(chk, lhs))
                del chk
                for _ in range(3):
                    qmg.stencil_apply(dmf, lhs, rhs)
                torch.cuda.synchronize()
                m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                m0.record()
                for _ in range(args.steps):
                    qmg.stencil_apply(dmf, lhs, rhs)
                m1.record()
                torch.cuda.synchronize()
                mf_ms = m0.elapsed_time(m1) / args.steps
                matrix_free = {"ms_per_apply": mf_ms, "bytes_per_site": 96.0, "GBps_on_its_bytes": 96.0 * V / (mf_ms * 1e-3) / 1e9,
                               "speedup_vs_stored_blocks": ms_per_step / mf_ms, "output_bit_identical_to_stored_blocks": same,
                               "kernel": "qmg::wilson_mf_tile_kernel<16, 8>", "opt_in": "Wilson2D::enable_matrix_free_apply / qmg_stencil_desc.wilson_gauge"}
            del dmf
        except Exception as exc:      # context only
            matrix_free = {"error": repr(exc)[:200]}
    del gauge

    # end to end: host rhs -> device, apply, device lhs -> host, through the C ABI
    e2e_steps = max(1, min(args.steps, 3))
    hin, hout = C.c_void_p(), C.c_void_p()
    qmg.check(lib.qmg_malloc_host(C.byref(hin), C.c_size_t(16 * n)))
    qmg.check(lib.qmg_malloc_host(C.byref(hout), C.c_size_t(16 * n)))
    qmg.check(lib.qmg_memcpy_d2h(hin, qmg.ptr(rhs), C.c_size_t(16 * n)))
    qmg.stencil_apply_host(desc, hout, hin, dev_lhs=lhs, dev_rhs=rhs)      # warm-up (streams, events)
    # what this box's PCIe gives for the same buffers, one direction at a time (context for the e2e number, not part of it)
    barrier()
    t0 = time.perf_counter()
    qmg.check(lib.qmg_memcpy_h2d(qmg.ptr(rhs), hin, C.c_size_t(16 * n)))
    t1 = time.perf_counter()
    qmg.check(lib.qmg_memcpy_d2h(hout, qmg.ptr(lhs), C.c_size_t(16 * n)))
    t2 = time.perf_counter()
    pcie = {"h2d_GBps": 16 * n / (t1 - t0) / 1e9, "d2h_GBps": 16 * n / (t2 - t1) / 1e9}
    lib.qmg_device_numa_node.restype = C.c_int
    numa = int(lib.qmg_device_numa_node())
    if world > 1:
        # every rank copies at the same time: report the slowest GPU's rates and where each rank's staging memory lives
        import torch.distributed as dist
        t = torch.tensor([pcie["h2d_GBps"], pcie["d2h_GBps"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        pcie = {"h2d_GBps_slowest_gpu": float(t[0].item()), "d2h_GBps_slowest_gpu": float(t[1].item()), "all_gpus_copying_at_once": True}
        nodes = [None] * world
        dist.all_gather_object(nodes, numa)
        pcie["numa_node_of_each_gpu"] = nodes
    else:
        pcie["numa_node_of_gpu"] = numa
    pcie["pinned_staging"] = "allocated and first touched on the GPU's NUMA node (qmg_malloc_host)"
    # ... and with both directions busy at once, which is what the pipelined host-vector apply asks of the link: the time of
    # one upload and one download of the same buffers issued together bounds an e2e step from below (context, not part of e2e)
    try:
        import numpy as np
        h_in = torch.from_numpy(np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_double)), shape=(2 * n,)))
        h_out = torch.from_numpy(np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_double)), shape=(2 * n,)))
        d_in, d_out = torch.view_as_real(rhs).view(-1), torch.view_as_real(lhs).view(-1)
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        both = []
        for _ in range(3):
            barrier()
            tb0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_down):
                h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            both.append(time.perf_counter() - tb0)
        tb = min(both[1:])
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([tb], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tb = float(t.item())
        pcie["both_directions_at_once"] = {"seconds": tb, "GBps_each_way": 16 * n / tb / 1e9, "pinned": bool(h_in.is_pinned()),
                                           "e2e_bound_GBps": BYTES_PER_SITE * V * world / tb / 1e9,
                                           "note": "an e2e step cannot be shorter than this; e2e.value / e2e_bound_GBps is the pipeline's efficiency"}
        del h_in, h_out, d_in, d_out
    except Exception as exc:      # context only: never fail the bench over it
        pcie["both_directions_at_once"] = {"error": repr(exc)[:200]}
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # the host-vector entry of the C ABI: upload, apply and download pipelined over row chunks (sharded: plain route)
        qmg.stencil_apply_host(desc, hout, hin, dev_lhs=lhs, dev_rhs=rhs)
    barrier()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_val = BYTES_PER_SITE * V * world / e2e_sec / 1e9
    lib.qmg_free_host(hin)
    lib.qmg_free_host(hout)
    clocks = sampler.summary() if rank == 0 else None

    # second half of the metric: 3-level Wilson K-cycle solve time (single GPU leg; the sharded solve is reported by --kcycle-sharded)
    kcycle = kcycle_same = None
    if args.kcycle_L > 0:
        del clover, hopping, rhs, lhs, desc
        torch.cuda.empty_cache()
        kcycle = kcycle_run("gpu", args.kcycle_L, restart=args.kcycle_restart, world=world, rank=rank,
                            Yl=(args.kcycle_L // world if strong else None), link_compressed=not args.kcycle_stored_blocks)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([kcycle["seconds"], kcycle["setup_seconds"], kcycle["precond_apply_s"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)     # wall time between device synchronisations, max over ranks
            kcycle["seconds"], kcycle["setup_seconds"], kcycle["precond_apply_s"] = (float(v) for v in t.tolist())
            kcycle["comm"] = qmg.comm_counters()
        if world == 1 and args.cpu_kcycle_L > 0 and not args.no_cpu:
            # hand the big leg's 150 GB of parked blocks back first: the small leg's set-up would otherwise time the driver's frees
            lib.qmg_trim()
            torch.cuda.empty_cache()
            torch.cuda.synchronize()
            kcycle_same = kcycle_run("gpu", args.cpu_kcycle_L)
    # strong scaling of the solve: the SAME kcycle_L x kcycle_L lattice (a stack of N independently drawn slabs) on N GPUs;
    # its N = 1 point is the `kcycle` leg above
    kcycle_strong = None
    if world > 1 and not strong and args.kcycle_L > 0 and not args.no_strong and args.kcycle_L % (32 * world) == 0:
        import torch.distributed as dist
        torch.cuda.empty_cache()
        lib.qmg_trim()
        kcycle_strong = kcycle_run("gpu", args.kcycle_L, restart=args.kcycle_restart, world=world, rank=rank, Yl=args.kcycle_L // world,
                                   link_compressed=not args.kcycle_stored_blocks)
        t = torch.tensor([kcycle_strong["seconds"], kcycle_strong["setup_seconds"], kcycle_strong["precond_apply_s"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kcycle_strong["seconds"], kcycle_strong["setup_seconds"], kcycle_strong["precond_apply_s"] = (float(v) for v in t.tolist())
        kcycle_strong["lattice"] = [args.kcycle_L, args.kcycle_L]
        kcycle_strong["per_gpu_lattice"] = [args.kcycle_L, args.kcycle_L // world]
        kcycle_strong["note"] = "same global lattice size as the N = 1 `kcycle` leg, cut into %d y-slabs; time is the max over ranks" % world

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic = ncu_traffic(X, Y)
        achieved = BYTES_PER_SITE * V / (kern_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                gbs, sec = cpu_reference_apply(args.cpu_L, args.cpu_reps)
                cpu = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "reference",
                       "sample": "Wilson apply on %dx%d, %d applies, oracle/_ref (unmodified reference headers, g++ -O2, 1 thread; the reference is single-threaded)" % (args.cpu_L, args.cpu_L, args.cpu_reps),
                       "ms_per_apply": sec * 1e3, "host_cores_available": os.cpu_count()}
                if args.cpu_kcycle_L > 0:
                    cpu["kcycle"] = kcycle_run("ref", args.cpu_kcycle_L)
                    cpu["kcycle_gpu_same_config"] = kcycle_same
                    if kcycle is not None and cpu["kcycle"].get("iter"):
                        # explicit extrapolation of the CPU reference to the GPU leg's lattice: cost per site per outer iteration is
                        # size-independent on paper (it only grows on a CPU once the working set leaves the caches), so
                        #   t_cpu(L) ~ t_cpu(L0) x (L / L0)^2 x iterations(L) / iterations(L0)      -- a LOWER bound for the CPU
                        kc_cpu = cpu["kcycle"]
                        vol = (args.kcycle_L / float(args.cpu_kcycle_L)) ** 2
                        itr = kcycle["iter"] / float(kc_cpu["iter"])
                        ext = {"to_lattice": [args.kcycle_L, args.kcycle_L], "from_lattice": [args.cpu_kcycle_L, args.cpu_kcycle_L],
                               "solve_seconds": kc_cpu["seconds"] * vol * itr, "setup_seconds": kc_cpu["setup_seconds"] * vol,
                               "formula": "t_cpu(%d^2) x (%d/%d)^2 x iter_gpu(%d^2)/iter_cpu(%d^2) = %.2f s x %.0f x %d/%d; set-up: %.2f s x %.0f"
                                          % (args.cpu_kcycle_L, args.kcycle_L, args.cpu_kcycle_L, args.kcycle_L, args.cpu_kcycle_L,
                                             kc_cpu["seconds"], vol, kcycle["iter"], kc_cpu["iter"], kc_cpu["setup_seconds"], vol),
                               "gpu_solve_seconds_measured": kcycle["seconds"], "gpu_setup_seconds_measured": kcycle["setup_seconds"]}
                        ext["solve_speedup_vs_extrapolated_cpu"] = ext["solve_seconds"] / kcycle["seconds"]
                        cpu["kcycle_extrapolated"] = ext
                        sys.stderr.write("[bench] CPU reference K-cycle extrapolated to %dx%d: solve %.0f s (%.1f h), set-up %.0f s; GPU measured: solve %.2f s, set-up %.2f s\n"
                                         % (args.kcycle_L, args.kcycle_L, ext["solve_seconds"], ext["solve_seconds"] / 3600.0, ext["setup_seconds"],
                                            kcycle["seconds"], kcycle["setup_seconds"]))
            except Exception as e:  # the baseline is a report, never the product path
                cpu = {"value": None, "unit": "GB/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % e}
        emit(json.dumps({
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "wilson_stencil_apply_%dx%d_u1" % (X, Y * world), "per_gpu_lattice": [X, Y], "beta": beta,
                       "mass": {"stencil_apply": -0.075, "kcycle": -0.05,
                                "why": "the apply only adds the mass to the diagonal; the K-cycle legs use -0.05 because n13's -0.075 is beyond critical on the reference's own 128^2 / 256^2 configs (profiles/r03_oracle_mass_m0075_nonconvergence.log)"},
                       "bytes_per_site": BYTES_PER_SITE, "l2": "operands (%.1f GB per apply) far larger than the 126 MB L2; no flush needed" % (BYTES_PER_SITE * V / 1e9),
                       "parallelism": "y-slabs x%d, 1-row halo ring" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": (traffic or {}).get("bytes"), "traffic_unit": "DRAM bytes per launch (ncu --set full)",
                         "algorithmic_bytes_per_launch": BYTES_PER_SITE * V, "traffic_source": (traffic or {}).get("source"),
                         "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / 8000.0, "kernel": "qmg::stencil_kernel<2, 0, 0, 0>"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n, "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps,
                    "pcie_one_direction_at_a_time": pcie},
            "gpu_launches": launches,
            "clocks": clocks,
            "wilson_matrix_free": matrix_free,
            "kcycle": kcycle,
            "kcycle_strong": kcycle_strong,
            "shard_parity": shard_parity,
        }))
    if world > 1:
        import torch.distributed as dist
        qmg.comm_finalize()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    # Libraries print banners on stdout (NCCL: "NCCL version ..."); stdout must carry exactly one JSON line, so everything
    # else is sent to stderr and the line goes to a saved copy of the original descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--L", type=int, default=8192, help="per-GPU lattice is L x L")
    ap.add_argument("--kcycle-stored-blocks", action="store_true", dest="kcycle_stored_blocks",
                    help="time the K-cycle with the stored-block applies only (default: also with the link-compressed applies, which is the reported time)")
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak",
                    help="weak: L x L per GPU (an L x N L lattice); strong: the L x L lattice cut into N slabs of L / N rows")
    ap.add_argument("--cpu-L", type=int, default=2048, dest="cpu_L")
    ap.add_argument("--cpu-reps", type=int, default=5, dest="cpu_reps")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--kcycle-L", type=int, default=8192, dest="kcycle_L", help="3-level K-cycle solve on L x L per GPU after the stencil run (0 = skip)")
    ap.add_argument("--kcycle-restart", type=int, default=8, dest="kcycle_restart",
                    help="restart length of the outer flexible GCR (n13 uses 32; at 8192^2 per GPU 2 x 32 stored 2.1 GB vectors do not fit beside the hierarchy; 8, 16 and 32 need the same 22 iterations at 4096^2)")
    ap.add_argument("--cpu-kcycle-L", type=int, default=256, dest="cpu_kcycle_L",
                    help="K-cycle size for the CPU reference leg: BASELINE config 2's 256 x 256 on the reference's own l256t256b60 config (~35 s of CPU: set-up + solve; 0 = skip)")
    ap.add_argument("--parity-L", type=int, default=512, dest="parity_L", help="N > 1: lattice of the N-slabs-against-one-GPU K-cycle equivalence check (0 = skip)")
    ap.add_argument("--no-strong", action="store_true", dest="no_strong", help="N > 1: skip the strong-scaling K-cycle leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
