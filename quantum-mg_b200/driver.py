"""ctypes view of the flat driver API (host/qmg_capi_body.h) exported by libqmg_host.so: the reference's class API
-- Lattice2D, the operators, TransferMG, CoarseOperator2D, StatefulMultigridMG, the solvers and the whole n13 / n22
set-up + K-cycle solve -- driven from Python with numpy complex128 arrays in and out.  Everything behind it runs on the
GPU (include/qmg host classes over libqmg_b200.so); there is no CPU path here.

`Backend()` binds the product library.  The same driver text compiled against the unmodified reference headers is the
oracle; the tests bind that one through tests/capi.py (`Backend("ref")`), which passes its own library path and symbol
prefix to the generic constructor below -- this module never names or loads anything under oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.path.join(_HERE, "libqmg_host.so")

CD = np.complex128


def _c(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def carr(a):
    return np.ascontiguousarray(a, dtype=CD)


class Backend:
    def __init__(self, kind="gpu", lib_path=None, prefix=None):
        self.kind = kind
        if lib_path is None:
            if kind != "gpu":
                raise ValueError("quantum-mg_b200.driver only knows the 'gpu' backend (got %r)" % (kind,))
            lib_path, prefix = GPU_LIB, "qmgh_"
            if not os.path.exists(lib_path):
                raise RuntimeError("libqmg_host.so not built (run __graft_entry__.build()); there is no CPU fallback")
        self.lib, self.prefix = C.CDLL(lib_path), prefix
        for name in ("lattice_new", "wilson_new", "staggered_new", "laplace_new", "dwf_new", "generic_new",
                     "transfer_new", "transfer_new_asym", "coarse_new", "mg_new", "mg_get_stencil"):
            self.fn(name).restype = C.c_void_p
        self.fn("stencil_get").restype = C.c_long
        self.fn("stencil_time_apply").restype = C.c_double

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- lattice
    def lattice(self, X, Y, nc):
        return Lattice(self, X, Y, nc)


class Lattice:
    def __init__(self, be, X, Y, nc):
        self.be, self.X, self.Y, self.nc = be, X, Y, nc
        self.h = C.c_void_p(be.fn("lattice_new")(X, Y, nc))
        out = (C.c_int * 6)()
        be.fn("lattice_sizes")(self.h, out)
        self.volume, self.size_cv, self.size_cm, self.size_gauge, self.size_hopping, self.size_corner = list(out)

    def coord_to_index(self, x, y):
        return self.be.fn("lattice_coord_to_index")(self.h, x, y)

    def index_to_coord(self, i):
        xy = (C.c_int * 2)()
        self.be.fn("lattice_index_to_coord")(self.h, i, xy)
        return xy[0], xy[1]

    def cshift(self, rhs, cdir, eo, dof, lhs=None):
        rhs = carr(rhs)
        lhs = np.zeros(self.volume * dof, CD) if lhs is None else carr(lhs).copy()
        self.be.fn("cshift")(_c(lhs), _c(rhs), cdir, eo, dof, self.h)
        return lhs

    # ---- operators
    # ---- time-slice reductions / wall sources (reductions/reductions.h)
    def timeslice(self, op, a, b=None):
        aa = carr(a)
        bb = None if b is None else carr(b)
        out = np.zeros(self.Y * (2 if op == 2 else 1), np.float64)
        self.be.fn("timeslice")(self.h, op, _c(aa), _c(bb), _c(out))
        return out.view(CD) if op == 2 else out

    def wall_source(self, timeslice, color, seed, deviation=1.0, mean=0.0):
        out = np.full(self.size_cv, 7.0 + 0j, CD)
        self.be.fn("wall_source")(self.h, timeslice, color, C.c_uint(seed), C.c_double(deviation), C.c_double(mean), _c(out))
        return out

    # ---- U(1) gauge side (u1/u1_utils.h); nc = 1 lattices
    def u1_observables(self, gauge):
        g = carr(gauge)
        out = (C.c_double * 3)()
        self.be.fn("u1_observables")(self.h, _c(g), out)
        return complex(out[0], out[1]), out[2]

    def u1_action(self, phases, beta):
        p = np.ascontiguousarray(phases, dtype=np.float64)
        f = self.be.fn("u1_action")
        f.restype = C.c_double
        return f(self.h, _c(p), C.c_double(beta))

    def u1_polar(self, phases):
        p = np.ascontiguousarray(phases, dtype=np.float64)
        out = np.zeros(p.size, CD)
        self.be.fn("u1_polar")(self.h, _c(p), _c(out))
        return out

    def u1_gauge_trans(self, gauge, trans):
        g, t = carr(gauge).copy(), carr(trans)
        self.be.fn("u1_gauge_trans")(self.h, _c(g), _c(t))
        return g

    def u1_ape_smear(self, gauge, alpha, n_iter):
        g = carr(gauge)
        out = np.zeros(g.size, CD)
        self.be.fn("u1_ape_smear")(self.h, _c(out), _c(g), C.c_double(alpha), n_iter)
        return out

    def u1_instanton(self, gauge, Q, x0, y0):
        g = carr(gauge).copy()
        self.be.fn("u1_instanton")(self.h, _c(g), C.c_double(Q), x0, y0)
        return g

    def u1_noncompact_instanton(self, phases, Q):
        p = np.ascontiguousarray(phases, dtype=np.float64).copy()
        self.be.fn("u1_noncompact_instanton")(self.h, _c(p), C.c_double(Q))
        return p

    def u1_heatbath(self, phases, beta, n_update, seed):
        p = np.ascontiguousarray(phases, dtype=np.float64).copy()
        self.be.fn("u1_heatbath")(self.h, _c(p), C.c_double(beta), n_update, C.c_uint(seed))
        return p

    def u1_file(self, kind, path, gauge=None, phases=None):
        n = 2 * self.X * self.Y
        if kind == 0:
            gauge = np.zeros(n, CD)
        if kind == 2:
            phases = np.zeros(n, np.float64)
        g = None if gauge is None else carr(gauge)
        p = None if phases is None else np.ascontiguousarray(phases, dtype=np.float64)
        self.be.fn("u1_file")(self.h, kind, path.encode(), _c(g), _c(p))
        return g if kind == 0 else (p if kind == 2 else None)

    def u1_create(self, kind, beta=1.0, seed=1):
        n = self.X * self.Y * (1 if kind == 3 else 2)
        out = np.zeros(n, CD)
        self.be.fn("u1_create")(self.h, kind, C.c_double(beta), C.c_uint(seed), _c(out))
        return out

    def wilson(self, mass, gauge, wilson_coeff=1.0):
        m = complex(mass)
        g = carr(gauge)
        return Stencil(self, C.c_void_p(self.be.fn("wilson_new")(self.h, C.c_double(m.real), C.c_double(m.imag), _c(g), C.c_double(wilson_coeff))))

    def staggered(self, mass, gauge):
        m = complex(mass)
        g = carr(gauge)
        return Stencil(self, C.c_void_p(self.be.fn("staggered_new")(self.h, C.c_double(m.real), C.c_double(m.imag), _c(g))))

    def laplace(self, msq, gauge):
        m = complex(msq)
        g = carr(gauge)
        return Stencil(self, C.c_void_p(self.be.fn("laplace_new")(self.h, C.c_double(m.real), C.c_double(m.imag), _c(g))))

    def dwf(self, mass, gauge, Ls, M5=-1.0):
        m = complex(mass)
        g = carr(gauge)
        return Stencil(self, C.c_void_p(self.be.fn("dwf_new")(self.h, C.c_double(m.real), C.c_double(m.imag), _c(g), Ls, C.c_double(M5))))

    def generic(self, clover, hopping, shift=0.0, eo_shift=0.0, dof_shift=0.0, is_chiral=True, def_chirality=1):
        s = [complex(shift), complex(eo_shift), complex(dof_shift)]
        s6 = (C.c_double * 6)(s[0].real, s[0].imag, s[1].real, s[1].imag, s[2].real, s[2].imag)
        cl = None if clover is None else carr(clover)
        hp = None if hopping is None else carr(hopping)
        return Stencil(self, C.c_void_p(self.be.fn("generic_new")(self.h, int(is_chiral), def_chirality, s6, _c(cl), _c(hp))))


class Stencil:
    NAMES = ["clover", "hopping", "dagger_clover", "dagger_hopping", "rbjacobi_clover", "rbjacobi_hopping",
             "rbjacobi_cinv", "rbj_dagger_clover", "rbj_dagger_hopping", "rbj_dagger_cinv"]

    def __init__(self, lat, h, owner=None):
        self.lat, self.be, self.h, self.owner = lat, lat.be, h, owner

    def get(self, name):
        which = self.NAMES.index(name)
        n = self.be.fn("stencil_get")(self.h, which, None)
        if n == 0:
            return None
        out = np.zeros(n, CD)
        self.be.fn("stencil_get")(self.h, which, _c(out))
        return out

    def add_to(self, name, noise):
        noise = carr(noise)
        self.be.fn("stencil_add_to")(self.h, self.NAMES.index(name), _c(noise))

    def shifts(self):
        out = (C.c_double * 6)()
        self.be.fn("stencil_get_shifts")(self.h, out)
        return complex(out[0], out[1]), complex(out[2], out[3]), complex(out[4], out[5])

    def update_shifts(self, shift, eo_shift, dof_shift):
        s = [complex(shift), complex(eo_shift), complex(dof_shift)]
        self.be.fn("stencil_update_shifts")(self.h, (C.c_double * 6)(s[0].real, s[0].imag, s[1].real, s[1].imag, s[2].real, s[2].imag))

    def build(self, dagger=False, rbjacobi=False, rbj_dagger=False):
        self.be.fn("stencil_build")(self.h, (1 if dagger else 0) | (2 if rbjacobi else 0) | (4 if rbj_dagger else 0))

    def built(self):
        return self.be.fn("stencil_built")(self.h)

    def apply(self, rhs, type=0, lhs=None):
        rhs = carr(rhs)
        lhs = np.zeros(self.lat.size_cv, CD) if lhs is None else carr(lhs).copy()
        self.be.fn("stencil_apply")(self.h, type, _c(lhs), _c(rhs))
        return lhs

    def apply_piece(self, piece, rhs, dir=0, lhs=None):
        rhs = carr(rhs)
        lhs = np.zeros(self.lat.size_cv, CD) if lhs is None else carr(lhs).copy()
        self.be.fn("stencil_apply_piece")(self.h, piece, dir, _c(lhs), _c(rhs))
        return lhs

    def prepare(self, b, type):
        b = carr(b)
        out = np.zeros(self.lat.size_cv, CD)
        self.be.fn("stencil_prepare")(self.h, type, _c(out), _c(b))
        return out

    def reconstruct(self, y, b, type):
        y, b = carr(y), carr(b)
        out = np.zeros(self.lat.size_cv, CD)
        self.be.fn("stencil_reconstruct")(self.h, type, _c(out), _c(y), _c(b))
        return out

    def chiral(self, op, a, b=None):
        a = carr(a).copy()
        bb = None if b is None else carr(b).copy()
        self.be.fn("stencil_chiral")(self.h, op, _c(a), _c(bb))
        return a, bb

    def gamma5_hermitian(self, on=True):
        return int(self.be.fn("stencil_gamma5_hermitian")(self.h, 1 if on else 0))

    def matrix_free(self, gauge=None, on=1):
        """B200 extension (Wilson2D only): apply from the gauge links instead of the stored blocks, same bits.  on = 1 needs the gauge
        array the operator was built from; 0 switches off, 2 / 3 pause / resume.  Returns 1 when active."""
        return int(self.be.fn("stencil_matrix_free")(self.h, _c(carr(gauge)) if gauge is not None else None, int(on)))

    def eigs(self, type, nev, ncv=None, high=False, tol=1e-8, want_vectors=False):
        """B200 build only: nev extreme eigenpairs of the Hermitian operator `type` (arpack_dcn's Lanczos restatement)."""
        ncv = 3 * nev if ncv is None else ncv
        ev = np.zeros(nev, np.float64)
        vec = np.zeros((nev, self.lat.size_cv), CD) if want_vectors else None
        ok = self.be.fn("stencil_eigs")(self.h, type, nev, ncv, 1 if high else 0, C.c_double(tol), _c(ev), _c(vec))
        return (ok == 1), ev, vec

    def coarse_sigma(self, type, v):
        """CoarseOperator2D::apply_sigma(out, v, QMGSigmaTypeCoarse type in 6..9); out starts as zeros."""
        out = np.zeros(self.lat.size_cv, CD)
        vv = carr(v)
        assert self.lat.be.fn("coarse_apply_sigma")(self.h, type, _c(out), _c(vv)) == 1
        return out

    def time_apply(self, rhs, type=0, warm=1, reps=5):
        rhs = carr(rhs)
        return self.be.fn("stencil_time_apply")(self.h, type, warm, reps, _c(rhs))

    def solve(self, solver, b, type=0, x0=None, n=None, max_iter=1000, tol=1e-8, iparam=32, dparam=0.85, verbosity=0):
        b = carr(b)
        x = np.zeros(self.lat.size_cv, CD) if x0 is None else carr(x0).copy()
        n = self.lat.size_cv if n is None else n
        info = (C.c_double * 4)()
        self.be.fn("solve")(self.h, solver, type, _c(x), _c(b), n, max_iter, C.c_double(tol), iparam, C.c_double(dparam), verbosity, info)
        return x, dict(resSq=info[0], iter=int(info[1]), success=bool(info[2]), ops=int(info[3]))

    def free(self):
        if self.h is not None:
            self.be.fn("stencil_free")(self.h)
            self.h = None


class Transfer:
    def __init__(self, fine, coarse, nullvecs, block_ortho=True, save_decomp=False, doubling=0, restrict_vecs=None):
        self.fine, self.coarse, self.be = fine, coarse, fine.be
        nv = carr(nullvecs).reshape(coarse.nc, fine.size_cv)
        if restrict_vecs is None:
            self.h = C.c_void_p(self.be.fn("transfer_new")(fine.h, coarse.h, _c(nv), int(block_ortho), int(save_decomp), doubling))
        else:
            rv = carr(restrict_vecs).reshape(coarse.nc, fine.size_cv)
            self.h = C.c_void_p(self.be.fn("transfer_new_asym")(fine.h, coarse.h, _c(nv), _c(rv), int(block_ortho), int(save_decomp), doubling))

    def nullvecs(self, which=0):
        out = np.zeros((self.coarse.nc, self.fine.size_cv), CD)
        n = self.be.fn("transfer_get_nullvecs")(self.h, which, _c(out))
        return out if n else None

    def prolong(self, coarse_v, fine_v=None):
        cv = carr(coarse_v)
        fv = np.zeros(self.fine.size_cv, CD) if fine_v is None else carr(fine_v).copy()
        self.be.fn("transfer_prolong")(self.h, _c(cv), _c(fv))
        return fv

    def restrict(self, fine_v, coarse_v=None):
        fv = carr(fine_v)
        cv = np.zeros(self.coarse.size_cv, CD) if coarse_v is None else carr(coarse_v).copy()
        self.be.fn("transfer_restrict")(self.h, _c(fv), _c(cv))
        return cv

    def props(self):
        p = self.be.fn("transfer_props")(self.h)
        return dict(symmetric=bool(p & 1), has_decomp=bool(p & 2), init=bool(p & 4), doubling=p >> 4)

    def cholesky(self):
        out = np.zeros(self.coarse.size_cm, CD)
        self.be.fn("transfer_get_cholesky")(self.h, _c(out))
        return out

    def LU(self):
        L, U = np.zeros(self.coarse.size_cm, CD), np.zeros(self.coarse.size_cm, CD)
        self.be.fn("transfer_get_LU")(self.h, _c(L), _c(U))
        return L, U

    def coarse_operator(self, fine_stencil, is_chiral=True, use_rbjacobi=False, build_extra=0):
        h = C.c_void_p(self.be.fn("coarse_new")(self.coarse.h, fine_stencil.h, self.fine.h, self.h, int(is_chiral), int(use_rbjacobi), build_extra))
        return Stencil(self.coarse, h, owner=(self, fine_stencil))

    def free(self):
        if self.h is not None:
            self.be.fn("transfer_free")(self.h)
            self.h = None


class Multigrid:
    """StatefulMultigridMG driven like tests/n13_wilson_kcycle/wilson_kcycle.cpp."""

    def __init__(self, lat0, stencil0, coarsest_type=0, coarsest_tol=0.2, coarsest_iters=1000, coarsest_restart=32):
        self.be, self.lats, self.keep = lat0.be, [lat0], [stencil0]
        self.h = C.c_void_p(self.be.fn("mg_new")(lat0.h, stencil0.h, coarsest_type, C.c_double(coarsest_tol), coarsest_iters, coarsest_restart))

    def push_level(self, new_lat, transfer, fine_stencil_app=0, inner_tol=0.2, inner_iters=1000, inner_restart=32,
                   pre_iters=2, post_iters=2, pre_tol=1e-15, post_tol=1e-15, pre_cgne=False, post_cgne=False,
                   build_stencil=True, is_chiral=True, build_from=0, build_extra=0, nvecs=None):
        ip = (C.c_int * 7)(fine_stencil_app, inner_iters, inner_restart, pre_iters, post_iters, int(pre_cgne), int(post_cgne))
        dp = (C.c_double * 3)(inner_tol, pre_tol, post_tol)
        nv = None if nvecs is None else carr(nvecs)
        self.be.fn("mg_push_level")(self.h, new_lat.h, transfer.h, ip, dp, int(build_stencil), int(is_chiral), build_from, build_extra, _c(nv))
        self.lats.append(new_lat)
        self.keep.append(transfer)

    def num_levels(self):
        return self.be.fn("mg_num_levels")(self.h)

    def stencil(self, level):
        return Stencil(self.lats[level], C.c_void_p(self.be.fn("mg_get_stencil")(self.h, level)), owner=self)

    def precond(self, rhs, verbosity=0):
        rhs = carr(rhs)
        lhs = np.zeros_like(rhs)
        self.be.fn("mg_precond")(self.h, _c(lhs), _c(rhs), verbosity)
        return lhs

    def solve(self, b, outer_type=0, x0=None, max_iter=1000, tol=1e-10, restart=32, verbosity=0):
        b = carr(b)
        x = np.zeros_like(b) if x0 is None else carr(x0).copy()
        info = (C.c_double * 5)()
        self.be.fn("mg_solve")(self.h, outer_type, _c(x), _c(b), max_iter, C.c_double(tol), restart, verbosity, info)
        return x, dict(resSq=info[0], iter=int(info[1]), success=bool(info[2]), ops=int(info[3]), seconds=info[4])

    def tracker(self, level):
        out = (C.c_int * 6)()
        self.be.fn("mg_tracker")(self.h, level, out)
        return dict(nullvec=out[0], krylov=out[1], presmooth=out[2], postsmooth=out[3], total=out[4], iters=out[5])

    def reset_tracker(self):
        self.be.fn("mg_reset_tracker")(self.h)

    def executed(self, level):
        """Operator applications actually launched at `level` (B200: the fused K-cycle skips A.0 and unread residuals)."""
        f = self.be.fn("mg_executed")
        f.restype = C.c_long
        return int(f(self.h, level))

    def set_fused(self, on):
        """B200 extension: 1 / True fused K-cycle with the pre-smoother's residual handed over (default), 2 fused with the explicit
        residual (bit-identical to 0), 0 / False the reference's sweep-for-sweep sequence; returns the old setting."""
        return int(self.be.fn("mg_set_fused")(self.h, int(on)))

    def apply_stencil(self, level, rhs, type=0):
        rhs = carr(rhs)
        lhs = np.zeros_like(rhs)
        self.be.fn("mg_apply_stencil")(self.h, level, type, _c(lhs), _c(rhs))
        return lhs

    def free(self):
        if self.h is not None:
            self.be.fn("mg_free")(self.h)
            self.h = None


class KCycle:
    """The whole n13-style setup + solve as native calls (kcycle_* in qmg_capi_body.h)."""

    def __init__(self, be, L, mass, gauge, n_refine=1, block=4, coarse_dof=8, pre_iters=2, post_iters=2, inner_tol=0.2, inner_iters=1000,
                 inner_restart=32, coarsest_tol=0.2, coarsest_iters=1000, coarsest_restart=32, null_max_iter=500, null_tol=5e-5, null_L=6,
                 level_app=0, coarsest_app=0, pre_tol=1e-15, post_tol=1e-15, seed=1337, verbosity=0, Y=None, staggered=False, adaptive_setups=None):
        self.be, self.X, self.Y = be, L, (L if Y is None else Y)
        self.nc0 = 1 if staggered else 2
        for name in ("kcycle_new", "kcycle_new_adaptive", "kcycle_mg"):
            be.fn(name).restype = C.c_void_p
        be.fn("kcycle_time_precond").restype = C.c_double
        ip = (C.c_int * 15)(n_refine, block, block, coarse_dof, pre_iters, post_iters, inner_iters, inner_restart, coarsest_iters,
                            coarsest_restart, null_max_iter, null_L, level_app, coarsest_app, 1 if staggered else 0)
        dp = (C.c_double * 5)(inner_tol, coarsest_tol, null_tol, pre_tol, post_tol)
        g = carr(gauge)
        self.n_levels = n_refine + 1
        if adaptive_setups is None:
            self.h = C.c_void_p(be.fn("kcycle_new")(self.X, self.Y, C.c_double(mass), _c(g), ip, dp, C.c_uint(seed), verbosity))
        else:       # tests/n22_wilson_kcycle_adaptive: Richardson-relaxed test vectors refined by the K-cycle itself
            self.h = C.c_void_p(be.fn("kcycle_new_adaptive")(self.X, self.Y, C.c_double(mass), _c(g), ip, dp, int(adaptive_setups), C.c_uint(seed), verbosity))
        self._mg = C.c_void_p(be.fn("kcycle_mg")(self.h))

    def solve(self, b=None, outer_type=0, max_iter=1000, tol=1e-10, restart=32, verbosity=0, want_x=False):
        n = self.X * self.Y * self.nc0
        bb = None if b is None else carr(b)
        x = np.zeros(n, CD) if want_x else None
        info = (C.c_double * 8)()
        self.be.fn("kcycle_solve")(self.h, _c(bb), _c(x), outer_type, max_iter, C.c_double(tol), restart, verbosity, info)
        out = dict(resSq=info[0], iter=int(info[1]), success=bool(info[2]), ops=int(info[3]), seconds=info[4], check_relres=info[5],
                   setup_seconds=info[6], null_ops=int(info[7]))
        return (x, out) if want_x else out

    def tracker(self, level):
        out = (C.c_int * 6)()
        self.be.fn("mg_tracker")(self._mg, level, out)
        return dict(nullvec=out[0], krylov=out[1], presmooth=out[2], postsmooth=out[3], total=out[4], iters=out[5])

    def executed(self, level):
        f = self.be.fn("mg_executed")
        f.restype = C.c_long
        return int(f(self._mg, level))

    def set_fused(self, on):
        return int(self.be.fn("mg_set_fused")(self._mg, int(on)))

    def update_links(self, gauge):
        """New gauge links into the same Wilson operator and a fresh hierarchy (the n16 measurement-loop step)."""
        g = carr(gauge)
        self.be.fn("kcycle_update_links")(self.h, _c(g))
        self._mg = C.c_void_p(self.be.fn("kcycle_mg")(self.h))

    def pion(self, x0=0, y0=0, max_iter=1000, tol=1e-10, restart=32, verbosity=0):
        """Folded would-be pion correlator from a point source (n16 :452-506): (Y values, info dict)."""
        out = np.zeros(self.Y, np.float64)
        info = (C.c_double * 3)()
        self.be.fn("kcycle_pion")(self.h, x0, y0, max_iter, C.c_double(tol), restart, verbosity, _c(out), info)
        return out, dict(iters=int(info[0]), success=bool(info[1]), seconds=info[2])

    def gamma5_hermitian(self, on=True, tile_levels_only=False):
        """B200 extension: link-compressed applies on every level that passes the check (tile_levels_only: only on the levels
        with nc >= 4, where the shared-memory patch kernels make it faster); returns how many levels switched."""
        return int(self.be.fn("kcycle_gamma5_hermitian")(self.h, (2 if tile_levels_only else 1) if on else 0))

    def matrix_free(self, on=True):
        """B200 extension: pause / resume the matrix-free apply of the Wilson fine operator (set up at construction when
        kcycle_setup_matrix_free(1) was in force); returns 1 when it is active afterwards."""
        return int(self.be.fn("kcycle_matrix_free")(self.h, 1 if on else 0))

    def deflate_coarsest(self, num_low, num_high=0):
        """B200 build only: eigenpairs of the coarsest normal operator for the deflated coarsest solve; returns their eigenvalues."""
        ev = np.zeros(num_low + num_high, np.float64)
        n = int(self.be.fn("kcycle_deflate_coarsest")(self.h, num_low, num_high, _c(ev)))
        return ev[:n]

    def time_precond(self, warm=1, reps=3):
        return self.be.fn("kcycle_time_precond")(self.h, warm, reps) / reps

    def free(self):
        if self.h is not None:
            self.be.fn("kcycle_free")(self.h)
            self.h = None
