"""ctypes binding of libqmg_b200.so (the C ABI in include/qmg_b200.h).

Host-side plumbing only: device memory is torch tensors (complex128, whose
interleaved (re, im) layout is exactly the reference's complex<double>), the
kernels are the hand-written sm_100a ones in csrc/.  There is no CPU path: if the
shared library is missing, or no CUDA device is visible, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqmg_b200.so")
HOST_LIB_PATH = os.path.join(_HERE, "libqmg_host.so")

# pieces bitmask / flags (include/qmg_b200.h)
APPLY_CLOVER, APPLY_HOP_TO_EVEN, APPLY_HOP_TO_ODD, APPLY_SHIFT, APPLY_ALL = 1, 2, 4, 8, 15
APPLY_IDENTITY_CLOVER, APPLY_ACCUMULATE, APPLY_EVEN_ROWS_ONLY, APPLY_ODD_ROWS_ONLY = 16, 32, 64, 128

# QMGStencilType (/root/reference/stencil/stencil_2d.h:63-74)
MATVEC_ORIGINAL, MATVEC_DAGGER, MATVEC_RIGHT_JACOBI, MATVEC_RIGHT_SCHUR = 0, 1, 2, 3
MATVEC_M_MDAGGER, MATVEC_MDAGGER_M, MATVEC_RBJ_DAGGER, MATVEC_RBJ_M_MDAGGER, MATVEC_RBJ_MDAGGER_M = 4, 5, 6, 7, 8


class StencilDesc(C.Structure):
    _fields_ = [("X", C.c_int), ("Y", C.c_int), ("nc", C.c_int),
                ("clover", C.c_void_p), ("hopping", C.c_void_p),
                ("shift", C.c_double * 2), ("eo_shift", C.c_double * 2), ("dof_shift", C.c_double * 2),
                ("halo_ym", C.c_void_p), ("halo_yp", C.c_void_p),
                ("gamma5_hermitian", C.c_int), ("hop_halo_ym", C.c_void_p),
                ("wilson_gauge", C.c_void_p), ("wilson_w", C.c_double), ("wilson_gauge_halo_ym", C.c_void_p)]


class TransferDesc(C.Structure):
    _fields_ = [("Xf", C.c_int), ("Yf", C.c_int), ("ncf", C.c_int),
                ("Xc", C.c_int), ("Yc", C.c_int), ("ncc", C.c_int)]


class QmgError(RuntimeError):
    pass


_lib = None


def exported_symbols():
    """Every symbol include/qmg_b200.h declares (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "qmg_b200.h")
    text = open(hdr).read()
    return sorted(set(re.findall(r"\b(qmg_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load libqmg_b200.so; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QmgError("libqmg_b200.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.qmg_last_error.restype = C.c_char_p
        _lib.qmg_get_stream.restype = C.c_void_p
        _lib.qmg_kernel_launches.restype = C.c_long
    return _lib


def check(rc):
    if rc != 0:
        raise QmgError(lib().qmg_last_error().decode() or "libqmg_b200 call failed (rc=%d)" % rc)


def init(device=None, use_torch_stream=True):
    """Select the device (default: torch's current one) and adopt torch's current stream."""
    import torch
    if not torch.cuda.is_available():
        raise QmgError("no CUDA device visible: quantum-mg_b200 has no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    torch.cuda.set_device(dev)
    check(lib().qmg_init(dev))
    if use_torch_stream:
        check(lib().qmg_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return dev


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def cvec(n, zero=True):
    import torch
    f = torch.zeros if zero else torch.empty
    return f(int(n), dtype=torch.complex128, device="cuda")


def to_device(a):
    import numpy as np
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)).cuda()


def stencil_desc(X, Y, nc, clover=None, hopping=None, shift=0.0, eo_shift=0.0, dof_shift=0.0, halo_ym=None, halo_yp=None,
                 gamma5_hermitian=False, hop_halo_ym=None, wilson_gauge=None, wilson_w=1.0, wilson_gauge_halo_ym=None):
    d = StencilDesc()
    d.X, d.Y, d.nc = int(X), int(Y), int(nc)
    d.clover = clover.data_ptr() if clover is not None else None
    d.hopping = hopping.data_ptr() if hopping is not None else None
    for name, v in (("shift", shift), ("eo_shift", eo_shift), ("dof_shift", dof_shift)):
        v = complex(v)
        getattr(d, name)[0] = v.real
        getattr(d, name)[1] = v.imag
    d.halo_ym = halo_ym.data_ptr() if halo_ym is not None else None
    d.halo_yp = halo_yp.data_ptr() if halo_yp is not None else None
    d.gamma5_hermitian = 1 if gamma5_hermitian else 0
    d.hop_halo_ym = hop_halo_ym.data_ptr() if hop_halo_ym is not None else None
    d.wilson_gauge = wilson_gauge.data_ptr() if wilson_gauge is not None else None
    d.wilson_w = float(wilson_w)
    d.wilson_gauge_halo_ym = wilson_gauge_halo_ym.data_ptr() if wilson_gauge_halo_ym is not None else None
    d._keep = (clover, hopping, halo_ym, halo_yp, hop_halo_ym, wilson_gauge, wilson_gauge_halo_ym)
    return d


def wilson_mf_deviation(desc):
    """sum |stored - regenerated|^2 of an nc = 2 set against the Wilson blocks of desc.wilson_gauge: 0.0 exactly licenses the matrix-free apply."""
    out = (C.c_double * 2)()
    check(lib().qmg_wilson_mf_deviation(C.byref(desc), out))
    return out[0]


def stencil_gamma5_deviation(desc):
    """Relative distance of the stored backward blocks from the gamma5-hermitian relation (0 for Wilson up to rounding)."""
    out = (C.c_double * 2)()
    check(lib().qmg_stencil_gamma5_deviation(C.byref(desc), out))
    return (out[0] / out[1]) ** 0.5 if out[1] > 0 else 0.0


def stencil_apply(desc, lhs, rhs, pieces=APPLY_ALL, dir_mask=15):
    check(lib().qmg_stencil_apply(C.byref(desc), C.c_int(pieces), C.c_int(dir_mask), ptr(lhs), ptr(rhs)))


def stencil_apply_host(desc, lhs_host, rhs_host, pieces=APPLY_ALL, dir_mask=15, dev_lhs=None, dev_rhs=None, rows_per_chunk=0):
    """lhs_host = M rhs_host for HOST vectors (numpy complex128 arrays or raw pointers), pipelined over row chunks."""
    def hp(a):
        return a if isinstance(a, C.c_void_p) else C.c_void_p(a.ctypes.data)
    check(lib().qmg_stencil_apply_host(C.byref(desc), C.c_int(pieces), C.c_int(dir_mask), hp(lhs_host), hp(rhs_host),
                                       ptr(dev_lhs), ptr(dev_rhs), C.c_int(rows_per_chunk)))


def stencil_apply_dot(desc, lhs, rhs, dot_with=None, pieces=APPLY_ALL):
    out = (C.c_double * 3)()
    check(lib().qmg_stencil_apply_dot(C.byref(desc), C.c_int(pieces), ptr(lhs), ptr(rhs), ptr(dot_with), out))
    return complex(out[0], out[1]), out[2]


def fill_wilson(X, Y, gauge, wilson_coeff=1.0):
    V = X * Y
    clover, hopping = cvec(V * 4), cvec(V * 16)
    check(lib().qmg_fill_wilson(X, Y, C.c_double(wilson_coeff), ptr(gauge), ptr(clover), ptr(hopping)))
    return clover, hopping


def fill_staggered(X, Y, gauge):
    hopping = cvec(X * Y * 4)
    check(lib().qmg_fill_staggered(X, Y, ptr(gauge), ptr(hopping)))
    return hopping


def fill_laplace(X, Y, gauge):
    clover, hopping = cvec(X * Y), cvec(X * Y * 4)
    check(lib().qmg_fill_laplace(X, Y, ptr(gauge), ptr(clover), ptr(hopping)))
    return clover, hopping


def fill_dwf(X, Y, Ls, gauge, mass, wilson_coeff=1.0):
    nc = 2 * Ls
    clover, hopping = cvec(X * Y * nc * nc), cvec(X * Y * nc * nc * 4)
    m = complex(mass)
    check(lib().qmg_fill_dwf(X, Y, Ls, C.c_double(wilson_coeff), C.c_double(m.real), C.c_double(m.imag), ptr(gauge), ptr(clover), ptr(hopping)))
    return clover, hopping


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def transfer_desc(Xf, Yf, ncf, Xc, Yc, ncc):
    d = TransferDesc()
    d.Xf, d.Yf, d.ncf, d.Xc, d.Yc, d.ncc = int(Xf), int(Yf), int(ncf), int(Xc), int(Yc), int(ncc)
    return d


def prolong(tdesc, nullvecs, coarse, fine):
    check(lib().qmg_prolong(C.byref(tdesc), _ptr_array(nullvecs), len(nullvecs), ptr(coarse), ptr(fine)))


def restrict(tdesc, nullvecs, fine, coarse):
    check(lib().qmg_restrict(C.byref(tdesc), _ptr_array(nullvecs), len(nullvecs), ptr(fine), ptr(coarse)))


def block_orthonormalize(tdesc, nullvecs, cholesky=None):
    check(lib().qmg_block_orthonormalize(C.byref(tdesc), _ptr_array(nullvecs), len(nullvecs), ptr(cholesky)))


def coarse_build(tdesc, fine_desc, prolong_vecs, restrict_vecs=None):
    Vc = tdesc.Xc * tdesc.Yc
    n2 = tdesc.ncc * tdesc.ncc
    clover, hopping = cvec(Vc * n2), cvec(4 * Vc * n2)
    rv = _ptr_array(restrict_vecs if restrict_vecs is not None else prolong_vecs)
    check(lib().qmg_coarse_build(C.byref(tdesc), C.byref(fine_desc), _ptr_array(prolong_vecs), rv, ptr(clover), ptr(hopping)))
    return clover, hopping


def dot(x, y):
    out = (C.c_double * 2)()
    check(lib().qmg_dot(ptr(x), ptr(y), C.c_long(x.numel()), out))
    return complex(out[0], out[1])


def norm2sq(x):
    out = C.c_double()
    check(lib().qmg_norm2sq(ptr(x), C.c_long(x.numel()), C.byref(out)))
    return out.value


def kernel_launches():
    return int(lib().qmg_kernel_launches())


# ---- y-slab sharding (include/qmg_b200.h "sharding"): one process per GPU, torch.distributed for the rendezvous ----
def comm_init(group=None):
    """Join the ring of slabs: rank 0 creates the NCCL id, torch.distributed (any backend) hands it round, and every
    rank calls qmg_comm_init.  After this every lattice given to the library is this rank's (X, Y / world_size) slab."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise QmgError("comm_init needs an initialised torch.distributed process group")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib().qmg_comm_unique_id(buf))
    ids = [bytes(buf.raw)]
    dist.broadcast_object_list(ids, src=0, group=group)
    check(lib().qmg_comm_init(world, rank, C.c_char_p(ids[0])))
    return world, rank


def comm_finalize():
    check(lib().qmg_comm_finalize())


def comm_set_loopback(on):
    check(lib().qmg_comm_set_loopback(1 if on else 0))


def comm_counters():
    l = lib()
    l.qmg_comm_halo_exchanges.restype = C.c_long
    l.qmg_comm_allreduces.restype = C.c_long
    l.qmg_comm_p2p_halo_exchanges.restype = C.c_long
    return dict(halo_exchanges=int(l.qmg_comm_halo_exchanges()), allreduces=int(l.qmg_comm_allreduces()),
                size=int(l.qmg_comm_size()), rank=int(l.qmg_comm_rank()), active=bool(l.qmg_comm_active()), p2p=bool(l.qmg_comm_p2p()), p2p_halo_exchanges=int(l.qmg_comm_p2p_halo_exchanges()))


def u1_ape_smear(gauge, X, Y, alpha, n_iter, textbook=False):
    out = cvec(2 * X * Y)
    check(lib().qmg_u1_ape_smear(ptr(out), ptr(gauge), C.c_int(X), C.c_int(Y), C.c_double(alpha), C.c_int(n_iter), C.c_int(1 if textbook else 0)))
    return out
