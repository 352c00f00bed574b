"""CPU suite, part 2: the drop-in boundary.  The C-ABI library loads without a GPU, exports every symbol that
include/qmg_b200.h declares, refuses to compute without a device (no CPU fallback), and the host-class library exports
the same flat driver API as the oracle with identical host-side (Lattice2D) answers."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import capi
import latutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libqmg_b200_exports_header_symbols():
    import qmg
    lib = qmg.lib()
    syms = qmg.exported_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_header_cites_reference_for_every_entry():
    text = open(os.path.join(ROOT, "include", "qmg_b200.h")).read()
    assert text.count("stencil_2d.h") >= 8 and "transfer.h" in text and "coarse.h" in text and "extern \"C\"" in text
    assert "torch" not in text.lower().replace("torch's current stream", "")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import qmg
    lib = qmg.lib()
    assert lib.qmg_device_count() == 0
    x = np.zeros(8, np.complex128)
    out = C.c_double()
    rc = lib.qmg_norm2sq(x.ctypes.data_as(C.c_void_p), C.c_long(8), C.byref(out))
    assert rc != 0
    assert b"no CPU fallback" in lib.qmg_last_error()
    with pytest.raises(qmg.QmgError):
        qmg.init(0)


def test_product_does_not_touch_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may execute oracle/."""
    for base in ("quantum-mg_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if not f.endswith((".py", ".h", ".cuh", ".cu", ".cpp")):
                    continue
                for line in open(os.path.join(dirpath, f), errors="ignore"):
                    code = line.split("//")[0]
                    if "#include" in code or "import " in code or "CDLL" in code or "dlopen" in code:
                        assert "oracle" not in code and "qlinalg_shim" not in code and "libqmg_ref" not in code, (f, line)
    out = subprocess.run(["ldd", os.path.join(ROOT, "quantum-mg_b200", "libqmg_host.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "libqmg_ref" not in out and "libqmg_b200" in out


def _exports(path, prefix):
    out = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True).stdout
    return sorted(set(m.group(1) for m in re.finditer(r" T " + prefix + r"(\w+)", out)))


@pytest.mark.skipif(not capi.have_ref(), reason="oracle/_ref not built")
def test_host_library_mirrors_oracle_driver_api():
    ref = _exports(capi.REF_LIB, "ref_")
    gpu = _exports(capi.GPU_LIB, "qmgh_")
    assert ref == gpu and len(gpu) > 45


@pytest.mark.skipif(not capi.have_ref(), reason="oracle/_ref not built")
def test_lattice2d_parity_without_gpu():
    """Lattice2D is host-only integer geometry: the product class must agree with the reference's for every index map."""
    ref, gpu = capi.Backend("ref"), capi.Backend("gpu")
    for X, Y, nc in ((6, 4, 2), (8, 8, 8), (2, 2, 1), (1, 1, 4)):
        a, b = ref.lattice(X, Y, nc), gpu.lattice(X, Y, nc)
        assert (a.volume, a.size_cv, a.size_cm, a.size_gauge, a.size_hopping, a.size_corner) == (b.volume, b.size_cv, b.size_cm, b.size_gauge, b.size_hopping, b.size_corner)
        xyc = (C.c_int * 3)()
        for i in range(X * Y):
            assert a.index_to_coord(i) == b.index_to_coord(i)
        for x in range(X):
            for y in range(Y):
                assert a.coord_to_index(x, y) == b.coord_to_index(x, y)
                for fn, args in (("lattice_cm_index", (x, y, nc - 1, 0)), ("lattice_hopping_index", (x, y, 0, nc - 1, 3)), ("lattice_gauge_index", (x, y, 0, 0, 1))):
                    assert ref.fn(fn)(a.h, *args) == gpu.fn(fn)(b.h, *args)
        for i in range(0, a.size_cv, 3):
            ref.fn("lattice_cv_index_to_coord")(a.h, i, xyc); ra = tuple(xyc)
            gpu.fn("lattice_cv_index_to_coord")(b.h, i, xyc)
            assert ra == tuple(xyc)


def test_bench_reference_arm_runs():
    """bench.py --impl reference times the reference's own CPU apply and prints one JSON line with the contract keys."""
    if not capi.have_ref():
        pytest.skip("oracle/_ref not built")
    import json
    out = subprocess.run(["python", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-L", "256", "--cpu-reps", "2", "--kcycle-L", "0", "--cpu-kcycle-L", "0"],
                         stdout=subprocess.PIPE, text=True, timeout=600).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "GB/s"


def test_unmodified_reference_drivers_compile_against_the_b200_headers():
    """The drop-in claim of INTEGRATION.md, checked where the reference tree is present (this container): every test driver
    the reference ships parses UNMODIFIED against include/qmg (same class names, members, free functions and solver entry
    points).  Excluded: n02, whose own local header is stale against the reference's Stencil2D as well."""
    ref_tests = "/root/reference/tests"
    if not os.path.isdir(ref_tests):
        pytest.skip("reference tree not present on this machine")
    failures = []
    for d in sorted(os.listdir(ref_tests)):
        if not d.startswith("n") or d.startswith("n02"):
            continue
        for f in sorted(os.listdir(os.path.join(ref_tests, d))):
            if not f.endswith(".cpp"):
                continue
            src = os.path.join(ref_tests, d, f)
            r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-I" + os.path.join(ROOT, "include", "qmg"),
                                "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ref_tests, d), src],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                failures.append((d, r.stdout[-600:]))
    assert not failures, failures


def test_c_abi_header_is_plain_c(tmp_path):
    """include/qmg_b200.h is the drop-in boundary: it must be consumable from C (cgo / JNI / ctypes style bindings), i.e. parse
    as strict C99 with no C++ or torch types, and a C translation unit using it must link against libqmg_b200.so."""
    src = tmp_path / "use_abi.c"
    src.write_text('#include "qmg_b200.h"\n#include <stdio.h>\nint main(void) { printf("%d %d\\n", qmg_device_count(), qmg_comm_size()); return 0; }\n')
    exe = tmp_path / "use_abi"
    lib_dir = os.path.join(ROOT, "quantum-mg_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src),
                        "-L" + lib_dir, "-lqmg_b200", "-Wl,-rpath," + lib_dir, "-o", str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.split()[1] == "1"


def test_ctypes_stencil_desc_matches_the_c_header(tmp_path):
    """qmg.StencilDesc, and the copy a binder would paste from INTEGRATION.md section 3, mirror qmg_stencil_desc field for field
    (a binder that omits the trailing gamma5_hermitian / hop_halo_ym hands the library 12 bytes of garbage)."""
    import ctypes as C
    import qmg
    hdr = open(os.path.join(ROOT, "include", "qmg_b200.h")).read()
    body = re.search(r"typedef struct qmg_stencil_desc\s*\{(.*?)\}\s*qmg_stencil_desc;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [re.sub(r"\[.*\]", "", n).strip(" *") for n in decl.split(None, 1)[1].replace("qmg_cplx*", "").replace("const", "").split(",")]
    names = [n.split()[-1] for n in names]
    assert names == [f[0] for f in qmg.StencilDesc._fields_]
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    snippet = doc[doc.index("class StencilDesc(C.Structure)"):doc.index("desc = StencilDesc()")]
    assert re.findall(r'\("(\w+)",', snippet) == names
    # same size and offsets as the C compiler sees them
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qmg_b200.h"\nint main(void) { printf("%zu", sizeof(qmg_stencil_desc));\n'
                   + "".join('printf(" %%zu", offsetof(qmg_stencil_desc, %s));\n' % n for n in names) + "return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert got[0] == C.sizeof(qmg.StencilDesc)
    assert got[1:] == [getattr(qmg.StencilDesc, n).offset for n in names]
