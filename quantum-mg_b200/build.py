"""In-tree build of the sm_100a shared libraries (no JIT cache, no torch extension).

  libqmg_b200.so   csrc/*.cu       kernels + the C ABI of include/qmg_b200.h
  libqmg_host.so   host/*.cpp      the reference-API host classes (include/qmg/) behind
                                   the flat driver API (host/qmg_capi_body.h), linked to libqmg_b200.so

nvcc cross-compiles without a GPU; the .so files travel to the GPU box with the tree.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--extended-lambda",
              "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def build_kernels(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
    hdrs = glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(ROOT, "include", "qmg_b200.h")]
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            out = _run([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
            if verbose:
                print(out)
    lib = os.path.join(HERE, "libqmg_b200.so")
    if force or _newer(lib, objs):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ldl"])
    return lib


def build_host(force=False):
    srcs = sorted(glob.glob(os.path.join(HERE, "host", "*.cpp")))
    if not srcs:
        return None
    deps = srcs + glob.glob(os.path.join(HERE, "host", "*.h")) + glob.glob(os.path.join(ROOT, "include", "qmg", "*", "*.h")) \
        + [os.path.join(ROOT, "include", "qmg_b200.h")]
    lib = os.path.join(HERE, "libqmg_host.so")
    if force or _newer(lib, deps + [os.path.join(HERE, "libqmg_b200.so")]):
        cxx = shutil.which("g++") or "g++"
        _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "include"),
              "-I" + os.path.join(ROOT, "include", "qmg")] + srcs +
             ["-L" + HERE, "-lqmg_b200", "-Wl,-rpath,$ORIGIN", "-o", lib])
    return lib


def build_all(force=False, verbose=False):
    libs = [build_kernels(force, verbose)]
    h = build_host(force)
    if h:
        libs.append(h)
    return libs


if __name__ == "__main__":
    print("\n".join(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)))
