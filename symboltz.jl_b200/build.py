"""Problem-build-time compilation: sympy model -> generated CUDA header -> nvcc (sm_100a) -> in-tree shared library.

Mirrors what `CosmologyProblem(M, pars)` does in the reference (src/solve.jl:129-236: symbolic compilation, "expensive,
do not repeat", docs/src/solve.md:26-28): one shared library per model structure (lmax, nx, w0wa), cached under
symboltz.jl_b200/_build/<key>/ and reused for every parameter set."""
import json
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "_build")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the B200 path cannot be built (there is no CPU fallback)")


def _digest(sources):
    import hashlib
    h = hashlib.sha1()
    for s in sources:
        with open(s, "rb") as f:
            h.update(hashlib.sha1(f.read()).digest())
    return h.hexdigest()


def _newer(target, sources):
    """Is `target` an up-to-date build of `sources`?  Decided by a content hash stamped next to the target (file times do not
    survive every way a tree gets copied to a GPU box; a needless rebuild there costs minutes of GPU time), with the
    modification times as the fallback for targets built before the stamps existed."""
    if not os.path.exists(target):
        return False
    stamp = target + ".stamp"
    if os.path.exists(stamp):
        return open(stamp).read().strip() == _digest(sources)
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _stamp(target, sources):
    with open(target + ".stamp", "w") as f:
        f.write(_digest(sources))


_FRESH = {}


def model_key(lmax, nx, w0wa):
    return f"l{lmax}_x{nx}_{'w0wa' if w0wa else 'lcdm'}"


def build_model(lmax=10, nx=4, w0wa=False, force=False, verbose=False):
    """Generate + compile the per-model engine. Returns (path to .so, info dict)."""
    key = model_key(lmax, nx, w0wa)
    if not force and key in _FRESH:  # checked once per process: parameter sweeps construct thousands of problems of one model
        return _FRESH[key]
    d = os.path.join(BUILD_DIR, key)
    so = os.path.join(d, f"libsbm_{key}.so")
    hdr = os.path.join(d, "sb_model_gen.h")
    meta = os.path.join(d, "info.json")
    gen_srcs = [os.path.join(_HERE, "codegen", f) for f in ("model.py", "lower.py")]
    c_srcs = [os.path.join(CSRC, f) for f in ("sb_engine.cu", "sb_debug.cpp", "sb_rodas.h")]
    if not force and _newer(so, gen_srcs + c_srcs) and os.path.exists(meta):
        _FRESH[key] = (so, json.load(open(meta)))
        return _FRESH[key]
    os.makedirs(d, exist_ok=True)
    if force or not (_newer(hdr, gen_srcs) and os.path.exists(meta)):
        from .codegen.model import Model
        from .codegen.lower import generate
        text, info = generate(Model(lmax=lmax, nx=nx, w0wa=w0wa))
        info.pop("L", None)
        with open(hdr, "w") as f:
            f.write(text)
        json.dump(info, open(meta, "w"))
        _stamp(hdr, gen_srcs)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", d, "-I", CSRC, "-o", so, c_srcs[0], c_srcs[1]]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    _stamp(so, gen_srcs + c_srcs)
    _FRESH[key] = (so, json.load(open(meta)))
    return _FRESH[key]


def build_los(force=False, verbose=False):
    so = os.path.join(BUILD_DIR, "libsbl.so")
    src = os.path.join(CSRC, "sb_los.cu")
    if not force and _newer(so, [src]):
        return so
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", so, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    _stamp(so, [src])
    return so


def build_comm(force=False, verbose=False):
    """libsbc.so: NCCL communicator + exchange steps behind the C ABI (csrc/sb_comm.cu).  Links the NCCL the process will see at run time
    (soname libnccl.so.2: torch's bundled copy when torch is loaded first, the system copy for a plain C / Julia host)."""
    so = os.path.join(BUILD_DIR, "libsbc.so")
    src = os.path.join(CSRC, "sb_comm.cu")
    if not force and _newer(so, [src]):
        return so
    os.makedirs(BUILD_DIR, exist_ok=True)
    inc, lib = "/usr/include", "/usr/lib/x86_64-linux-gnu"
    try:
        import nvidia.nccl as _n  # torch's bundled NCCL (same soname), preferred when present
        base = list(_n.__path__)[0]
        if os.path.exists(os.path.join(base, "include", "nccl.h")):
            inc, lib = os.path.join(base, "include"), os.path.join(base, "lib")
    except Exception:
        pass
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", inc, "-o", so, src, "-L", lib, "-l:libnccl.so.2", "-Xlinker", "-rpath," + lib]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    _stamp(so, [src])
    return so


# BASELINE configs: 1/2 (ΛCDM lmax 10), the reference test fixture (lmax 5), 4 (w0waCDM), 3 (massive-ν momentum grids nx = 8 and 16), the reference's "High lmax" test (32)
DEFAULT_MODELS = [dict(lmax=10, nx=4, w0wa=False), dict(lmax=5, nx=4, w0wa=False), dict(lmax=10, nx=4, w0wa=True), dict(lmax=10, nx=8, w0wa=False), dict(lmax=32, nx=4, w0wa=False),
                  dict(lmax=10, nx=16, w0wa=False)]


def build_all(force=False, verbose=False, jobs=None):
    """Everything the tests and the bench load, the model engines in parallel (each is one nvcc process of 0.5-2 minutes)."""
    import concurrent.futures as cf
    out = [build_los(force=force, verbose=verbose), build_comm(force=force, verbose=verbose)]
    jobs = jobs or min(len(DEFAULT_MODELS), os.cpu_count() or 1)
    with cf.ThreadPoolExecutor(jobs) as pool:
        out += [r[0] for r in pool.map(lambda m: build_model(force=force, verbose=verbose, **m), DEFAULT_MODELS)]
    return out


if __name__ == "__main__":  # python -m symboltz.jl_b200.build [--lmax 10 --nx 4 --w0wa] : what a non-Python host runs at CosmologyProblem time
    import argparse
    ap = argparse.ArgumentParser(description="generate + compile the per-model engine (cached); prints the path of the shared library")
    ap.add_argument("--lmax", type=int, default=10)
    ap.add_argument("--nx", type=int, default=4)
    ap.add_argument("--w0wa", action="store_true")
    ap.add_argument("--all", action="store_true", help="build the default model set and the line-of-sight library")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if a.all:
        print("\n".join(build_all(force=a.force, verbose=True)))
    else:
        print(build_model(a.lmax, a.nx, a.w0wa, force=a.force, verbose=True)[0])
        print(build_los(force=a.force))
