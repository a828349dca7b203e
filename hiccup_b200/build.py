"""Build libhiccup_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m hiccup_b200.build [--verbose]

The library is plain CUDA runtime code (no torch types); nvcc cross-compiles it without a GPU.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libhiccup_b200.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra_flags=()):
    """out / extra_flags: an experimental variant beside the product library (development aid; load it with
    HIC_LIB_PATH=...)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if out is not None:
        return _build_to(nvcc, out, verbose, list(extra_flags))
    if not force and not _stale():
        return LIB_PATH
    return _build_to(nvcc, LIB_PATH, verbose, [])


def _build_to(nvcc, LIB_PATH, verbose, extra):
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build %s" % LIB_PATH)
    cmd = [nvcc, *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr"]
    cmd += extra
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH + ".tmp", *sources()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libhiccup_b200.so")
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    if "--out" in sys.argv:           # python -m hiccup_b200.build --out path.so -DHIC_K1_TH=32 ...
        o = sys.argv[sys.argv.index("--out") + 1]
        print(build(out=o, verbose="--verbose" in sys.argv, extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
    else:
        print(build(force=True, verbose="--verbose" in sys.argv))
