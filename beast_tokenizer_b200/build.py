"""Build libbeast_b200.so (sm_100a only) in-tree with nvcc.

    python -m beast_tokenizer_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/beast_b200.h); it links the static
CUDA runtime, so it has no dependency on torch.  nvcc cross-compiles without a GPU.
"""
import glob
import os
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbeast_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


PYLISTS_SRC = os.path.join(CSRC, "pylists.c")
PYLISTS = os.path.join(PKG, "_pylists" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pylists(force=False):
    """The CPython helper of the list-returning BPE API (CSR <-> List[List[int]]): plain C, gcc, in-tree."""
    if not force and os.path.exists(PYLISTS) and os.path.getmtime(PYLISTS) >= os.path.getmtime(PYLISTS_SRC):
        return PYLISTS
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-Wall", "-I", sysconfig.get_paths()["include"],
           PYLISTS_SRC, "-o", PYLISTS]
    subprocess.check_call(cmd)
    return PYLISTS


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    build_pylists(force)
    if not force and not _stale():
        return LIB
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            print(f"--- {os.path.basename(src)}\n{out}", file=sys.stderr)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
