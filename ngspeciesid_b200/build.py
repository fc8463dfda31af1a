"""Builds ngspeciesid_b200/libngsid.so (CUDA kernels + C ABI) for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ngsid_api.cu")
OUT = os.path.join(HERE, "libngsid.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--fmad=false", "-Xcompiler", "-fPIC,-fopenmp", "-shared", "-Xptxas", "-v", "-ldl", "-lgomp"]


def sources():
    d = os.path.join(HERE, "csrc")
    return ([os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh", ".h"))]
            + [os.path.join(HERE, "..", "include", "ngsid.h")])


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build_library(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT, SRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for libngsid.so")
    logdir = os.path.join(HERE, "..", "build")          # git-ignored; the register / spill report of the last build
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, "ptxas.log"), "w") as f:
        f.write(res.stdout)
    return OUT


HOSTPACK_SRC = os.path.join(HERE, "csrc", "hostpack.c")


def hostpack_path():
    import sysconfig
    return os.path.join(HERE, "_hostpack" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_hostpack(force=False):
    """The CPython helper of the modules/ layer (csrc/hostpack.c: tuples of str -> byte arrays), gcc."""
    import sysconfig
    out = hostpack_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(HOSTPACK_SRC):
        return out
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"],
           "-o", out, HOSTPACK_SRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("gcc failed for _hostpack")
    return out


def load_hostpack():
    """The _hostpack module, compiled on first use when __graft_entry__.build() has not been run (gcc, 1 s)."""
    import importlib
    build_hostpack()
    return importlib.import_module("ngspeciesid_b200._hostpack")


if __name__ == "__main__":
    build_library(force=True, verbose=True)
    build_hostpack(force=True)
    print(OUT)
