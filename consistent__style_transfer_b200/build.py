"""Builds libwmd_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwmd_b200.so")
SOURCES = ["wmd_b200.cu"]
HEADERS = ["common.cuh", "nbow.cuh", "cost.cuh", "cost_fast.cuh", "solve.cuh", "solve_wide.cuh", "fused.cuh", "rwmd.cuh", "allpairs.cuh", "emd.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # float32 distances must round every op separately (numpy parity)
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def csrpack_path() -> str:
    import sysconfig
    return os.path.join(HERE, "_csrpack" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_csrpack(force: bool = False) -> str:
    """The host-side list -> CSR packer (csrc/csrpack.c, CPython C API, no CUDA)."""
    import sysconfig
    out, src = csrpack_path(), os.path.join(CSRC, "csrpack.c")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler for _csrpack")
    tmp = out + f".{os.getpid()}.tmp"
    res = subprocess.run([cc, "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], "-o", tmp, src],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building _csrpack failed: " + res.stderr[-400:])
    os.replace(tmp, out)
    return out


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libwmd_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [
        os.path.join(os.path.dirname(HERE), "include", "wmd_b200.h"), os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:                       # ranks of one job may all find the library stale
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():                          # another process built it while we waited
            return LIB
        tmp = LIB + f".{os.getpid()}.tmp"
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed building libwmd_b200.so")
        os.replace(tmp, LIB)                                       # never a half-written file at the final path
        with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as f:
            f.write(res.stdout + res.stderr)
    try:
        build_csrpack(force)
    except Exception as exc:                                       # the packer is an accelerator, not a dependency
        sys.stderr.write(f"[build] _csrpack not built: {exc}\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
