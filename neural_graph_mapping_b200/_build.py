"""Build libngm_b200.so in-tree with nvcc for sm_100a (no torch dependency, no JIT cache)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libngm_b200.so")
STAMP = os.path.join(PKG_DIR, ".libngm_b200.stamp")
# the same sources with -DNGM_DEBUG_EXPORTS: adds the ngm_debug_* entry points (include/ngm_b200_debug.h) and the
# traced twin of the tcgen05 kernel; used by tests/tools only, never by the package
DEBUG_LIB_PATH = os.path.join(PKG_DIR, "libngm_b200_debug.so")
DEBUG_STAMP = os.path.join(PKG_DIR, ".libngm_b200_debug.stamp")

SOURCES = ["abi.cu", "sampler.cu", "composite.cu", "composite_bwd.cu", "encode.cu", "adam.cu", "targets.cu", "field_simt.cu", "field_tc.cu", "field_tc_bwd.cu", "knn.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _source_hash(extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "ngm_b200.h"), os.path.join(INCLUDE, "ngm_b200_debug.h")]
    for f in files:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())  # not the absolute path: the tree is copied to other boxes
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Compile every CUDA source to objects and link libngm_b200.so next to the package
    (``debug``: libngm_b200_debug.so, the same sources with -DNGM_DEBUG_EXPORTS)."""
    lib_path, stamp = (DEBUG_LIB_PATH, DEBUG_STAMP) if debug else (LIB_PATH, STAMP)
    defines = ["-DNGM_DEBUG_EXPORTS"] if debug else []
    want = _source_hash(" ".join(defines))

    def fresh() -> bool:
        return os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read() == want

    if not force and fresh():
        return lib_path
    # One builder at a time per tree: under torch.distributed.run every rank calls build(); the others wait on the
    # lock and then find the stamp fresh (concurrent nvcc runs into one object directory corrupt each other's files).
    import fcntl

    with open(lib_path + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return lib_path
            return _build_locked(lib_path, stamp, defines, want, debug, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(lib_path: str, stamp: str, defines, want: str, debug: bool, verbose: bool) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build_debug" if debug else "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(obj)
    tmp_lib = lib_path + f".tmp{os.getpid()}"  # link beside the target, then rename: a reader never sees a partial file
    link = [nvcc, "-shared", "-o", tmp_lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp_lib, lib_path)
    open(os.path.join(objdir, "build.log"), "w").write("\n".join(log))
    open(stamp, "w").write(want)
    if verbose:
        print("\n".join(log))
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv))
