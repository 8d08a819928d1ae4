"""zinc_b200/build.py -- in-tree build of libzipgpu.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m zinc_b200.build [--force] [--verbose]

Objects go to zinc_b200/csrc/_build/, the library to zinc_b200/libzipgpu.so (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libzipgpu.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr",
           "-I", INCLUDE]
SOURCES = ["raa_encode.cu", "commit_ws.cu", "commit_ws16k.cu", "commit_wsc.cu", "raa_big.cu", "encode_f.cu", "merkle.cu", "open_columns.cu", "combine_rows.cu", "sparse_encode.cu", "sparse_umma.cu", "peer_roots.cu", "microbench.cu", "zipgpu.cu", "mgpu.cu", "rand_compat.cpp"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(INCLUDE, "zipgpu.h"), os.path.abspath(__file__)]


def _stale(target: str, srcs: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src: str, verbose: bool, objdir: str = OBJ, extra: tuple = ()) -> str:
    obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if _stale(obj, [path] + _deps()):
        cmd = [NVCC, *ARCH, *CUFLAGS, *extra, "-c", path, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """an A/B build for experiments: zinc_b200/libzipgpu_<name>.so compiled with extra -D flags (load it with ZIPGPU_LIB)"""
    objdir = os.path.join(CSRC, "_build_" + name)
    os.makedirs(objdir, exist_ok=True)
    extra = tuple("-D" + d for d in defines)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, False, objdir, extra), SOURCES))
    lib = os.path.join(HERE, f"libzipgpu_{name}.so")
    r = subprocess.run([NVCC, *ARCH, "-shared", "-o", lib, *objs, "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
