"""Builds libb2deflate.so (the C-ABI product library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
Run: python deflate-library-java_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb2deflate.so")
OBJ = os.path.join(HERE, "build")
CU = ["api.cu", "inflate.cu", "deflate.cu", "crc32.cu"]
C = ["corpus.c"]
HDRS = ["common.cuh", "kernels.h", os.path.join("..", "..", "include", "b2deflate.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "--use_fast_math"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HDRS]
    objs = []
    procs = []
    for f in CU + C:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            if f.endswith(".cu"):
                cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            else:
                cmd = ["gcc", "-O2", "-fPIC", "-fvisibility=hidden", "-std=gnu99", "-c", src, "-o", obj]
            procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"== {f}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("libb2deflate build failed")
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL",
                                                       "-Wno-deprecated-gpu-targets"]
        subprocess.check_call(cmd)
    build_host(force)
    return OUT


def build_host(force=False):
    """The host-side mirror of the reference's CLIs (host/gzip.cpp, host/gunzip.cpp) -> bin/gzip, bin/gunzip."""
    host = os.path.join(HERE, "host")
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    deps = [os.path.join(host, "b2d_streams.hpp"), os.path.join(HERE, "..", "include", "b2deflate.h"), OUT]
    for name in ("gzip", "gunzip", "zpipe", "stream_tests"):
        src = os.path.join(host, name + ".cpp")
        exe = os.path.join(bindir, name)
        if force or _stale(exe, [src] + deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, src, "-L" + HERE, "-lb2deflate",
                                   "-Wl,-rpath,$ORIGIN/..", "-pthread"])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
