#!/usr/bin/env python3
"""Builds variants of libb2deflate.so for A/B runs on the GPU box: tools/variants.py <file.cu> NAME=-DFLAG[,-DFLAG2] ...
Each variant recompiles one source with the extra flags and links it with the other objects of the normal build into
deflate-library-java_b200/build/variants/libb2d_NAME.so (git-ignored, travels with gpurun).  The probes take B2D_SO=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "deflate-library-java_b200"))
import build as B
B.build()
src = sys.argv[1]
vdir = os.path.join(B.OBJ, "variants")
os.makedirs(vdir, exist_ok=True)
procs = []
for spec in sys.argv[2:]:
    name, _, flags = spec.partition("=")
    obj = os.path.join(vdir, f"{src}.{name}.o")
    cmd = [B.NVCC] + B.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-c", os.path.join(B.CSRC, src), "-o", obj]
    procs.append((name, obj, subprocess.Popen(cmd)))
for name, obj, p in procs:
    assert p.wait() == 0, name
    objs = [obj if f == src else os.path.join(B.OBJ, f + ".o") for f in B.CU + B.C]
    out = os.path.join(vdir, f"libb2d_{name}.so")
    subprocess.check_call([B.NVCC, "-shared", "-o", out] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL", "-Wno-deprecated-gpu-targets"])
    print(out)
