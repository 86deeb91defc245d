#!/usr/bin/env python
"""Builds an experimental variant of libgem_b200.so with extra -D flags (measurement only):

    python tools/build_variant.py NAME -DGEM_PLANAR_W=16 -DGEM_PLANAR_H=16 ...

writes globalegomocap_b200/libgem_b200_NAME.so (select it with GEM_B200_LIB=<path>).  Only the translation units that
see the flags are rebuilt; the rest is linked from the regular build.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from globalegomocap_b200 import build as B  # noqa: E402

AFFECTED = ("api.cu", "energy.cu", "gemm_tap_tc.cu")


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    B.build()
    out_dir = os.path.join(B.HERE, "_build_var", name)
    os.makedirs(out_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        cmd = [B._nvcc(), *B.ARCH, *B.COMMON, *B.SOURCES[src], *defs, "-c", os.path.join(B.CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(3) as ex:
        var = dict(zip(AFFECTED, ex.map(compile_one, AFFECTED)))
    objs = [var.get(s, os.path.join(B.BUILD, s.replace(".cu", ".o"))) for s in B.SOURCES]
    lib = os.path.join(B.HERE, f"libgem_b200_{name}.so")
    subprocess.run([B._nvcc(), *B.ARCH, "-shared", "-o", lib, *objs], check=True)
    print(lib)


if __name__ == "__main__":
    main()
