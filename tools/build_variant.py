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



def affected(defs):
    """translation units (or the headers they include) that mention one of the macros"""
    names = [d[2:].split("=")[0] for d in defs if d.startswith("-D")]
    headers = {h: open(os.path.join(B.CSRC, h)).read() for h in os.listdir(B.CSRC) if h.endswith((".cuh", ".h"))}
    hit_headers = {h for h, txt in headers.items() if any(n in txt for n in names)}
    out = []
    for src in B.SOURCES:
        txt = open(os.path.join(B.CSRC, src)).read()
        if any(n in txt for n in names) or any(f'"{h}"' in txt for h in hit_headers):
            out.append(src)
    return tuple(out)



def main():
    name, defs = sys.argv[1], sys.argv[2:]
    AFFECTED = affected(defs)
    print("rebuilding", AFFECTED)
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
