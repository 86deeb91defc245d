#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS opcodes in libgem_b200.so (tcgen05 MMA / TMEM loads / TMA tensor
and bulk copies / tensor-core barriers), the evidence DESIGN.md section 5 quotes.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "globalegomocap_b200", "libgem_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCMMA|UTCBAR|UTCCP|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|"
                 r"ELECT|REDUX|CCTL)\b((?:\.[A-Z0-9_x]+)*)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    arch = None
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = (arch, collections.Counter(), [0])
            continue
        if name is None:
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            kernels[name][2][0] += 1
            m = PAT.search(line)
            if m:
                kernels[name][1][m.group(1) + m.group(2)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass globalegomocap_b200/libgem_b200.so: Blackwell-specific opcodes per kernel")
    print("# (UTCHMMA = tcgen05.mma kind::f16/tf32, .2CTA = cta_group::2; LDTM = tcgen05.ld; UTMALDG / UTMASTG = TMA tensor")
    print("#  load / store; UBLKCP = cp.async.bulk; UTCBAR = tcgen05.commit; SYNCS = mbarrier ops)")
    total = collections.Counter()
    for (mangled, (arch, cnt, n)), nice in zip(kernels.items(), demangle):
        if not cnt:
            continue
        short = nice.replace("(anonymous namespace)::", "").replace("gem::", "")
        short = re.sub(r"\(.*$", "", short).replace("void ", "")
        print(f"\n{short}   [{arch}, {n[0]} instructions]")
        for k, v in sorted(cnt.items()):
            print(f"    {k:28s} {v}")
            total[k.split('.')[0]] += v
    print("\n# totals by base opcode")
    for k, v in sorted(total.items()):
        print(f"    {k:28s} {v}")
    archs = sorted({a for a, _, _ in kernels.values()})
    print("\n# cubin architectures:", ", ".join(archs), f"({len(kernels)} kernels)")


if __name__ == "__main__":
    sys.exit(main())
