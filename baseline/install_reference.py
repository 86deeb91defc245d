#!/usr/bin/env python
"""Installs the UNMODIFIED reference (read-only at /root/reference) into baseline/_ref for `bench.py --impl reference`.

The reference ships no setup.py / pyproject.toml, so `pip install /root/reference` has nothing to build.  As the bench
contract allows, the install runs from a scratch copy under /tmp to which ONLY a packaging stub (setup.py) is added;
no source file of the reference is modified, and nothing of it enters the repository (baseline/_ref/ is git-ignored
but travels to the GPU box with the gpurun snapshot).

    python baseline/install_reference.py            # -> baseline/_ref/{optimizer.py, networks/, utils/, ...}
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
TARGET = os.path.join(HERE, "_ref")

SETUP = '''
from setuptools import setup, find_packages
setup(name="globalegomocap-reference", version="0.0.0",
      py_modules=["optimizer", "optimize_whole_sequence", "calculate_errors"],
      packages=find_packages(include=["networks", "networks.*", "utils", "utils.*"]),
      package_data={"utils.fisheye": ["*.json", "*.mat"]}, zip_safe=False)
'''


def install() -> str:
    """Returns 'installed', 'present' or 'unavailable: <why>'."""
    if os.path.exists(os.path.join(TARGET, "optimizer.py")):
        return "present"
    if not os.path.isdir(REF):
        return "unavailable: /root/reference does not exist here"
    tmp = tempfile.mkdtemp(prefix="gem_ref_install_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src)
        with open(os.path.join(src, "setup.py"), "w") as f:
            f.write(SETUP)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            return "unavailable: pip install failed: " + (res.stderr.strip().splitlines() or ["?"])[-1]
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install())
