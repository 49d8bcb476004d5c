#!/usr/bin/env python
"""Builds the UNMODIFIED reference (douban/dpark) into the git-ignored baseline/_ref/ so that
`bench.py --impl reference` can time the reference's own shuffle on the GPU box's host cores.

Test/measurement infrastructure, not product code: nothing under dpark_b200/ imports it.

Recipe (verified there and by tests/golden/make_golden.py):
  1. copy /root/reference/dpark                 -> baseline/_ref/dpark          (sources stay out of git: .gitignore)
  2. cythonize -2 dpark/portable_hash.pyx       (language_level 2: the .pyx uses `long`)
  3. gcc dpark/utils/crc32c.c + crc32c_mod.c    -> dpark/utils/crc32c<EXT_SUFFIX>  (the package import needs it)
The four modules absent from this image (addict, pymesos, lz4framed, dpark.utils.recursion) are injected
as stubs at run time by oracle/ref_runner.py; no reference file is edited.

`pip install --target baseline/_ref /root/reference` (the base contract's form) is tried first and kept when it
works; in this image it fails (setup.py demands pymesos/addict/lz4framed wheels that the wheelhouse lacks and
cythonizes with language_level 3, on which portable_hash.pyx does not compile), so the explicit recipe runs.
Only runs where /root/reference exists (the build container); the GPU box uses the prebuilt tree that
travels with gpurun.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DPARK_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def built():
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return (os.path.exists(os.path.join(DST, "dpark", "portable_hash" + ext)) and
            os.path.exists(os.path.join(DST, "dpark", "utils", "crc32c" + ext)))


def build(force=False):
    if not os.path.exists(os.path.join(REF, "dpark", "portable_hash.pyx")):
        return built()
    if built() and not force:
        return True
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, "dpark"), os.path.join(DST, "dpark"))
    subprocess.check_call(["cythonize", "-i", "-2", "dpark/portable_hash.pyx"], cwd=DST,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    inc = sysconfig.get_paths()["include"]
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-msse4.2", "-I" + inc, "dpark/utils/crc32c.c",
                           "dpark/utils/crc32c_mod.c", "-o", "dpark/utils/crc32c" + ext], cwd=DST)
    # build leftovers that are not needed at run time
    for junk in ("dpark/portable_hash.c", "build"):
        p = os.path.join(DST, junk)
        if os.path.isdir(p):
            shutil.rmtree(p, ignore_errors=True)
        elif os.path.exists(p):
            os.unlink(p)
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("baseline/_ref %s" % ("ready" if ok else "NOT built (reference absent)"))
    sys.exit(0 if ok else 1)
