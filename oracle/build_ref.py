"""oracle/build_ref.py -- TEST INFRASTRUCTURE.

Compiles the reference's OWN Cython sources of the hot path, from where they lie under
/root/reference, into extension modules under ``oracle/_ref/`` (git-ignored, travels to the GPU box).

  syconn/extraction/block_processing_C.pyx        -> oracle/_ref/block_processing_C*.so
  syconn/extraction/find_object_properties_C.pyx  -> oracle/_ref/find_object_properties_C*.so

Recipe = the reference's setup.py:8-16 (C++11, same Cython directives).  No reference source is copied
into the repository: cythonize runs on a scratch copy in a temp dir, only the built ``.so`` files are
kept.  One dead line is dropped from the scratch copy of find_object_properties_C.pyx (line 16, an
unused ``ctypedef vector[n_type[:, :, :]]`` that Cython >= 3 refuses to compile; no behaviour).
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SYK_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODS = ["block_processing_C", "find_object_properties_C"]

SETUP = """
from setuptools import setup, Extension
from Cython.Build import cythonize
exts = [Extension(m, [m + ".pyx"], extra_compile_args=["-std=c++11", "-O3", "-w"], language="c++") for m in %r]
setup(ext_modules=cythonize(exts, compiler_directives={
    'language_level': 3, 'boundscheck': False, 'wraparound': False, 'initializedcheck': False,
    'cdivision': False, 'overflowcheck': True}))
"""


def have_ref():
    return all(glob.glob(os.path.join(OUT, m + "*.so")) for m in MODS)


def build(force=False):
    if have_ref() and not force:
        return True
    src = os.path.join(REF, "syconn", "extraction")
    if not os.path.isdir(src):
        return False
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for m in MODS:
            with open(os.path.join(src, m + ".pyx")) as f:
                lines = f.readlines()
            if m == "find_object_properties_C":
                lines = [ln for ln in lines if "ctypedef vector[n_type[:, :, :]] uintarr_vec" not in ln]
            with open(os.path.join(tmp, m + ".pyx"), "w") as f:
                f.writelines(lines)
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(SETUP % (MODS,))
        subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp,
                              stdout=subprocess.DEVNULL)
        for so in glob.glob(os.path.join(tmp, "*.so")):
            shutil.copy(so, OUT)
    return have_ref()


if __name__ == "__main__":
    print("built" if build(force="--force" in sys.argv) else "reference not available")
