"""Build the UNMODIFIED reference (arthurmensch/modl) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under modl_b200/ may import this.

The reference's hot path is Cython + NumPy.  Its own build system
(numpy.distutils, /root/reference/setup.py:22-38) no longer exists on
NumPy 2 / Python 3.12, so this recipe stages a scratch copy of the
reference tree under a temp dir, cythonizes the six .pyx modules with plain
setuptools (mirroring modl/decomposition/setup.py:11-20,
modl/utils/math/setup.py:11-15, modl/utils/randomkit/setup.py:15-28,
modl/input_data/setup.py), builds them in that scratch dir, and installs the
result (compiled .so + the reference's own .py files) into oracle/_ref/.

oracle/_ref/ is git-ignored (never enters history) but NOT gpurun-ignored,
so the compiled reference travels to the GPU box where /root/reference does
not exist.  `modl/__init__.py` and `modl/decomposition/__init__.py` are
installed EMPTY because the originals eagerly import nilearn/nibabel
(modl/__init__.py:1-4, modl/decomposition/fmri.py:18-21), absent here.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

SETUP_PY = r'''
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
inc = [numpy.get_include(), "modl/utils/randomkit"]
ext = [
    Extension("modl.decomposition.dict_fact_fast",
              ["modl/decomposition/dict_fact_fast.pyx"], include_dirs=inc),
    Extension("modl.decomposition.recsys_fast",
              ["modl/decomposition/recsys_fast.pyx"], include_dirs=inc),
    Extension("modl.utils.math.enet",
              ["modl/utils/math/enet.pyx"], include_dirs=inc),
    Extension("modl.utils.randomkit.random_fast",
              ["modl/utils/randomkit/random_fast.pyx",
               "modl/utils/randomkit/randomkit.c",
               "modl/utils/randomkit/distributions.c"],
              language="c++", include_dirs=inc),
    Extension("modl.utils.randomkit.sampler",
              ["modl/utils/randomkit/sampler.pyx"],
              language="c++", include_dirs=inc),
    Extension("modl.input_data.image_fast",
              ["modl/input_data/image_fast.pyx"], include_dirs=inc),
]
setup(name="modl_ref", ext_modules=cythonize(ext, language_level=3, quiet=True))
'''

# files of the reference that the hot path (and the "next" rows) need at run time
KEEP_PY = [
    "modl/decomposition/dict_fact.py",
    "modl/decomposition/recsys.py",
    "modl/decomposition/image.py",
    "modl/decomposition/fmri.py",
    "modl/decomposition/stability.py",
    "modl/utils/__init__.py",
    "modl/utils/math/__init__.py",
    "modl/utils/randomkit/__init__.py",
    "modl/utils/recsys/__init__.py",
    "modl/utils/recsys/cross_validation.py",
    "modl/input_data/image.py",
    "modl/feature_extraction/__init__.py",
    "modl/feature_extraction/image.py",
]
EMPTY_INIT = [
    "modl/__init__.py",
    "modl/decomposition/__init__.py",
    "modl/input_data/__init__.py",
]


def built(dest=DEST):
    need = [
        "modl/decomposition/dict_fact.py",
        "modl/feature_extraction/image.py",
    ]
    if not all(os.path.exists(os.path.join(dest, n)) for n in need):
        return False
    for pkg, stem in (("modl/decomposition", "dict_fact_fast"),
                      ("modl/utils/math", "enet"),
                      ("modl/utils/randomkit", "random_fast"),
                      ("modl/utils/randomkit", "sampler")):
        d = os.path.join(dest, pkg)
        if not os.path.isdir(d) or not any(
                f.startswith(stem + ".") and f.endswith(".so") for f in os.listdir(d)):
            return False
    return True


def build(reference="/root/reference", force=False, quiet=True):
    if built() and not force:
        return DEST
    if not os.path.isdir(os.path.join(reference, "modl")):
        raise FileNotFoundError(
            "reference tree not found at %s (oracle/_ref must be prebuilt)" % reference)
    tmp = tempfile.mkdtemp(prefix="modl_ref_build_")
    try:
        shutil.copytree(os.path.join(reference, "modl"), os.path.join(tmp, "modl"))
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(SETUP_PY)
        env = dict(os.environ)
        env.setdefault("CFLAGS", "-O2")
        out = subprocess.run([sys.executable, "setup.py", "build_ext", "-i"],
                             cwd=tmp, env=env, capture_output=quiet, text=True)
        if out.returncode != 0:
            raise RuntimeError("reference build failed:\n%s\n%s" % (out.stdout, out.stderr))
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        for rel in KEEP_PY:
            dst = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copy2(os.path.join(tmp, rel), dst)
        for rel in EMPTY_INIT:
            dst = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            open(dst, "w").close()
        for root, _, files in os.walk(os.path.join(tmp, "modl")):
            for fn in files:
                if fn.endswith(".so"):
                    rel = os.path.relpath(os.path.join(root, fn), tmp)
                    dst = os.path.join(DEST, rel)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copy2(os.path.join(root, fn), dst)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not built():
        raise RuntimeError("reference build incomplete")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(build(a.reference, a.force, quiet=False))
