"""Build the REAL reference `algos` extension (test infrastructure only).

Compiles /root/reference/graphormer/algos.pyx *unmodified* with Cython into
oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  The
source is read where it lies; only build outputs land in oracle/_ref/.  No
reference source is copied into the repository history.

Needs language_level=2 (algos.pyx:15 uses the bare Python-2 name `long`).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import the result.
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/graphormer/algos.pyx"
OUT = os.path.join(HERE, "_ref")


def have_ref():
    if not os.path.isdir(OUT):
        return False
    return any(f.startswith("algos") and f.endswith(".so") for f in os.listdir(OUT))


def build(force=False):
    if have_ref() and not force:
        return True
    if not os.path.exists(REF_PYX):
        return False
    import numpy
    from setuptools import setup, Extension
    from Cython.Build import cythonize

    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="mobgt_ref_")
    try:
        # the reference tree is read-only: cythonize from a scratch copy in /tmp
        pyx = os.path.join(tmp, "algos.pyx")
        shutil.copy(REF_PYX, pyx)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            setup(
                name="algos",
                ext_modules=cythonize(
                    [Extension("algos", ["algos.pyx"], include_dirs=[numpy.get_include()],
                               extra_compile_args=["-O2"])],
                    language_level=2, quiet=True),
                script_args=["-q", "build_ext", "--build-lib", OUT, "--build-temp", os.path.join(tmp, "bt")],
            )
        finally:
            os.chdir(cwd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return have_ref()


def load():
    """Import the compiled reference module, or return None if it was never built."""
    if not have_ref():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import importlib
    return importlib.import_module("algos")


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok, os.listdir(OUT) if ok else "")
