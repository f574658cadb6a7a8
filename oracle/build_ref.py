"""Build the LIVE reference oracle: the reference's own Cython module
pdspy/interferometry/libinterferometry.pyx, compiled where it lies under
/root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).

TEST INFRASTRUCTURE ONLY.  Nothing in the product (pdspy_b200/) imports this.
No reference source is copied into the repo: the .pyx is read in place, the
generated .c and the built .so land in oracle/_ref/ only.

The .pyx imports h5py and astropy at module scope (libinterferometry.pyx:3-4)
but only uses them in read/write/asFITS; two empty stub modules written into
oracle/_ref/stubs/ satisfy the import.  Flags follow the reference's setup.py:8-13
(-ffast-math, -lm, NPY_NO_DEPRECATED_API=0).
"""
import os, subprocess, sys, sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("PDSPY_REFERENCE", "/root/reference")
PYX = os.path.join(REF, "pdspy", "interferometry", "libinterferometry.pyx")


def so_path():
    return os.path.join(OUT, "libinterferometry" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    """Returns the path of the built module, or None if the reference tree is absent."""
    if os.path.exists(so_path()) and not force:
        return so_path()
    if not os.path.exists(PYX):
        return None
    import numpy
    os.makedirs(os.path.join(OUT, "stubs"), exist_ok=True)
    for stub in ("h5py", "astropy"):
        with open(os.path.join(OUT, "stubs", stub + ".py"), "w") as f:
            f.write("# empty stub so the live reference oracle imports (test infrastructure)\n")
    cfile = os.path.join(OUT, "libinterferometry.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", PYX, "-o", cfile])
    inc = sysconfig.get_paths()["include"]
    cmd = ["gcc", "-shared", "-fPIC", "-O2", "-ffast-math", "-DNPY_NO_DEPRECATED_API=0",
           "-I", inc, "-I", numpy.get_include(), cfile, "-o", so_path(), "-lm"]
    subprocess.check_call(cmd)
    return so_path()


def load():
    """Import the live reference module (building it if possible). None if unavailable."""
    p = build()
    if p is None:
        return None
    import importlib.util
    stubs = os.path.join(OUT, "stubs")
    if stubs not in sys.path:
        sys.path.append(stubs)
    spec = importlib.util.spec_from_file_location("libinterferometry", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
