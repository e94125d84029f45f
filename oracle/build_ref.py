"""Build oracle/_ref/: the reference's own cython_kmer_index.pyx compiled as-is (TEST INFRASTRUCTURE).

Reads /root/reference/graph_kmer_index/cython_kmer_index.pyx where it lies; generated C and the
extension module go to oracle/_ref/ only.  No reference source is copied into the repository.
Skips silently when the reference is absent (GPU box) -- the prebuilt .so travels with gpurun.
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PYX = "/root/reference/graph_kmer_index/cython_kmer_index.pyx"


def build(force=False):
    if not os.path.exists(PYX):
        return False
    os.makedirs(OUT, exist_ok=True)
    existing = glob.glob(os.path.join(OUT, "cython_kmer_index*.so"))
    if existing and not force and os.path.getmtime(existing[0]) >= os.path.getmtime(PYX):
        return True
    import numpy as np
    c_file = os.path.join(OUT, "cython_kmer_index.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", PYX, "-o", c_file])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "cython_kmer_index" + ext)
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", inc, "-I", np.get_include(),
                           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", c_file, "-o", so])
    os.remove(c_file)   # generated from the reference: do not keep a derived copy of its source around
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "reference absent: oracle/_ref not built")
