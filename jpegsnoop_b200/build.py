"""In-tree build of libjsgpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False):
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libjsgpu.so build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout)
    return os.path.join(HERE, "libjsgpu.so")
