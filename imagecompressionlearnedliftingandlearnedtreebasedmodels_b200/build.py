"""In-tree build of the sm_100a CUDA library (``libll_b200.so``) with nvcc.

The library is a plain C-ABI shared object (``include/ll_api.h``); it does not link
against torch.  ``nvcc`` cross-compiles without a GPU, so this runs in the dev
container; the built ``.so`` is git-ignored but travels to the GPU box.
"""
import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libll_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(os.path.dirname(PKG_DIR), "include", "ll_api.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a into ``libll_b200.so``.  Returns the path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libll_b200.so")
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(PKG_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
