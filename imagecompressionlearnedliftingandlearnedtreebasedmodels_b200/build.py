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
# measurement / unit probes (include/ll_probe.h): test and bench tooling, kept out of the product library
PROBE_LIB_PATH = os.path.join(PKG_DIR, "libll_probe.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def probe_sources():
    return sorted(glob.glob(os.path.join(CSRC, "probe", "*.cu"))) + [os.path.join(CSRC, "ll_api.cu")]


def _headers():
    inc = os.path.join(os.path.dirname(PKG_DIR), "include")
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(inc, "ll_api.h"), os.path.join(inc, "ll_probe.h")]


def _stale(lib, srcs):
    if not os.path.isfile(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in srcs + _headers() if os.path.exists(d))


def _compile(lib, srcs, log_name, verbose):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError(f"nvcc not found: cannot build {os.path.basename(lib)}")
    tmp = lib + ".tmp"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(PKG_DIR, log_name), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    os.replace(tmp, lib)
    if verbose:
        print(log)


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a into ``libll_b200.so`` (the product ABI, ``include/ll_api.h``) and
    ``csrc/probe/*.cu`` into ``libll_probe.so`` (``include/ll_probe.h``).  Returns the product library's path."""
    if force or _stale(LIB_PATH, sources()):
        _compile(LIB_PATH, sources(), "build.log", verbose)
    if force or _stale(PROBE_LIB_PATH, probe_sources() + sources()):      # the probe builds #include product sources
        _compile(PROBE_LIB_PATH, probe_sources(), "build_probe.log", verbose)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
