"""In-tree build of the C-ABI CUDA library (libtcsfm_b200.so) for sm_100a.

    python -m tcsfm_b200.build            # or: __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so is written next to this file so that
it travels with a snapshot of the repository; it is git-ignored.
"""
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtcsfm_b200.so")
SOURCES = ["cabi.cu", "warp_kernels.cu", "ssim_kernels.cu", "pair_kernels.cu", "pair_fast_kernels.cu", "frame_kernels.cu", "photo_kernels.cu", "smooth_kernels.cu", "pft_kernels.cu"]
HEADERS = ["common.cuh", "tcsfm_math.cuh", "tile.cuh", os.path.join("..", "..", "include", "tcsfm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              # parity-critical arithmetic uses explicit *_rn intrinsics (never contracted);
              # the default --fmad=true only touches the gradient / reduction arithmetic
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def source_digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    stamp = LIB_PATH + ".stamp"
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == source_digest()


def build_variant(out_path, defines):
    """Tuning builds: same sources with extra -D macros, written to `out_path`."""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D%s" % d for d in defines] + ["-I", CSRC, "-o", out_path] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed")
    return proc.stderr


def build(force=False, verbose=False):
    if not force and is_current():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libtcsfm_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    with open(LIB_PATH + ".stamp", "w") as f:
        f.write(source_digest())
    with open(LIB_PATH + ".ptxas.log", "w") as f:
        f.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
