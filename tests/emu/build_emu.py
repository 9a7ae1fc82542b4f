"""TEST INFRASTRUCTURE ONLY: compiles csrc/*.cu with g++ against cuda_emu.h into
tests/emu/_build/libtcsfm_emu.so so that kernel logic can be exercised on the
GPU-less build container.  Never loaded by the package, never timed."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "tightly-coupled-sfm_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libtcsfm_emu.so")
SOURCES = ["cabi.cu", "warp_kernels.cu", "ssim_kernels.cu", "pair_kernels.cu", "pair_fast_kernels.cu", "frame_kernels.cu", "photo_kernels.cu", "smooth_kernels.cu", "pft_kernels.cu"]


def digest():
    h = hashlib.sha256()
    for d, names in ((CSRC, sorted(os.listdir(CSRC))), (HERE, ["cuda_emu.h", "build_emu.py"]),
                     (os.path.join(ROOT, "include"), ["tcsfm.h"])):
        for n in names:
            with open(os.path.join(d, n), "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build(force=False, sanitize=None):
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = LIB if not sanitize else LIB.replace(".so", "_%s.so" % sanitize)
    stamp = lib + ".stamp"
    dg = digest()
    if not force and os.path.isfile(lib) and os.path.isfile(stamp) and open(stamp).read() == dg:
        return lib
    objs = []
    for i, src in enumerate(SOURCES):
        obj = os.path.join(OUT_DIR, src.replace(".cu", (".%s.o" % sanitize) if sanitize else ".o"))
        cmd = ["g++", "-std=c++20", "-O2", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-mfma",
               "-DTCSFM_HOST_EMU", "-I", HERE, "-I", CSRC, "-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj,
               "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable"]
        if i == 0:
            cmd.insert(1, "-DTCSFM_EMU_DEFINE_GLOBALS")
        if sanitize:
            cmd.insert(1, "-fsanitize=%s" % sanitize)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    link = ["g++", "-shared", "-pthread", "-o", lib] + objs
    if sanitize:
        link.insert(1, "-fsanitize=%s" % sanitize)
    subprocess.run(link, check=True)
    with open(stamp, "w") as f:
        f.write(dg)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
