"""Build librgbdgan_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m rgbd_gan_b200.build [--force] [--verbose]

The built library is git-ignored (*.so) but travels with gpurun snapshots.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librgbdgan_b200.so")
SOURCES = ["api.cu", "consistency.cu", "deepvoxels.cu", "poses.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "sweep.cuh"),
           os.path.join(os.path.dirname(_HERE), "include", "rgbdgan_b200.h")]

# No -use_fast_math: IEEE division/sqrt and no flush-to-zero are part of the parity contract.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--compiler-options", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static",
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """`defines`/`out` build a tuning variant (e.g. defines=["RGBD_KPIX=2"], out="lib/variants/kpix2.so")"""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not stale():
        return LIB_PATH
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed (%d): %s" % (res.returncode, " ".join(cmd)))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
