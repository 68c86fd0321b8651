"""Build libso3d.so (the sm_100a kernels + C ABI) in-tree with nvcc.

    python -m diffusion_extensions_b200.build

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the tree.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SRCS = [os.path.join(CSRC, f) for f in ("so3d_kernels.cu", "so3d_pairwise.cu", "so3d_denoiser.cu", "so3d_loop.cu")]
DEPS = SRCS + [os.path.join(CSRC, f) for f in ("so3d_math.cuh", "so3d_lanes.cuh", "so3d_tma.cuh", "so3d_common.cuh", "so3d_cdf_smem.cuh")] + [os.path.join(HERE, "..", "include", "so3d.h")]
LIB = os.path.join(HERE, "libso3d.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # cudart is linked statically (nvcc default): the library only needs the driver at load time and
    # shares the primary context with torch, so torch's stream handles (CUstream) are valid here.
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: build a variant of the library elsewhere (A/B runs via SO3D_LIB_PATH)."""
    if out is None and not force and up_to_date():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", out or LIB, *SRCS]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
