"""Build a variant of libso3d.so for A/B measurements:  python tests/tools/build_variant.py NAME [-DFLAG=..]...
-> build/variants/libso3d_NAME.so, selected at run time with SO3D_LIB_PATH=build/variants/libso3d_NAME.so.

`--kernels-only` recompiles csrc/so3d_kernels.cu with the flags and links it against cached objects of the other two
translation units (build/obj/): a third of the build time when only the row-engine kernels vary."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diffusion_extensions_b200 import build as B

args = sys.argv[1:]
kernels_only = "--kernels-only" in args or any(a.startswith("--tu=") for a in args)
tu = next((a[5:] for a in args if a.startswith("--tu=")), "so3d_kernels.cu")   # the translation unit that gets the flags
args = [a for a in args if a != "--kernels-only" and not a.startswith("--tu=")]
name, flags = args[0], args[1:]
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libso3d_{name}.so")
if not kernels_only:
    print(B.build(extra_flags=flags, out=out))
else:
    nvcc = B.find_nvcc()
    base = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
    obj_dir = os.path.join(ROOT, "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    main = next(s for s in B.SRCS if os.path.basename(s) == tu)
    for src in B.SRCS:
        if src == main:
            continue
        o = os.path.join(obj_dir, os.path.basename(src) + ".o")
        if not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(d) for d in B.DEPS + [os.path.join(B.CSRC, "so3d_lanes.cuh")]):
            subprocess.check_call([nvcc, *base, "-c", "-o", o, src])
        objs.append(o)
    subprocess.check_call([nvcc, *base, "-shared", *flags, "-o", out, main, *objs])
    print(out)
