"""Build a variant of libso3d.so for A/B measurements:  python tests/tools/build_variant.py NAME [-DFLAG=..]...
-> build/variants/libso3d_NAME.so, selected at run time with SO3D_LIB_PATH=build/variants/libso3d_NAME.so."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diffusion_extensions_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
print(B.build(extra_flags=flags, out=os.path.join(out_dir, f"libso3d_{name}.so")))
