"""Build tuning variants of the library under build/variants/ (they travel to the GPU box with gpurun):
    python tools/build_variants.py name1:DEF1=V,DEF2=V name2:...
Each is selected at run time with P3D_LIB=<path> (see tools/run_variants.sh)."""
import importlib
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bn = importlib.import_module("part-based-3d-reconstruction_b200.build_native")
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)


def one(spec):
    name, _, defs = spec.partition(":")
    defines = [d for d in defs.split(",") if d]
    out = os.path.join(out_dir, f"lib_{name}.so")
    bn.build(force=True, defines=defines, out=out)
    return out


if __name__ == "__main__":
    for f in os.listdir(out_dir):
        os.remove(os.path.join(out_dir, f))
    with ThreadPoolExecutor(4) as ex:
        for path in ex.map(one, sys.argv[1:]):
            print("built", path)
