"""Copy the reference's own `utils/*.py` verbatim into baseline/_ref/utils/ (git-ignored, shipped to the GPU box by
gpurun) so that bench.py's CPU arms time the UNMODIFIED reference functions.  Run where
/root/reference exists (the build container); __graft_entry__.build() calls it.  Nothing is copied into tracked paths."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("P3D_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=True) -> bool:
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        from harness import FILES
    finally:
        sys.path.pop(0)
    src = os.path.join(SRC, "utils")
    if not os.path.isdir(src):
        if verbose:
            print(f"install_ref: {src} not found; leaving {DST} as it is")
        return False
    os.makedirs(os.path.join(DST, "utils"), exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(DST, "utils", f))
    for extra in ("eval_helpers.py", "eval_helpers_intra.py"):                # star-imported nowhere, kept for completeness
        if os.path.exists(os.path.join(src, extra)):
            shutil.copyfile(os.path.join(src, extra), os.path.join(DST, "utils", extra))
    if verbose:
        print(f"install_ref: copied {len(FILES)} files to {DST}/utils")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
