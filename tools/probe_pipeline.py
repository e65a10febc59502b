"""Host-side profile of the Bibi@256 carving pipeline (global_carve + partwise_carve, NumPy in / NumPy out)."""
import contextlib, cProfile, importlib, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
mu = importlib.import_module(PKG + ".utils.mask_utils")
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 256
name = sys.argv[2] if len(sys.argv) > 2 else "Bibi"
data = os.path.join(ROOT, "tests", "golden", "data")
sem, sem_ext, binary = mu.load_and_prepare_masks(data, name, "front", dim, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
sym = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
ext_d = {"main_door": 20, "windows": 10}
def run():
    g = vc.global_carve(binary, sem_ext, 90)
    return vc.partwise_carve(g, sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym, ext_d)
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); run(); t = time.perf_counter() - t0
print(f"{name}@{dim}: {1e3 * t:.1f} ms", file=sys.stderr)
pr = cProfile.Profile()
with contextlib.redirect_stdout(io.StringIO()):
    pr.enable(); run(); pr.disable()
st = pstats.Stats(pr, stream=sys.stderr); st.sort_stats("tottime").print_stats(22)
