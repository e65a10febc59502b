"""Timing of the GPU marching cubes and of the whole meshify_colored_voxel_grid call on the carved Bibi@256 grid and on the
synthetic 512^3 monument (occupancy only)."""
import contextlib, importlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
mu = importlib.import_module(PKG + ".utils.mask_utils"); syn = importlib.import_module(PKG + ".synthetic")
vu = importlib.import_module(PKG + ".utils.voxel_utils")
data = os.path.join(ROOT, "tests", "golden", "data")
sem, sem_ext, binary = mu.load_and_prepare_masks(data, "Bibi", "front", 256, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
sym = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
with contextlib.redirect_stdout(io.StringIO()):
    grid = vc.partwise_carve(vc.global_carve(binary, sem_ext, 90, return_tensor=True), sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym,
                             {"main_door": 20, "windows": 10})


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


mask = grid.any(dim=-1).to(torch.uint8)
ms, (v, f, n) = timed(lambda: vu.marching_cubes_binary(mask))
print(f"Bibi@256 {tuple(mask.shape)}: marching cubes {ms:.3f} ms -> {v.shape[0]} vertices, {f.shape[0]} faces")
t0 = time.perf_counter(); out = vu.meshify_colored_voxel_grid(grid); t1 = time.perf_counter()
print(f"  meshify_colored_voxel_grid whole call (incl. sklearn nearest neighbour on the host): {t1 - t0:.2f} s")
t0 = time.perf_counter(); out = vu.meshify_colored_voxel_grid(grid, stride=2); t1 = time.perf_counter()
print(f"  stride 2: {t1 - t0:.2f} s, {out[0].shape[0]} vertices")
lab = syn.monument_labels(512, torch.device("cuda"))
m5 = (lab > 0).to(torch.uint8)
ms, (v, f, n) = timed(lambda: vu.marching_cubes_binary(m5), reps=3)
print(f"synthetic 512^3: marching cubes {ms:.3f} ms -> {v.shape[0]} vertices, {f.shape[0]} faces ({m5.numel() / ms / 1e6:.1f} Gvoxel/s)")
