"""Where the time of one left_right_guided_carve call goes (Bibi@256, front_minarets at 5 degrees): device time per phase."""
import contextlib, importlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
mu = importlib.import_module(PKG + ".utils.mask_utils")
data = os.path.join(ROOT, "tests", "golden", "data")
sem, sem_ext, binary = mu.load_and_prepare_masks(data, "Bibi", "front", 256, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
g = vc.global_carve(binary, sem_ext, 90, return_tensor=True)
pm = vc._PackedMask(sem_ext)
grid = vc.part_carve(g, pm, jobs)
col = cfg.PART_COLORS_NP["front_minarets"]
dev = grid.device


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = [ev()]
    carved = grid.clone(); m2 = pm.device_match(col, dev); m3 = vc._colour_mask(grid, col); e.append(ev())
    W, H, D, _ = grid.shape
    labels = torch.empty((W, H, D), dtype=torch.int32, device=dev); ncomp = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = int(vc.lib.p3d_label6_workspace_bytes(m3.numel())); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    vc.check(vc.lib.p3d_label6(vc.ptr(m3), W, H, D, vc.ptr(labels), vc.ptr(ncomp), vc.ptr(ws), wsb, vc.stream_ptr())); e.append(ev())
    bbox = torch.empty((1024, 6), dtype=torch.int32, device=dev); sums = torch.empty((1024, 4), dtype=torch.int64, device=dev)
    vc.check(vc.lib.p3d_component_stats(vc.ptr(labels), W, H, D, 1024, vc.ptr(bbox), vc.ptr(sums), vc.stream_ptr())); e.append(ev())
    n = int(ncomp.cpu().item()); t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = vc.left_right_guided_carve(grid, pm, col, 5)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    names = ["clone+masks", "label6", "stats"]
    print(f"rep {rep}: components {n}; " + "; ".join(f"{nm} {e[i].elapsed_time(e[i + 1]):.3f} ms" for i, nm in enumerate(names)) +
          f"; host until n known {1e3 * (t1 - t0):.3f} ms; whole LR call {1e3 * (t2 - t1):.3f} ms")
