"""Scratch: depth-buffer evaluator timing on the full Taj grid (python tools/probe_depth.py)."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
eh = importlib.import_module(PKG + ".utils.eval_helpers_intra"); mu = importlib.import_module(PKG + ".utils.mask_utils")
vu = importlib.import_module(PKG + ".utils.voxel_utils"); cfg = importlib.import_module(PKG + ".utils.config")
data = os.path.join("tests", "golden", "data")
grid = np.load(os.path.join(data, "results", "1.Orthographic_Voxel_Carving", "Taj_voxel_grid.npz"))["voxel_grid"]
front = mu.load_mask(data, "Taj", "front", int(max(grid.shape)))
H, W = front.shape[:2]
g = torch.from_numpy(grid).cuda()
for dt in (np.float32, np.float64):
    cam = eh.load_camera_json(os.path.join(data, "results", "2.Perspective_Camera_Estimation", "Taj_camera_params_final.json"), "front")
    cam = {k: (v.astype(dt) if hasattr(v, "astype") else v) for k, v in cam.items()}
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        z = eh.compute_global_depth_buffer(g, cam, H, W, return_tensor=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        pts = vu.device_points_by_parts(g, cfg.PART_COLORS, ["dome"])[0]
        torch.cuda.synchronize(); t2 = time.perf_counter()
        m = eh.project_part_visible(pts, cam, z, H, W, return_tensor=True)
        torch.cuda.synchronize(); t3 = time.perf_counter()
    print(np.dtype(dt).name, "depth buffer %.2f ms, points %.2f ms, part_visible %.2f ms (%d pts)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, pts.shape[0]))
