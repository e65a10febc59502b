"""Host-side cost of the carving API: Bibi@256 pipeline (NumPy in/out and device-chained) and part_carve at 512^3."""
import contextlib, importlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
mu = importlib.import_module(PKG + ".utils.mask_utils"); syn = importlib.import_module(PKG + ".synthetic")
data = os.path.join(ROOT, "tests", "golden", "data")
sem, sem_ext, binary = mu.load_and_prepare_masks(data, "Bibi", "front", 256, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
sym = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
ext_d = {"main_door": 20, "windows": 10}


def best(fn, n=5):
    b = None
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        b = dt if b is None else min(b, dt)
    return b * 1e3, r


with contextlib.redirect_stdout(io.StringIO()):
    t_np, _ = best(lambda: vc.partwise_carve(vc.global_carve(binary, sem_ext, 90), sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym, ext_d))
    t_dev, _ = best(lambda: vc.partwise_carve(vc.global_carve(binary, sem_ext, 90, return_tensor=True), sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym, ext_d))
    g = vc.global_carve(binary, sem_ext, 90, return_tensor=True)
    pm_e, pm_f = vc._PackedMask(sem_ext), vc._PackedMask(sem)
    t_g, _ = best(lambda: vc.global_carve(binary, sem_ext, 90, return_tensor=True))
    t_pc, pc = best(lambda: vc.part_carve(g, pm_e, jobs))
    cur = pc
    t_lr = []
    for part, angle in sym.items():
        t, cur2 = best(lambda: vc.left_right_guided_carve(cur, pm_e, cfg.PART_COLORS_NP[part], angle))
        t_lr.append((part, round(t, 3)))
        cur = cur2
print(f"Bibi@256 numpy in/out {t_np:.2f} ms; device chain {t_dev:.2f} ms; global_carve {t_g:.3f}; part_carve {t_pc:.3f}; LR {t_lr}")
dev = torch.device("cuda")
for N in (512,):
    lab = syn.monument_labels(N, dev)
    front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
    del lab
    lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]
    ext = torch.from_numpy(lut[front]).to(dev); binm = (front > 0).astype(np.uint8)
    out = vc.global_carve(binm, ext, 90, return_tensor=True)
    t_g, _ = best(lambda: vc.global_carve(binm, ext, 90, return_tensor=True))
    t_pc, _ = best(lambda: vc.part_carve(out, ext, jobs))
    t_pc_np, _ = best(lambda: vc.part_carve(out, lut[front], jobs))
    print(f"{N}^3: global_carve call {t_g:.3f} ms; part_carve call {t_pc:.3f} ms (device mask), {t_pc_np:.3f} ms (NumPy mask)")

# stage timing of the Bibi@256 device chain (synchronised between stages)
def stage(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {label}: {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
    return r
with contextlib.redirect_stdout(io.StringIO()):
    for rep in range(2):
        print("rep", rep, file=sys.stderr)
        pm_e, pm_f = vc._PackedMask(sem_ext), vc._PackedMask(sem)
        g = stage("global_carve", lambda: vc.global_carve(binary, sem_ext, 90, return_tensor=True))
        cur = stage("part_carve", lambda: vc.part_carve(g, pm_e, jobs))
        for part, angle in sym.items():
            cur = stage("lr " + part, lambda: vc.left_right_guided_carve(cur, pm_e, cfg.PART_COLORS_NP[part], angle))
        def extr():
            for part, depth in ext_d.items():
                mask = pm_f.device_match(cfg.PART_COLORS_NP[part], cur.device)
                for axis, direction in ((2, "+"), (2, "-"), (0, "+"), (0, "-")):
                    vc._extrude_inplace(cur, mask, axis, direction, depth, cfg.PART_COLORS_NP[part])
        stage("extrude x8", extr)
        W, H, D, _ = cur.shape
        oriented = torch.empty((D, H, W, 3), dtype=torch.uint8, device=cur.device)
        stage("reorient", lambda: vc.check(vc.lib.p3d_reorient(vc.ptr(cur), W, H, D, vc.ptr(oriented), vc.stream_ptr())))
        stage("recolour", lambda: vc.recolor_backward_components(oriented, cfg.PART_COLORS_NP["front_minarets"], new_color=cfg.PART_COLORS_NP["back_minarets"], k=2, sort_axis=0))

# portrait mask (Charminar@256: grid 177 x 256 x 177 -- no multiple of 32, table-driven kernels): whole calls
g5 = np.load(os.path.join(ROOT, "tests", "golden", "real5_golden.npz"))
for key in ("real_Charminar_256", "real_Itimad_256"):
    ext5, bin5 = torch.from_numpy(g5[key + "_ext"]).cuda(), torch.from_numpy(g5[key + "_bin"]).cuda()
    gg = vc.global_carve(bin5, ext5, 90)
    t_g5, _ = best(lambda: vc.global_carve(bin5, ext5, 90))
    t_p5, _ = best(lambda: vc.part_carve(gg, ext5, jobs))
    print(f"{key}: grid {tuple(gg.shape[:3])} ({gg.numel() / 1e6:.1f} MB): global_carve call {t_g5:.3f} ms, part_carve call {t_p5:.3f} ms")
