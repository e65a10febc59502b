"""Scratch stage timing of the carving pipeline: python tools/probe_carve.py [N | bibi256]"""
import contextlib, importlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
mu = importlib.import_module(PKG + ".utils.mask_utils"); syn = importlib.import_module(PKG + ".synthetic")
arg = sys.argv[1] if len(sys.argv) > 1 else "bibi256"
if arg.startswith("bibi"):
    md = int(arg[4:])
    sem, ext, binm = mu.load_and_prepare_masks(os.path.join("tests", "golden", "data"), "Bibi", "front", md, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
else:
    N = int(arg)
    lab = syn.monument_labels(N, "cuda")
    front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
    lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]
    sem = lut[front]; ext = sem.copy()
    for q in cfg.INTERIOR_PARTS: ext[np.all(sem == cfg.PART_COLORS_NP[q], axis=-1)] = cfg.PART_COLORS_NP["full_building"]
    binm = (front > 0).astype(np.uint8); del lab
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
sym = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
extd = {"main_door": 20, "windows": 10}
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{label:28s} {1e3*(time.perf_counter()-t0):9.2f} ms"); return r
for rep in range(2):
    print("--- rep", rep, "grid", (binm.shape[1], binm.shape[0], binm.shape[1]))
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        pass
    g = T("global_carve (tensor)", lambda: vc.global_carve(binm, ext, 90, return_tensor=True))
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        pass
    p = T("part_carve", lambda: vc.part_carve(g, ext, jobs))
    cur = p
    for part, ang in sym.items():
        with contextlib.redirect_stdout(out):
            cur = T(f"left_right {part} {ang}", lambda: vc.left_right_guided_carve(cur, ext, cfg.PART_COLORS_NP[part], ang)) if False else cur
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with contextlib.redirect_stdout(out):
            cur = vc.left_right_guided_carve(cur, ext, cfg.PART_COLORS_NP[part], ang)
        torch.cuda.synchronize(); print(f"{'left_right ' + part:28s} {1e3*(time.perf_counter()-t0):9.2f} ms")
    def extr():
        c = cur.clone()
        for part, depth in extd.items():
            m = np.all(sem == cfg.PART_COLORS_NP[part], axis=-1)
            for axis, d in ((2, "+"), (2, "-"), (0, "+"), (0, "-")):
                vc._extrude_inplace(c, m, axis, d, depth, cfg.PART_COLORS_NP[part])
        return c
    e = T("extrude x8", extr)
    def reor():
        W, H, D, _ = e.shape
        o = torch.empty((D, H, W, 3), dtype=torch.uint8, device=e.device)
        vc.check(vc.lib.p3d_reorient(vc.ptr(e), W, H, D, vc.ptr(o), vc.stream_ptr())); return o
    o = T("reorient", reor)
    r = T("recolor_backward", lambda: vc.recolor_backward_components(o, cfg.PART_COLORS_NP["front_minarets"], cfg.PART_COLORS_NP["back_minarets"], 2, 0))
    with contextlib.redirect_stdout(out):
        full = T("partwise_carve total", lambda: vc.partwise_carve(g, ext, sem, cfg.PART_COLORS_NP, jobs, sym, extd))
    print("components log lines:", len(out.getvalue().split("\n")))
