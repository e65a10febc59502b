"""Ragged (D % 32 != 0) against aligned bit-packed carve kernels at comparable sizes: whole calls with device-resident
inputs, CUDA events over 20 back-to-back calls.  python tools/probe_ragged.py [W ...]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
syn = importlib.import_module(PKG + ".synthetic")
dev = torch.device("cuda")
N = 512
lab = syn.monument_labels(N, dev)
front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]
jobs90 = [([n], 90) for n in ("full_building", "chhatris", "plinth", "front_minarets", "small_minarets", "dome")]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for W in [int(v) for v in sys.argv[1:]] or [512, 497, 481, 255]:
    lo = (N - W) // 2
    f = np.ascontiguousarray(front[:, lo:lo + W])              # (H = 512, W) crop around the centre
    ext = torch.from_numpy(lut[f]).to(dev); binm = torch.from_numpy((f > 0).astype(np.uint8)).to(dev)
    g = vc.global_carve(binm, ext, 90)
    plan = vc._fold_plan(W, N, W, dev)
    pm = vc._PackedMask(ext)
    # knock out 2 % of the voxels so that the clear pass of part_carve has work
    holes = torch.rand(g.shape[:3], device=dev) < 0.02
    ga = g.clone(); ga[holes] = 0
    vox = W * N * W
    t_g = timed(lambda: vc.global_carve(binm, ext, 90))
    t_p = timed(lambda: vc.part_carve(g, pm, jobs90))
    t_a = timed(lambda: vc.part_carve(ga, pm, jobs90))
    print(f"W=D={W} H={N} ({vox * 3 / 1e6:.0f} MB) bit path: {plan is not None and plan[1] is not None}; "
          f"global_carve {t_g:.4f} ms = {vox * 3 / t_g / 1e6:.0f} GB/s; part_carve {t_p:.4f} ms = {vox * 6 / t_p / 1e6:.0f} GB/s; "
          f"part_carve (2 % holes) {t_a:.4f} ms = {vox * 6 / t_a / 1e6:.0f} GB/s")
