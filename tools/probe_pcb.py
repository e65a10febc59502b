"""Kernel-only timing of the bit-level part_carve (pack_group_bits + part_copy_bits + part_clear) at N^3, graph-replayed
(tuning probe; env knobs P3D_PCB_WAVES / P3D_PCB_SMEM)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
syn = importlib.import_module(PKG + ".synthetic")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lab = syn.monument_labels(N, "cuda"); front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy(); del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]; ext = lut[front]; binm = (front > 0).astype(np.uint8)
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
g = vc.global_carve(binm, ext, 90, return_tensor=True)
p = vc.part_carve(g, ext, jobs)
launch = vc._LAST_PART_CARVE_LAUNCH
for _ in range(3): launch()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, stream=side):
    for _ in range(20): launch()
best = 1e9
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print(f"N={N} lib={os.path.basename(os.environ.get('P3D_LIB','default'))} waves={os.environ.get('P3D_PCB_WAVES','-')} smem={os.environ.get('P3D_PCB_SMEM','-')} part_carve_ms={best:.4f} GB/s={6*N**3/best/1e6:.0f}")
