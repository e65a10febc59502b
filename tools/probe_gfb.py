"""Kernel-only timing of global_fold_bits_kernel<RGB> at N^3 (tuning probe; env knobs P3D_GFB_WAVES / P3D_GFB_SMEM)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
syn = importlib.import_module(PKG + ".synthetic"); nv = importlib.import_module(PKG + ".utils._native")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
lab = syn.monument_labels(N, dev); front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy(); del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]; ext = torch.from_numpy(lut[front]).to(dev); binm = (front > 0).astype(np.uint8)
M, off = vc._pass_transform((N, N, N), 90)
table, foldable = vc._fold_table(N, N, M, off, dev)
bits = vc._fold_bits(table, N, N, (N, N, M.tobytes(), off.tobytes(), str(dev)))
m_hw = torch.from_numpy(binm).to(dev)
kout = torch.empty((N, N, N, 3), dtype=torch.uint8, device=dev)
wpr = (N + 31) // 32 + 2
mbits = torch.empty((N, wpr), dtype=torch.int32, device=dev)
nv.check(nv.lib.p3d_pack_mask_bits(nv.ptr(m_hw), N, N, nv.ptr(mbits), wpr, nv.stream_ptr()))
def launch():
    nv.check(nv.lib.p3d_global_carve_fold_bits(N, N, N, 0, N, nv.ptr(bits[0]), bits[1], nv.ptr(mbits), wpr, nv.ptr(ext), 1, nv.ptr(kout), nv.stream_ptr()))
for _ in range(5): launch()
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(20): launch()
    k1.record(); torch.cuda.synchronize()
    best = min(best, k0.elapsed_time(k1) / 20)
print(f"N={N} waves={os.environ.get('P3D_GFB_WAVES','-')} smem={os.environ.get('P3D_GFB_SMEM','-')} kernel_ms={best:.4f} GB/s={3*N**3/best/1e6:.0f}")
