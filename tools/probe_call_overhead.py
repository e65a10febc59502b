"""Host-side cost of the slab carving calls (what bounds them once the slab kernel takes < 0.1 ms at N = 8): CPU time per
call without synchronisation, then a cProfile of 200 calls.  One GPU; the slab is 1/8 of a 1024^3 grid."""
import cProfile, importlib, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config")
syn = importlib.import_module(PKG + ".synthetic")
N = int(os.environ.get("PROBE_N", "1024"))
dev = torch.device("cuda")
lab = syn.monument_labels(N, dev)
front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]
ext = torch.from_numpy(lut[front]).to(dev); binm_d = torch.from_numpy((front > 0).astype(np.uint8)).to(dev)
span = (0, N // 8)
jobs90 = [([n], 90) for n in ("full_building", "chhatris", "plinth", "front_minarets", "small_minarets", "dome")]


def cpu_us(fn, n=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


def prof(fn, n=200, top=25):
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(top)
    print(s.getvalue()[:6000])


gc = lambda: vc.global_carve(binm_d, ext, 90, return_tensor=True, x_range=span)
print("global_carve slab: cpu %.1f us / call, incl. gpu %.1f us" % cpu_us(gc))
gfull = vc.global_carve(binm_d, ext, 90, return_tensor=True)
pm = vc._PackedMask(ext)
pc = lambda: vc.part_carve(gfull, pm, jobs90, x_range=span)
print("part_carve slab (replicated input): cpu %.1f us / call, incl. gpu %.1f us" % cpu_us(pc))
pcf = lambda: vc.part_carve(gfull[:N // 8], pm, jobs90) if False else None
slab_in = gfull[span[0]:span[1]].contiguous()
def sl():
    job = vc.PartCarveSlab(slab_in, pm, jobs90, N, span)
    job.begin()
    return job.finish()
print("PartCarveSlab new+begin+finish: cpu %.1f us / call, incl. gpu %.1f us" % cpu_us(sl))
job = vc.PartCarveSlab(slab_in, pm, jobs90, N, span)
def st():
    job.begin()
    return job.finish()
print("PartCarveSlab reused begin+finish: cpu %.1f us / call, incl. gpu %.1f us" % cpu_us(st))
if os.environ.get("PROBE_PROFILE", "1") == "1":
    prof(gc)
    prof(pc)
    prof(sl)
