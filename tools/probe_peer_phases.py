"""Phase timing of the peer form of part_carve_sharded (1024^3, x-slab per rank) under torchrun: where the time beyond the
two kernels goes.  Timing only: the variants without barriers race on purpose."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
PKG = "part-based-3d-reconstruction_b200"
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
syn = importlib.import_module(PKG + ".synthetic"); vc = importlib.import_module(PKG + ".utils.voxel_carving_utils")
cfg = importlib.import_module(PKG + ".utils.config"); sw = importlib.import_module(PKG + ".utils.sweep")
N = int(os.environ.get("PROBE_N", "1024"))
lab = syn.monument_labels(N, dev)
front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]
ext = torch.from_numpy(lut[front]).to(dev); binm = (front > 0).astype(np.uint8)
jobs90 = [([n], 90) for n in ("full_building", "chhatris", "plinth", "front_minarets", "small_minarets", "dome")]
gfull = vc.global_carve(binm, ext, 90, return_tensor=True)
span = sw.shard_range(N, world, rank)
slab_in = gfull[span[0]:span[1]].contiguous()
want = vc.part_carve(gfull, ext, jobs90, x_range=span)
del gfull
nbytes = vc.PartCarveSlab.workspace_bytes(N, N, N, len(jobs90))
buf, hdl, ptrs = sw.symmetric_workspace(nbytes, dev)
job = vc.PartCarveSlab(slab_in, ext, jobs90, N, span, workspace=buf)


def full():
    job.begin(); sw.peer_barrier(hdl); out = job.finish(peers=ptrs, n_ranks=world); sw.peer_barrier(hdl); return out


buf1, hdl1, ptrs1 = sw.symmetric_workspace(nbytes, dev, slot=1)
job1 = vc.PartCarveSlab(slab_in, ext, jobs90, N, span, workspace=buf1)
turn = [0]


def full_two_slots():                   # the shipped form: two workspaces in turn, one barrier per call
    j, h, p = (job, hdl, ptrs) if turn[0] % 2 == 0 else (job1, hdl1, ptrs1)
    turn[0] += 1
    j.begin(); sw.peer_barrier(h); return j.finish(peers=p, n_ranks=world)


variants = {
    "full_two_slots_one_barrier": full_two_slots,
    "full": full,
    "pass_a": lambda: job.begin(),
    "pass_a+b_local": lambda: (job.begin(), job.finish()),
    "pass_a+b_peers_no_barrier": lambda: (job.begin(), job.finish(peers=ptrs, n_ranks=world)),
    "barriers_only": lambda: (sw.peer_barrier(hdl), sw.peer_barrier(hdl)),
    "pass_a+barriers": lambda: (job.begin(), sw.peer_barrier(hdl), sw.peer_barrier(hdl)),
}
ok = bool(torch.equal(full(), want)) and all(bool(torch.equal(full_two_slots(), want)) for _ in range(3))
res = {}
for name, fn in variants.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(20):
        fn()
    eb.record(); torch.cuda.synchronize()
    t = torch.tensor([ea.elapsed_time(eb) / 20], dtype=torch.float64, device=dev)
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    res[name] = [round(float(tmin.item()), 4), round(float(t.item()), 4)]
okt = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"n_gpus": world, "correct": bool(okt.item()), "ms_min_max_over_ranks": res}))
dist.destroy_process_group()
