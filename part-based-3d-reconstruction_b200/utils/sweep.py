"""Multi-GPU candidate sweep: candidates are sharded over ranks (one process per GPU), every rank
scores its contiguous block with no data-path collective, and ONE 16-byte-per-rank all-gather of
(score, global index) picks the winner.  Selection rule = the reference's strict `>` loop
(utils/camera_estimation.py:646): the FIRST candidate with the greatest score wins.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nv
from ._native import check, lib, ptr, stream_ptr


def shard_range(K: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank` when K candidates are split over `world` ranks."""
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def select_best(pairs):
    """pairs: iterable of (score, global_index); index < 0 marks an empty shard.  Greatest score, ties ->
    lowest index."""
    best_s, best_i = -np.inf, -1
    for s, i in pairs:
        i = int(i)
        if i < 0:
            continue
        if s > best_s or (s == best_s and i < best_i):
            best_s, best_i = float(s), i
    return best_s, best_i


class BestReducer:
    """Device-side, sync-free reduction of the per-rank best candidate (NCCL all-gather of 16 B per rank
    followed by a one-warp select kernel)."""

    launches_per_reduce = 2

    def __init__(self, device, world: int, group=None):
        self.device, self.world, self.group = device, world, group
        self.pair = torch.zeros(2, dtype=torch.int64, device=device)
        self.gathered = torch.zeros((world, 2), dtype=torch.int64, device=device)
        self.out = torch.zeros(2, dtype=torch.int64, device=device)

    def reduce(self, scores: torch.Tensor, best: torch.Tensor, offset: int):
        """scores (K) f64 + best (2) i64 of the local block whose first candidate has global index `offset`.
        Returns a LazyBest; reading it synchronises."""
        check(lib.p3d_best_pack(ptr(scores), ptr(best), int(offset), ptr(self.pair), stream_ptr()), "p3d_best_pack")
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered.view(-1), self.pair, group=self.group)
            src, n = self.gathered, self.world
        else:
            src, n = self.pair, 1
        check(lib.p3d_best_select(ptr(src), n, ptr(self.out), stream_ptr()), "p3d_best_select")
        nv.launch_count += 2
        return LazyBest(self.out)


class LazyBest:
    def __init__(self, t):
        self.t = t

    def __getitem__(self, i):
        v = self.t.cpu().numpy()
        return (float(v[:1].view(np.float64)[0]), int(v[1]))[i]


def all_gather_best(score: float, index: int, group=None, device=None):
    """Backend-agnostic version (NCCL on GPUs, gloo on CPUs): returns the global (score, index)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(score), int(index)
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    pair = torch.tensor([np.float64(score).view(np.int64), index], dtype=torch.int64, device=device)
    out = torch.empty((world, 2), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out.view(-1), pair, group=group)
    host = out.cpu().numpy()
    return select_best(zip(host[:, 0].copy().view(np.float64), host[:, 1]))


def score_candidates_sharded(scorer, candidates, group=None, gather_scores=False):
    """Score this rank's shard of `candidates` (K,9) with `scorer.score` and agree on the best.

    Returns (best_score, best_global_index, local_scores, (lo, hi)) -- or, with gather_scores=True, the full
    (K,) score vector in place of local_scores (one extra all-gather of 8*K bytes)."""
    cand = np.asarray(candidates).reshape(-1, 9)
    K = cand.shape[0]
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    lo, hi = shard_range(K, world, rank)
    if hi > lo:
        scores, _, best = scorer.score(cand[lo:hi])
        local = (float(scores[best]), lo + int(best))
    else:
        scores, local = np.zeros(0), (-np.inf, -1)
    best_score, best_index = all_gather_best(local[0], local[1], group)
    if gather_scores and world > 1:
        backend_dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        sizes = [shard_range(K, world, r) for r in range(world)]
        width = max(h - l for l, h in sizes)
        buf = torch.zeros(width, dtype=torch.float64, device=backend_dev)
        buf[:hi - lo] = torch.from_numpy(np.asarray(scores, dtype=np.float64)).to(backend_dev)
        out = torch.empty((world, width), dtype=torch.float64, device=backend_dev)
        dist.all_gather_into_tensor(out.view(-1), buf, group=group)
        out = out.cpu().numpy()
        scores = np.concatenate([out[r, :h - l] for r, (l, h) in enumerate(sizes)])
    return best_score, best_index, scores, (lo, hi)


def score_deformations_sharded(viewer, part, deforms, stride=1, group=None):
    """Notebook 3's deformation sweep across GPUs: the (D,4) deformation rows are split into contiguous
    blocks, every rank scores its block with `viewer.score(part, rows, stride)` (no data-path collective) and one
    16-byte all-gather picks the first deformation with the greatest IoU -- the strict `>` loop of the reference's
    auto-align (deformation_estimation.py:195-197).  Returns (best_iou, best_global_index, local_ious, (lo, hi))."""
    rows = np.ascontiguousarray(deforms, dtype=np.float64).reshape(-1, 4)
    D = rows.shape[0]
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    lo, hi = shard_range(D, world, rank)
    if hi > lo:
        ious = np.asarray(viewer.score(part, rows[lo:hi], stride=stride)[0], dtype=np.float64)
        k = int(np.argmax(ious))                                  # first index of the maximum
        local = (float(ious[k]), lo + k)
    else:
        ious, local = np.zeros(0), (-np.inf, -1)
    best_iou, best_index = all_gather_best(local[0], local[1], group)
    return best_iou, best_index, ious, (lo, hi)


def carve_sharded(carve_slab, W: int, group=None, gather: bool = False):
    """Carve a (W,H,D,3) grid in x-slabs, one contiguous slab per rank (global_carve's output at [x,y,z] depends only
    on the 2-D masks, so there is no exchange).  `carve_slab(x0, x1)` returns this rank's (x1-x0,H,D,3) uint8 tensor,
    e.g. `lambda a, b: global_carve(binary, sem_ext, 90, return_tensor=True, x_range=(a, b))`.

    Returns (slab, (x0, x1)), or with gather=True the full grid on every rank (one all-gather of equal-size padded
    slabs; NCCL on GPUs, gloo on CPUs)."""
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    x0, x1 = shard_range(W, world, rank)
    slab = carve_slab(x0, x1)
    if not gather or world == 1:
        return slab, (x0, x1)
    spans = [shard_range(W, world, r) for r in range(world)]
    width = max(b - a for a, b in spans)
    padded = slab.new_zeros((width,) + tuple(slab.shape[1:]))
    padded[:x1 - x0] = slab
    out = slab.new_empty((world * width,) + tuple(slab.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    full = torch.cat([out[r * width:r * width + (b - a)] for r, (a, b) in enumerate(spans)], dim=0)
    return full, (0, W)


_SYMM = {}                # (group name, bytes, device, slot) -> (symmetric buffer, rendezvous handle, pointer table)
_PEER_TURN = {}           # (group name, bytes, device) -> number of peer-form calls so far (all ranks call in lockstep)


def symmetric_workspace(nbytes: int, device, group=None, slot: int = 0):
    """A uint8 workspace of >= nbytes in NVLink peer-mapped (torch symmetric) memory, identical on every rank, plus the
    device table of every rank's base pointer.  Collective on first use per (group, size, slot); cached afterwards."""
    import torch.distributed._symmetric_memory as symm
    group = group if group is not None else dist.group.WORLD
    key = (group.group_name, int(nbytes), str(device), int(slot))
    hit = _SYMM.get(key)
    if hit is None:
        buf = symm.empty(int(nbytes), dtype=torch.uint8, device=device)
        hdl = symm.rendezvous(buf, group.group_name)
        ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        hit = _SYMM[key] = (buf, hdl, ptrs)
    return hit


def peer_barrier(hdl, group=None):
    """Stream-ordered barrier over the ranks of a symmetric-memory rendezvous: one tiny kernel that raises a flag in every
    peer's signal pad over NVLink and waits for theirs (no NCCL launch, no payload).  Falls back to a one-element
    all-reduce when this torch build's handle has no barrier."""
    if hasattr(hdl, "barrier"):
        hdl.barrier(channel=0)
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        dist.all_reduce(torch.zeros(1, dtype=torch.int32, device=dev), group=group)


def part_carve_sharded(grid_slab, semantic_mask, group_jobs, W: int, group=None, slab_cls=None, exchange="alltoall"):
    """part_carve of a grid whose x rows are sharded over the ranks (rank r holds rows shard_range(W, world, r); W must
    divide evenly): pass A per slab, ONE exchange of occupancy bits, pass B per slab.  Returns (output slab, (x0, x1)).
    The exchange is real: the fold of voxel_carving_utils.py:139-160 reads occ[W - z, y, x], the x<->z transposed source.

    exchange = "alltoall" (default): pass B of the slab [x0, x1) reads only the z-bit range [x0 + c2, x1 + c2) of every
               source row, so each rank sends every other rank just that word range of its own rows -- W*H*D/(8 world)
               bytes received per rank instead of the whole bit array;
             = "peer": no exchange buffer at all -- the workspaces live in NVLink peer-mapped symmetric memory and pass B
               reads the other ranks' rows in place (p3d_part_carve_slab_pass_b_peers), ordered by ONE signal-pad barrier
               per call (after every rank's pass A: one flag per peer over NVLink, no NCCL launch; two workspaces used
               in turn make a second barrier after pass B unnecessary).  CUDA only;
             = "allgather": the whole bit array on every rank (W*H*D/8 bytes; the round-1 form).
    `slab_cls` replaces voxel_carving_utils.PartCarveSlab (same begin / occ / needed_words / finish interface) in the
    CPU tests of this plumbing."""
    if slab_cls is None:
        from . import voxel_carving_utils as vc
        slab_cls = vc.PartCarveSlab
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    if W % world:
        raise ValueError(f"part_carve_sharded: width {W} does not divide over {world} ranks")
    if exchange not in ("alltoall", "peer", "allgather"):
        raise ValueError(f"exchange={exchange!r}")
    x0, x1 = shard_range(W, world, rank)
    if world > 1 and exchange == "peer":
        H, D = int(grid_slab.shape[1]), int(grid_slab.shape[2])
        nbytes = slab_cls.workspace_bytes(W, H, D, min(len(list(group_jobs)), 32))
        # Two workspaces used in turn, ONE barrier per call: call k writes its occupancy rows into workspace k % 2 and
        # the barrier orders every rank's pass A(k) before any pass B(k).  A peer may read workspace k % 2 until its
        # pass B(k) ends, and that lies before the barrier of call k + 1 in its stream -- which this rank passes before
        # pass A(k + 2) touches the same workspace again.  (One workspace needs a second barrier after pass B.)
        gname = (group if group is not None else dist.group.WORLD).group_name
        ckey = (gname, int(nbytes), str(grid_slab.device))
        turn = _PEER_TURN.get(ckey, 0)
        _PEER_TURN[ckey] = turn + 1
        buf, hdl, ptrs = symmetric_workspace(nbytes, grid_slab.device, group, slot=turn & 1)
        job = slab_cls(grid_slab, semantic_mask, group_jobs, W, (x0, x1), workspace=buf).begin()
        peer_barrier(hdl, group)                             # every rank's pass A is done (stream-ordered)
        return job.finish(peers=ptrs, n_ranks=world), (x0, x1)
    job = slab_cls(grid_slab, semantic_mask, group_jobs, W, (x0, x1)).begin()
    if world > 1 and exchange == "allgather":
        mine = job.occ[x0:x1].clone()
        dist.all_gather_into_tensor(job.occ.view(-1), mine.view(-1), group=group)
    elif world > 1:
        spans = [shard_range(W, world, r) for r in range(world)]
        send = [job.occ[x0:x1, :, slice(*job.needed_words(a, b))].contiguous() for a, b in spans]
        lo, hi = job.needed_words(x0, x1)
        recv = [job.occ.new_empty((b - a, job.occ.shape[1], hi - lo)) for a, b in spans]
        # pairwise exchange as ONE batch of sends/receives (NCCL groups them into a single launch; gloo, used by the CPU
        # tests, has no list all-to-all)
        ops = []
        for r in range(world):
            if r != rank:
                peer = r if group is None else dist.get_global_rank(group, r)
                ops.append(dist.P2POp(dist.isend, send[r], peer, group))
                ops.append(dist.P2POp(dist.irecv, recv[r], peer, group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for r, (a, b) in enumerate(spans):
            if r != rank:
                job.occ[a:b, :, lo:hi] = recv[r]
    return job.finish(), (x0, x1)
