"""The sharded paths over real NCCL (needs >= 2 GPUs on the box; skipped otherwise): candidate sweep, deformation sweep
and slab-sharded global_carve / part_carve with all-gather, each against the single-GPU result of the same call."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pkg

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ce, de, sw, vc, cfg = (pkg("utils.camera_estimation"), pkg("utils.deformation_estimation"), pkg("utils.sweep"),
                           pkg("utils.voxel_carving_utils"), pkg("utils.config"))
    ag = np.load(os.path.join(GOLDEN, "aligner_golden.npz"))
    grid, image, base = ag["grid"], ag["image"], ag["free_saved"]
    parts = ["front_minarets", "back_minarets"]
    cand = ce.random_candidates(base, 37, np.random.default_rng(3))
    cand[30] = cand[4]                                                       # equal scores in different shards
    scorer = ce.CandidateScorer(grid, image, cfg.PART_COLORS, parts)
    best_s, best_i, scores, span = sw.score_candidates_sharded(scorer, cand, gather_scores=True)
    import contextlib, io
    labels = {k: v for k, v in cfg.PART_COLORS.items() if k != "background"}
    cam = ce.row_to_params(base)
    with contextlib.redirect_stdout(io.StringIO()):
        viewer = de.DeformViewer(grid, labels, image, cam, ["dome"])
    rng = np.random.default_rng(8)
    rows = np.column_stack([rng.uniform(0.8, 1.2, 21), rng.uniform(-20, 20, 21), rng.uniform(0.8, 1.2, 21), rng.uniform(-20, 20, 21)])
    d_iou, d_i, d_local, d_span = sw.score_deformations_sharded(viewer, "dome", rows)
    cg = np.load(os.path.join(GOLDEN, "carve_golden.npz"))
    binm, ext = cg["syn_rect40x64_bin"], cg["syn_rect40x64_ext"]
    full, _ = sw.carve_sharded(lambda a, b: vc.global_carve(binm, ext, 90, return_tensor=True, x_range=(a, b)), binm.shape[1], gather=True)
    from helpers import GROUP_JOBS
    # part_carve by output x-slab from the replicated global_carve grid (live-reference golden vector), no exchange
    pc, _ = sw.carve_sharded(lambda a, b: vc.part_carve(full, ext, GROUP_JOBS, x_range=(a, b)), binm.shape[1], gather=True)
    # ... and with the INPUT sharded as well: each rank passes only its rows, one all-gather of occupancy bits
    a, b = sw.shard_range(binm.shape[1], world, rank)
    ps, _ = sw.part_carve_sharded(full[a:b].contiguous(), ext, GROUP_JOBS, binm.shape[1])
    # every exchange form of the sharded-input part_carve on a 256-wide grid that is NOT 4-way symmetric (8 bit words per
    # row: the all-to-all moves a strict subset; "peer" reads the other rank's rows over NVLink), against part_carve of
    # the whole grid on this GPU
    rng2 = np.random.default_rng(5)
    W2, H2 = 256, 6
    sem2 = np.zeros((H2, W2, 3), np.uint8)
    sem2[:, :] = cfg.PART_COLORS["background"]
    sem2[:, 40:200] = cfg.PART_COLORS["full_building"]
    sem2[:, 90:120] = cfg.PART_COLORS["dome"]
    sem2[:2, 130:170] = cfg.PART_COLORS["plinth"]
    g2 = np.zeros((W2, H2, W2, 3), np.uint8)
    occ2 = rng2.random((W2, H2, W2)) < 0.5
    g2[occ2] = sem2.transpose(1, 0, 2)[:, :, None, :].repeat(W2, axis=2)[occ2]
    g2[np.all(g2 == np.asarray(cfg.PART_COLORS["background"], np.uint8), axis=-1)] = 0
    g2d = torch.from_numpy(g2).cuda()
    want2 = vc.part_carve(g2d, sem2, GROUP_JOBS)
    a2, b2 = sw.shard_range(W2, world, rank)
    # a second grid (other occupancy), carved in alternation with the first: the peer form uses two symmetric
    # workspaces in turn with one barrier per call, so a stale or prematurely overwritten workspace shows up as a
    # wrong slab on one of the five back-to-back calls
    g3d = torch.flip(g2d, dims=[2]).contiguous()
    want3 = vc.part_carve(g3d, sem2, GROUP_JOBS)
    modes_ok = {}
    for mode in ("alltoall", "allgather", "peer"):
        for rep in range(5):
            try:
                src, want = (g2d, want2) if rep % 2 == 0 else (g3d, want3)
                got2, _ = sw.part_carve_sharded(src[a2:b2].contiguous(), sem2, GROUP_JOBS, W2, exchange=mode)
                modes_ok[mode] = modes_ok.get(mode, True) and bool(torch.equal(got2, want[a2:b2]))
            except Exception as exc:                         # reported, asserted by the parent
                modes_ok[mode] = repr(exc)
                break
    import json
    with open(os.path.join(out_dir, f"modes{rank}.json"), "w") as fh:
        json.dump(modes_ok, fh)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), best_s=best_s, best_i=best_i, scores=scores, d_iou=d_iou, d_i=d_i,
             d_local=d_local, d_lo=d_span[0], carve=full.cpu().numpy(), partcarve=pc.cpu().numpy(),
             partcarve_slab=ps.cpu().numpy(), slab=np.array([a, b]))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_paths_over_nccl(tmp_path, carve_golden):
    import contextlib, io
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ce, de, cfg = pkg("utils.camera_estimation"), pkg("utils.deformation_estimation"), pkg("utils.config")
    ag = np.load(os.path.join(GOLDEN, "aligner_golden.npz"))
    grid, image, base = ag["grid"], ag["image"], ag["free_saved"]
    cand = ce.random_candidates(base, 37, np.random.default_rng(3))
    cand[30] = cand[4]
    want_scores, _, want_best = ce.CandidateScorer(grid, image, cfg.PART_COLORS, ["front_minarets", "back_minarets"]).score(cand)
    labels = {k: v for k, v in cfg.PART_COLORS.items() if k != "background"}
    with contextlib.redirect_stdout(io.StringIO()):
        viewer = de.DeformViewer(grid, labels, image, ce.row_to_params(base), ["dome"])
    rng = np.random.default_rng(8)
    rows = np.column_stack([rng.uniform(0.8, 1.2, 21), rng.uniform(-20, 20, 21), rng.uniform(0.8, 1.2, 21), rng.uniform(-20, 20, 21)])
    want_ious = viewer.score("dome", rows)[0]
    got_ious = np.zeros(21)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["scores"], want_scores) and int(z["best_i"]) == int(want_best) == int(np.argmax(want_scores))
        assert float(z["best_s"]) == want_scores[want_best]
        got_ious[int(z["d_lo"]):int(z["d_lo"]) + len(z["d_local"])] = z["d_local"]
        assert int(z["d_i"]) == int(np.argmax(want_ious)) and float(z["d_iou"]) == want_ious.max()
        assert np.array_equal(z["carve"], carve_golden["syn_rect40x64_global"])
        assert np.array_equal(z["partcarve"], carve_golden["syn_rect40x64_partcarve"])
        a, b = (int(v) for v in z["slab"])
        assert np.array_equal(z["partcarve_slab"], carve_golden["syn_rect40x64_partcarve"][a:b])
    assert np.array_equal(got_ious, want_ious)
    import json
    for r in range(world):
        modes = json.load(open(tmp_path / f"modes{r}.json"))
        assert modes == {"alltoall": True, "allgather": True, "peer": True}, modes
