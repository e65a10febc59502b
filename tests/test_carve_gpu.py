"""GPU parity tests of the carving path (stage 1): CUDA kernels through the reference-signature Python layer
against the oracle and the golden vectors produced by the live reference."""
import numpy as np
import pytest
import torch

from conftest import pkg
from helpers import EXTRUSION_DEPTHS, GROUP_JOBS, PART_SYMMETRY, sha, unpack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vc():
    assert torch.cuda.is_available()
    return pkg("utils.voxel_carving_utils")


def test_affine_golden_vectors(vc, carve_golden):
    """process_voxel_grid's resample against scipy outputs recorded from the live reference environment."""
    g = carve_golden
    for k in range(int(g["aff_n"])):
        shape = tuple(int(v) for v in g[f"aff{k}_shape"])
        vol = unpack(g[f"aff{k}_vol"], shape)
        ang = int(g[f"aff{k}_angle"])
        M, off = vc._pass_transform(shape, ang)
        dev = torch.device("cuda")
        vin = torch.from_numpy(vol).to(dev)
        out = torch.empty_like(vin)
        import ctypes
        nv = pkg("utils._native")
        nv.check(nv.lib.p3d_resample_carve(nv.ptr(vin), *shape, M.ctypes.data_as(ctypes.c_void_p),
                                           off.ctypes.data_as(ctypes.c_void_p), None, nv.ptr(out), nv.stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), unpack(g[f"aff{k}_out"], shape)), (k, shape, ang)


@pytest.mark.parametrize("shape", [(15, 4, 15), (16, 5, 16), (31, 3, 31), (63, 2, 63), (64, 3, 64), (43, 47, 44),
                                   (30, 9, 58), (100, 2, 100), (127, 2, 127), (128, 2, 128)])
@pytest.mark.parametrize("interval", [90, 45, 60, 5])
def test_process_voxel_grid_vs_oracle(vc, oracle, shape, interval):
    rng = np.random.default_rng(shape[0] * 1000 + interval)
    vol = (rng.random(shape) < 0.6).astype(np.uint8)
    mask = rng.random((shape[1], shape[0])) < 0.8              # (H,W)
    got = vc.process_voxel_grid(vol, mask, interval)
    want = oracle.process_voxel_grid(vol, mask, interval)
    assert got.dtype == np.uint8 and np.array_equal(got, want)


def test_fold_fast_path_is_taken_for_cubic_90(vc):
    dev = torch.device("cuda")
    for W in (15, 16, 64, 255, 256):
        M, off = vc._pass_transform((W, 3, W), 90)
        _, foldable = vc._fold_table(W, W, M, off, dev)
        assert foldable, W
    M, off = vc._pass_transform((43, 3, 44), 90)
    assert not vc._fold_table(43, 44, M, off, dev)[1]          # half-integer offsets: genuine blend
    # the bit-packed form of the fold is available for ragged widths too (portrait masks: Charminar is 88 / 177 / 246
    # wide), not only for multiples of 32; below 16 voxels the table-driven kernels remain
    for W in (16, 17, 31, 48, 63, 88, 100, 177, 246, 255, 256):
        plan = vc._fold_plan(W, 5, W, dev)
        assert plan is not None and plan[1] is not None and plan[1][2] is not None, W
        assert tuple(plan[1][0].shape) == (W, (W + 31) // 32)
    assert vc._fold_plan(15, 5, 15, dev)[1] is None
    M, off = vc._pass_transform((32, 3, 32), 5)
    assert not vc._fold_table(32, 32, M, off, dev)[1]


def test_label6_vs_scipy(vc, oracle, carve_golden):
    import scipy.ndimage
    g = carve_golden
    dev = torch.device("cuda")
    cases = [unpack(g[f"lab{i}_mask"], tuple(int(v) for v in g[f"lab{i}_shape"])) for i in range(int(g["lab_n"]))]
    rng = np.random.default_rng(4)
    cases += [(rng.random((33, 17, 70)) < p).astype(np.uint8) for p in (0.05, 0.35, 0.5, 0.9)]
    cases += [np.zeros((4, 5, 6), np.uint8), np.ones((7, 3, 9), np.uint8)]
    for m in cases:
        labels, n, bbox, sums = vc._label_components(torch.from_numpy(np.ascontiguousarray(m, dtype=np.uint8)).to(dev))
        want, wn = scipy.ndimage.label(m)
        assert n == wn and np.array_equal(labels.cpu().numpy(), want)
        for i in range(1, n + 1):
            idx = np.argwhere(want == i)
            assert np.array_equal(bbox[i - 1], np.concatenate([idx.min(0), idx.max(0)]))
            assert sums[i - 1, 0] == len(idx) and np.array_equal(sums[i - 1, 1:], idx.sum(0))


@pytest.mark.parametrize("case", ["Bibi_64", "Taj_96", "Akbar_128", "Bibi_256"])
def test_real_mask_carving_matches_reference(vc, carve_golden, case, capsys):
    g = carve_golden
    key = "real_" + case
    sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
    cfg = pkg("utils.config")
    grid = vc.global_carve(binm, ext, 90)
    assert grid.dtype == np.uint8 and grid.shape == (binm.shape[1], binm.shape[0], binm.shape[1], 3)
    assert sha(grid) == str(g[key + "_global_sha"])
    capsys.readouterr()
    out = vc.partwise_carve(grid, ext, sem, cfg.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
    log = capsys.readouterr().out
    assert sha(out) == str(g[key + "_partwise_sha"])
    assert np.count_nonzero(out.any(-1)) == int(g[key + "_partwise_occ"])
    assert log.strip("\n") == str(g[key + "_log"]).strip("\n")
    if key + "_partwise" in g.files:
        assert np.array_equal(out, g[key + "_partwise"])


@pytest.mark.parametrize("case", ["Charminar_128", "Charminar_256", "Itimad_256"])
def test_remaining_monuments_match_reference(vc, case, capsys):
    """The two monuments beside Bibi / Taj / Akbar (tests/golden/real5_golden.npz, live reference): Charminar is a PORTRAIT
    mask (grid width 88 / 177: neither a multiple of 32 nor of 4 -- every vectorised or bit-packed kernel sees its ragged
    tail), Itimad a landscape one; global_carve, part_carve and the whole partwise_carve chain with the printed log."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "real5_golden.npz"))
    key = "real_" + case
    sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
    cfg = pkg("utils.config")
    grid = vc.global_carve(binm, ext, 90)
    assert grid.shape == (binm.shape[1], binm.shape[0], binm.shape[1], 3)
    assert sha(grid) == str(g[key + "_global_sha"]) and np.count_nonzero(grid.any(-1)) == int(g[key + "_global_occ"])
    assert sha(vc.part_carve(grid, ext, GROUP_JOBS)) == str(g[key + "_partcarve_sha"])
    for x_range in ((0, 5), (grid.shape[0] // 2 - 3, grid.shape[0] - 1)):            # x-slabs tile the same bytes
        slab = vc.global_carve(binm, ext, 90, x_range=x_range)
        assert np.array_equal(slab, grid[x_range[0]:x_range[1]])
    capsys.readouterr()
    out = vc.partwise_carve(grid, ext, sem, cfg.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
    log = capsys.readouterr().out
    assert sha(out) == str(g[key + "_partwise_sha"]) and np.count_nonzero(out.any(-1)) == int(g[key + "_partwise_occ"])
    assert log.strip("\n") == str(g[key + "_log"]).strip("\n")
    if key + "_partwise" in g.files:
        assert np.array_equal(out, g[key + "_partwise"]) and np.array_equal(grid, g[key + "_global"])


def test_device_resident_masks_give_the_same_bytes(vc, carve_golden):
    """Masks handed over as CUDA tensors (no host round trip inside the call) against the NumPy-mask path, on the bit
    fast path (64 wide), a ragged width and the general-angle path."""
    g = carve_golden
    for key, angle in (("syn_sq64", 90), ("syn_rect40x64", 90), ("syn_rect50x31", 90), ("syn_rect40x64", 45)):
        ext, binm = g[key + "_ext"], g[key + "_bin"]
        want = vc.global_carve(binm, ext, angle)
        got = vc.global_carve(torch.from_numpy(binm).cuda(), torch.from_numpy(ext).cuda(), angle)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), want), (key, angle)
        got255 = vc.global_carve((torch.from_numpy(binm) * 255).to(torch.uint8).cuda(), torch.from_numpy(ext).cuda(), angle)
        assert np.array_equal(got255.cpu().numpy(), want), (key, angle)      # any non-zero value is foreground (:279-283)
        jobs = GROUP_JOBS
        assert np.array_equal(vc.part_carve(got, torch.from_numpy(ext).cuda(), jobs).cpu().numpy(), vc.part_carve(want, ext, jobs))


def test_lenient_inputs_like_the_reference(vc, carve_golden):
    """What the reference accepts beyond uint8 RGB: integer grids / masks of another dtype (values 0..255), an RGBA
    colour mask (it reads channels 0..2 only, :134-135).  Values that do not fit a byte are refused, not wrapped."""
    g = carve_golden
    ext, binm = g["syn_rect40x64_ext"], g["syn_rect40x64_bin"]
    want = vc.global_carve(binm, ext, 90)
    rgba = np.concatenate([ext, np.full(ext.shape[:2] + (1,), 255, np.uint8)], axis=-1)
    assert np.array_equal(vc.global_carve(binm.astype(np.int64), rgba, 90), want)
    carved = (want.any(-1)).astype(np.int32)
    assert np.array_equal(vc.apply_colored_mask_to_voxel_grid(carved, rgba), vc.apply_colored_mask_to_voxel_grid(carved.astype(np.uint8), ext))
    occ = want.any(-1).astype(np.int16)
    assert np.array_equal(vc.process_voxel_grid(occ, binm, 90), vc.process_voxel_grid(occ.astype(np.uint8), binm, 90))
    with pytest.raises(TypeError):
        vc.process_voxel_grid(occ * 300, binm, 90)
    with pytest.raises(TypeError):
        vc.process_voxel_grid(occ.astype(np.float32), binm, 90)


def test_synthetic_quirk_cases_match_reference(vc, carve_golden):
    """Square image (_mask_to_wh transposes), foreground in the last column, widths with the odd FP offsets."""
    g = carve_golden
    cfg = pkg("utils.config")
    for tag in g["syn_cases"]:
        key = f"syn_{tag}"
        sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
        grid = vc.global_carve(binm, ext, 90)
        assert np.array_equal(grid, g[key + "_global"]), tag
        assert np.array_equal(vc.part_carve(grid, ext, GROUP_JOBS), g[key + "_partcarve"]), tag
        out = vc.partwise_carve(grid, ext, sem, cfg.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
        assert np.array_equal(out, g[key + "_partwise"]), tag
        out2 = vc.partwise_carve(grid, ext, sem, cfg.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS,
                                 recolor_back_minarets=False)
        assert sha(out2) == str(g[key + "_partwise_norecolor_sha"]), tag


def test_general_angle_paths_vs_oracle(vc, oracle, carve_golden):
    """global_carve / part_carve away from the 90-degree fast path (angle_interval 45 and mixed group angles)."""
    g = carve_golden
    key = "syn_rect40x64"
    sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
    assert np.array_equal(vc.global_carve(binm, ext, 45), oracle.global_carve(binm, ext, 45))
    grid = g[key + "_global"]
    jobs = [(["full_building"], 45), (["chhatris", "dome"], 90), (["plinth"], 30)]
    assert np.array_equal(vc.part_carve(grid, ext, jobs), oracle.part_carve(grid, ext, jobs))


def test_building_blocks_vs_oracle(vc, oracle, carve_golden):
    g = carve_golden
    cfg = pkg("utils.config")
    key = "syn_sq64"
    sem, ext = g[key + "_sem"], g[key + "_ext"]
    grid = g[key + "_partcarve"]
    # inputs are not mutated, outputs are fresh arrays
    before = grid.copy()
    for part, depth in EXTRUSION_DEPTHS.items():
        m = np.all(sem == cfg.PART_COLORS_NP[part], axis=-1)
        for axis, direction in ((2, "+"), (2, "-"), (0, "+"), (0, "-")):
            got = vc.extrude_from_surface(grid, m, axis, direction, depth, cfg.PART_COLORS_NP[part])
            assert np.array_equal(got, oracle.extrude_from_surface(grid, m, axis, direction, depth, cfg.PART_COLORS_NP[part]))
        assert np.array_equal(vc.extrude_from_surface(grid, m, 2, "+", 3, None),
                              oracle.extrude_from_surface(grid, m, 2, "+", 3, None))
    assert np.array_equal(grid, before)
    for colour in ("front_minarets", "plinth", "full_building"):
        for k, axis in ((2, 0), (1, 2), (4, 1)):
            got = vc.recolor_backward_components(grid, cfg.PART_COLORS_NP[colour], cfg.PART_COLORS_NP["back_minarets"], k, axis)
            want = oracle.recolor_backward_components(grid, cfg.PART_COLORS_NP[colour], cfg.PART_COLORS_NP["back_minarets"], k, axis)
            assert np.array_equal(got, want), (colour, k, axis)
    for part, angle in PART_SYMMETRY.items():
        got = vc.left_right_guided_carve(grid, ext, cfg.PART_COLORS_NP[part], angle)
        assert np.array_equal(got, oracle.left_right_guided_carve(grid, ext, cfg.PART_COLORS_NP[part], angle)), part
    occ = (grid.any(-1)).astype(np.uint8)
    m2 = np.all(ext != cfg.PART_COLORS_NP["background"], axis=-1) if False else (~np.all(ext == cfg.PART_COLORS_NP["background"], axis=-1))
    assert np.array_equal(vc.carve_voxel_grid_with_masks(occ, m2), oracle.carve_with_mask(occ, m2))
    got = vc.carve_voxel_grid_with_masks(grid, m2)
    assert np.array_equal(got, np.where(oracle.mask_to_wh(m2, 64, 64)[:, :, None, None], grid, 0))
    assert np.array_equal(vc.apply_colored_mask_to_voxel_grid(occ, ext),
                          np.where((occ == 1)[..., None], ext.transpose(1, 0, 2)[:, :, None, :], 0))
    with pytest.raises(ValueError):
        vc.carve_voxel_grid_with_masks(occ, np.ones((5, 7), np.uint8))


def test_guided_carve_with_overlapping_component_boxes(vc, oracle, capsys):
    """left_right_guided_carve when the bounding boxes of two components of the colour interlock (two L-shaped blobs),
    when one box contains another, and with many small components: the reference's loop order decides which paste
    survives (:199-201), so the batched path falls back to one paste per component in id order; arrays and printed log
    against the oracle, at every notebook angle."""
    cfg = pkg("utils.config")
    colour, other = cfg.PART_COLORS_NP["front_minarets"], cfg.PART_COLORS_NP["dome"]
    rng = np.random.default_rng(9)
    W, H, D = 40, 30, 40
    grid = np.zeros((W, H, D, 3), np.uint8)
    grid[2:20, 2:6, 2:6] = colour                 # L number one
    grid[2:6, 2:22, 2:6] = colour
    grid[8:24, 18:22, 2:6] = colour               # L number two, interlocking boxes, not touching (gap at x = 6, 7 / y = 6..17)
    grid[20:24, 8:22, 2:6] = colour
    grid[26:38, 4:28, 10:36] = other              # a big blob of another colour ...
    grid[30:33, 10:14, 20:24] = colour            # ... with a small component of the colour inside its box
    grid[28:36, 6:26, 12:14] = colour             # and a plate whose box contains the small one's x/y range
    for _ in range(25):                           # speckle: many tiny components
        x, y, z = (int(rng.integers(1, n - 1)) for n in (W, H, D))
        if not grid[x - 1:x + 2, y - 1:y + 2, z - 1:z + 2].any():
            grid[x, y, z] = colour
    sem = np.zeros((H, W, 3), np.uint8)
    sem[:, :] = cfg.PART_COLORS_NP["background"]
    sem[2:24, 1:26] = colour
    sem[5:27, 27:37] = colour
    sem[rng.random((H, W)) < 0.1] = cfg.PART_COLORS_NP["background"]
    assert vc._boxes_overlap([(2, 2, 2, 18, 20, 4), (8, 8, 2, 16, 14, 4)])
    for angle in (5, 45, 60, 90):
        capsys.readouterr()
        got = vc.left_right_guided_carve(grid, sem, colour, angle)
        log = capsys.readouterr().out
        want_log = []
        want = oracle.left_right_guided_carve(grid, sem, colour, angle, log=want_log)
        assert np.array_equal(got, want), angle
        assert log.strip("\n") == "\n".join(want_log), angle


def test_device_tensor_chain(vc, carve_golden):
    """CUDA tensors in -> CUDA tensors out, same bytes as the NumPy path."""
    g = carve_golden
    cfg = pkg("utils.config")
    key = "real_Bibi_64"
    sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
    grid = vc.global_carve(binm, ext, 90, return_tensor=True)
    assert isinstance(grid, torch.Tensor) and grid.is_cuda
    out = vc.partwise_carve(grid, ext, sem, cfg.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
    assert isinstance(out, torch.Tensor) and np.array_equal(out.cpu().numpy(), g[key + "_partwise"])


def test_global_carve_512_synthetic_vs_oracle(vc, oracle):
    """Config-4-sized carve (512^3, vector store path) against the oracle's scipy restatement."""
    syn = pkg("synthetic")
    cfg = pkg("utils.config")
    N = 512
    lab = syn.monument_labels(N).numpy()
    front = lab.max(axis=0)[::-1]                              # (y down, x): silhouette labels seen from the front
    lut = syn.label_lut()
    lut0 = lut.copy()
    lut0[0] = cfg.PART_COLORS["background"]
    ext = lut0[front]
    binm = (front > 0).astype(np.uint8)
    got = vc.global_carve(binm, ext, 90)
    want = oracle.global_carve(binm, ext, 90)
    assert np.array_equal(got, want)


def test_x_slabs_tile_the_full_grid(vc, carve_golden):
    """global_carve(..., x_range) == the same rows of the full grid (the multi-GPU sharding unit), on the bit-packed
    fast path (D % 32 == 0), the table path and the general-angle path."""
    g = carve_golden
    for key, interval in (("syn_sq64", 90), ("syn_rect40x64", 90), ("syn_rect33x48", 90), ("syn_rect40x64", 45)):
        ext, binm = g[key + "_ext"], g[key + "_bin"]
        full = vc.global_carve(binm, ext, interval)
        W = full.shape[0]
        for a, b in ((0, W), (0, 17), (17, 40), (40, W), (5, 5)):
            assert np.array_equal(vc.global_carve(binm, ext, interval, x_range=(a, b)), full[a:b]), (key, a, b)
    with pytest.raises(ValueError):
        vc.global_carve(g["syn_sq64_bin"], g["syn_sq64_ext"], 90, x_range=(10, 99))


def test_part_carve_on_asymmetric_grids_vs_oracle(vc, oracle):
    """The clear pass of the bit-level part_carve (runs whose rotated source voxel is empty) only has work on grids
    that are NOT 4-way symmetric; global_carve's output never is.  Random sparse RGB grids with holes, every group at
    90 degrees, W = D multiples of 32 (bit path) and one odd width (table path), against the oracle."""
    rng = np.random.default_rng(77)
    names = ["full_building", "plinth", "dome", "front_minarets"]
    jobs = [([n], 90) for n in names]
    # (256, 5): one full 256 x 256 bit tile per row (vector loads of the clear pass); (512, 3): 2 x 2 tiles;
    # (288, 4): partial x and z tiles
    for (W, H), uniform in (((64, 40), False), ((64, 40), True), ((96, 33), True), ((32, 32), False), ((32, 32), True),
                            ((48, 20), True), ((256, 5), True), ((512, 3), True), ((288, 4), False), ((288, 4), True)):
        sem = np.empty((H, W, 3), np.uint8)
        sem[:] = oracle.PART_COLORS["background"]
        lab = rng.integers(0, len(names) + 1, (H, W))
        if uniform:                      # one part almost everywhere: the group term passes, the occupancy term decides
            lab[:] = 2
            lab[:, :3] = 0
            lab[H // 2:, W // 2:] = 3
        for k, n in enumerate(names):
            sem[lab == k + 1] = oracle.PART_COLORS[n]
        grid = np.zeros((W, H, W, 3), np.uint8)
        occ = rng.random((W, H, W)) < 0.55
        grid[occ] = sem.transpose(1, 0, 2)[:, :, None, :].repeat(W, axis=2)[occ]     # column colour where occupied
        got = vc.part_carve(grid, sem, jobs)
        want = oracle.part_carve(grid, sem, jobs)
        assert np.array_equal(got, want), (W, H)
        assert 0 < np.count_nonzero(want.any(-1)) < np.count_nonzero(grid.any(-1))   # the carve removed something


def test_part_carve_x_slabs_tile_the_full_grid(vc, oracle):
    """part_carve(..., x_range=(a, b)) -- the multi-GPU unit: every slab is computed from the whole input without
    exchange and the slabs concatenate to the full result.  Random asymmetric grids (the clear pass has work), slab
    borders inside and on 256-voxel tiles, a width whose bit rows are vector-aligned (512) and one that is not (288),
    and a job list with a general angle (full computation + slice)."""
    rng = np.random.default_rng(5)
    names = ["full_building", "plinth", "dome", "front_minarets"]
    for (W, H), cuts, jobs in (((512, 3), (0, 100, 256, 300, 512), [([n], 90) for n in names]),
                               ((288, 4), (0, 1, 33, 287, 288), [([n], 90) for n in names]),
                               ((64, 9), (0, 20, 64), [([n], 90) for n in names]),
                               ((32, 6), (0, 7, 32), [(["dome"], 45), (["plinth"], 90)])):
        sem = np.empty((H, W, 3), np.uint8)
        sem[:] = oracle.PART_COLORS["background"]
        lab = np.full((H, W), 2)
        lab[:, :3] = 0
        lab[H // 2:, W // 2:] = 3
        lab[0, ::5] = 1
        for k, n in enumerate(names):
            sem[lab == k + 1] = oracle.PART_COLORS[n]
        grid = np.zeros((W, H, W, 3), np.uint8)
        occ = rng.random((W, H, W)) < 0.6
        grid[occ] = sem.transpose(1, 0, 2)[:, :, None, :].repeat(W, axis=2)[occ]
        full = vc.part_carve(grid, sem, jobs)
        assert np.array_equal(full, oracle.part_carve(grid, sem, jobs)), (W, H)
        slabs = [vc.part_carve(grid, sem, jobs, x_range=(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
        assert [s.shape[0] for s in slabs] == [b - a for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(np.concatenate(slabs, axis=0), full), (W, H)
    with pytest.raises(ValueError):
        vc.part_carve(grid, sem, jobs, x_range=(3, 99))


def test_part_carve_sharded_input_two_slabs_emulated(vc, oracle):
    """PartCarveSlab: each 'rank' sees only its rows of the grid; the occupancy rows are exchanged by hand here (the
    all-gather of utils.sweep.part_carve_sharded) and the two output slabs concatenate to the oracle's part_carve."""
    rng = np.random.default_rng(9)
    names = ["full_building", "plinth", "dome", "front_minarets"]
    jobs = [([n], 90) for n in names]
    for (W, H), cut in (((64, 7), 32), ((512, 2), 256), ((96, 5), 40)):
        sem = np.empty((H, W, 3), np.uint8)
        sem[:] = oracle.PART_COLORS["background"]
        lab = np.full((H, W), 2)
        lab[:, :3] = 0
        lab[H // 2:, W // 2:] = 3
        for k, n in enumerate(names):
            sem[lab == k + 1] = oracle.PART_COLORS[n]
        grid = np.zeros((W, H, W, 3), np.uint8)
        occ = rng.random((W, H, W)) < 0.6
        grid[occ] = sem.transpose(1, 0, 2)[:, :, None, :].repeat(W, axis=2)[occ]
        a = vc.PartCarveSlab(np.ascontiguousarray(grid[:cut]), sem, jobs, W, (0, cut)).begin()
        b = vc.PartCarveSlab(np.ascontiguousarray(grid[cut:]), sem, jobs, W, (cut, W)).begin()
        a.occ[cut:] = b.occ[cut:]
        b.occ[:cut] = a.occ[:cut]
        got = np.concatenate([a.finish(), b.finish()], axis=0)
        assert np.array_equal(got, oracle.part_carve(grid, sem, jobs)), (W, H)
    with pytest.raises(ValueError):
        vc.PartCarveSlab(np.ascontiguousarray(grid[:cut]), sem, [(["dome"], 45)], W, (0, cut))


def test_part_carve_asymmetric_golden_full_and_slabs(vc):
    """The live reference's part_carve of asymmetric grids (tests/golden/partcarve_asym_golden.npz): the full call, the
    replicated-input slabs and the sharded-input slab pair all reproduce it byte for byte."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "partcarve_asym_golden.npz"))
    for tag in g["asym_cases"]:
        key = f"asym_{tag}"
        grid, ext, want = g[key + "_grid"], g[key + "_ext"], g[key + "_partcarve"]
        W = grid.shape[0]
        assert np.array_equal(vc.part_carve(grid, ext, GROUP_JOBS), want), tag
        cut = W // 2 + 3
        slabs = [vc.part_carve(grid, ext, GROUP_JOBS, x_range=r) for r in ((0, cut), (cut, W))]
        assert np.array_equal(np.concatenate(slabs, axis=0), want), tag
        if W % 32 == 0:
            a = vc.PartCarveSlab(np.ascontiguousarray(grid[:cut]), ext, GROUP_JOBS, W, (0, cut)).begin()
            b = vc.PartCarveSlab(np.ascontiguousarray(grid[cut:]), ext, GROUP_JOBS, W, (cut, W)).begin()
            a.occ[cut:] = b.occ[cut:]
            b.occ[:cut] = a.occ[:cut]
            assert np.array_equal(np.concatenate([a.finish(), b.finish()], axis=0), want), tag


def _random_part_scene(rng, oracle, W, H, names, uniform, density=0.55):
    sem = np.empty((H, W, 3), np.uint8)
    sem[:] = oracle.PART_COLORS["background"]
    lab = rng.integers(0, len(names) + 1, (H, W))
    if uniform:
        lab[:] = 2
        lab[:, :3] = 0
        lab[H // 2:, W // 2:] = 3
    for k, n in enumerate(names):
        sem[lab == k + 1] = oracle.PART_COLORS[n]
    grid = np.zeros((W, H, W, 3), np.uint8)
    occ = rng.random((W, H, W)) < density
    grid[occ] = sem.transpose(1, 0, 2)[:, :, None, :].repeat(W, axis=2)[occ]
    return sem, grid


def test_ragged_widths_take_the_bit_path_and_match_the_oracle(vc, oracle):
    """Widths that are no multiple of 32 (and heights that make the voxel count no multiple of 16: byte-wise tails) on
    the flat-group bit kernels: global_carve with x-slabs and part_carve on asymmetric grids (the clear pass works
    voxel by voxel there), against the oracle."""
    rng = np.random.default_rng(1234)
    names = ["full_building", "plinth", "dome", "front_minarets"]
    jobs = [([n], 90) for n in names]
    dev = torch.device("cuda")
    for (W, H), uniform in (((16, 3), False), ((17, 5), True), ((31, 7), False), ((48, 9), True), ((63, 4), False),
                            ((88, 11), True), ((100, 3), False), ((177, 6), True), ((246, 2), True), ((255, 3), False)):
        assert vc._fold_plan(W, H, W, dev)[1] is not None
        sem, grid = _random_part_scene(rng, oracle, W, H, names, uniform)
        got = vc.part_carve(grid, sem, jobs)
        want = oracle.part_carve(grid, sem, jobs)
        assert np.array_equal(got, want), (W, H)
        assert 0 < np.count_nonzero(want.any(-1)) < np.count_nonzero(grid.any(-1))
        cut = W // 2 + 1
        slabs = [vc.part_carve(grid, sem, jobs, x_range=r) for r in ((0, cut), (cut, W))]
        assert np.array_equal(np.concatenate(slabs, axis=0), want), (W, H)
        # global_carve: blobby binary mask + the random semantic image as colours
        binm = (rng.random((H, W)) < 0.7).astype(np.uint8)
        binm[:, W // 3:W // 3 + 2] = 1
        full = vc.global_carve(binm, sem, 90)
        assert np.array_equal(full, oracle.global_carve(binm, sem, 90)), (W, H)
        for a, b in ((0, 1), (1, cut), (cut, W)):
            assert np.array_equal(vc.global_carve(binm, sem, 90, x_range=(a, b)), full[a:b]), (W, H, a, b)
    with pytest.raises(ValueError):                          # the sharded-input form still needs whole 32-voxel words
        vc.PartCarveSlab(np.ascontiguousarray(grid[:cut]), sem, jobs, W, (0, cut))


def test_ragged_random_shapes_vs_oracle(vc, oracle):
    """Thirty random (W, H) with W in 16..90 (every residue mod 16 and mod 32, H from 1) -- rows whose byte length is no
    multiple of 4, groups straddling rows at every offset, voxel counts with every remainder mod 16 -- global_carve and
    part_carve against the oracle."""
    rng = np.random.default_rng(2024)
    names = ["full_building", "plinth", "dome", "front_minarets"]
    jobs = [([n], 90) for n in names]
    seen = set()
    for t in range(30):
        W, H = int(rng.integers(16, 91)), int(rng.integers(1, 8))
        seen.add((W * H * W) % 16)
        sem, grid = _random_part_scene(rng, oracle, W, H, names, uniform=bool(t % 2), density=float(rng.uniform(0.3, 0.8)))
        assert np.array_equal(vc.part_carve(grid, sem, jobs), oracle.part_carve(grid, sem, jobs)), (W, H)
        binm = (rng.random((H, W)) < rng.uniform(0.4, 0.95)).astype(np.uint8)
        assert np.array_equal(vc.global_carve(binm, sem, 90), oracle.global_carve(binm, sem, 90)), (W, H)
    assert len(seen) >= 6
