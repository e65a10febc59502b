"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, the Python binding covers them, and the product refuses to run without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import PKG, ROOT, pkg

HEADER = os.path.join(ROOT, "include", "p3d_b200.h")
LIB = os.path.join(ROOT, PKG, "csrc", "libp3d_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first: python __graft_entry__.py"
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in p3d_b200.h but not exported"


def test_binding_covers_header():
    nv = pkg("utils._native")
    missing = [n for n in declared_symbols() if n not in nv._SIGNATURES]
    assert not missing, missing
    assert nv.lib.p3d_version() >= 100


def test_pure_host_entry_points_work_without_gpu():
    nv = pkg("utils._native")
    assert nv.lib.p3d_points_workspace_bytes(4096 * 10) == 11 * 8
    assert nv.lib.p3d_sweep_workspace_bytes(8, 64, 64, 2, 8) > 8 * 64 * 64 * 4
    assert nv.lib.p3d_sweep_workspace_bytes(0, 64, 64, 2, 8) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_cuda():
    import numpy as np
    nv = pkg("utils._native")
    ce = pkg("utils.camera_estimation")
    pu = pkg("utils.projection_utils")
    with pytest.raises(nv.P3DError):
        ce.compute_partwise_iou(np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4, 3), np.uint8), {"a": (1, 2, 3)})
    with pytest.raises(nv.P3DError):
        pu.project_colored_voxels(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.uint8), np.zeros(3), np.ones(3),
                                  1.0, 0.0, 0.0, 4, 4)
    vc = pkg("utils.voxel_carving_utils")
    sem = np.zeros((4, 32, 3), np.uint8)
    with pytest.raises(nv.P3DError):
        vc.part_carve(np.zeros((32, 4, 32, 3), np.uint8), sem, [(["dome"], 90)], x_range=(0, 16))
    with pytest.raises(nv.P3DError):
        vc.PartCarveSlab(np.zeros((16, 4, 32, 3), np.uint8), sem, [(["dome"], 90)], 32, (0, 16))


def test_product_never_imports_the_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, PKG)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "p3d_oracle" in src:
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def test_group_image_matches_the_per_group_formula():
    """Host logic of part_carve: bit g of the group image = m_g & _mask_to_wh(m_g) (m_g.T for a square image, the
    reference's quirk at voxel_carving_utils.py:19-28), built with one transpose for all groups."""
    import numpy as np
    vc = pkg("utils.voxel_carving_utils")
    rng = np.random.default_rng(0)
    for H, W in ((17, 17), (64, 64), (40, 64), (33, 48), (1, 1)):
        jobs = [(rng.random((H, W)) < 0.4, 90) for _ in range(7)]
        want = np.zeros((H, W), np.uint32)
        for g, (m, _) in enumerate(jobs):
            want |= ((m & m.T) if W == H else m).astype(np.uint32) << np.uint32(g)
        assert np.array_equal(vc._group_image(jobs, H, W), want)


def test_sweep_workspace_follows_the_batch_rule():
    """p3d_sweep_workspace_bytes (host arithmetic only): one z-buffer set while all cameras fit the 128 MB floor, two
    sets (double-buffered batches) of at most 512 MB each beyond it, never fewer than 16 cameras per set at 2048^2."""
    nv = pkg("utils._native")
    ws = lambda K, H, W: int(nv.lib.p3d_sweep_workspace_bytes(K, H, W, 9, 8))
    mb = 1 << 20
    per = 1024 * 1024 * 4
    small = ws(32, 1024, 1024)                       # 32 cameras = the floor at 1024^2: one set of 32
    assert 32 * per <= small < 32 * per + 2 * mb
    big = ws(65536, 1024, 1024)                      # two sets of 128 cameras (512 MB each) + per-camera blocks
    assert 2 * 128 * per <= big < 2 * 128 * per + 64 * mb
    assert ws(33, 1024, 1024) >= 2 * 33 * per        # just above the floor: two sets, capacity limited by K
    huge = ws(4096, 2048, 2048)                      # 2048^2: 16 cameras per set (268 MB, the floor rule), two sets
    assert 2 * 16 * 4 * per <= huge < 2 * 32 * 4 * per + 8 * mb


def test_job_colours_are_cached_per_name_structure():
    """Host logic of part_carve: the per-group colour lists come from PART_COLORS by part name (cached per name
    structure) and are recognised as already flat by the group-image helper; anything else is normalised."""
    import numpy as np
    vc = pkg("utils.voxel_carving_utils")
    cfg = pkg("utils.config")
    jobs = [(["dome", "plinth"], 90), (["front_minarets"], 90), ([], 90)]
    flat = vc._job_colours(jobs)
    assert flat == tuple(tuple(tuple(int(v) for v in cfg.PART_COLORS[n]) for n in names) for names, _ in jobs)
    assert vc._job_colours([(list(n), 45) for n, _ in jobs]) is flat          # the angle is not part of the key
    assert vc._flat_colours(flat) is flat
    loose = [[np.asarray(cfg.PART_COLORS[n]) for n in names] for names, _ in jobs]
    assert vc._flat_colours(loose) == flat


def test_bit_workspace_covers_ragged_rows():
    """p3d_part_carve_bits_workspace_bytes (host arithmetic): two z-packed bit arrays of ceil(D/32) words per row plus
    the per-group mask rows, for widths that are no multiple of 32 as well."""
    nv = pkg("utils._native")
    for W, H, G in ((177, 256, 6), (88, 128, 6), (256, 160, 4), (16, 3, 1)):
        words, xwp = (W + 31) // 32, (W + 31) // 32 + 2
        need = 2 * W * H * words * 4 + G * H * xwp * 4
        got = int(nv.lib.p3d_part_carve_bits_workspace_bytes(W, H, W, G))
        assert need <= got <= need + 3 * 256, (W, H, G)
