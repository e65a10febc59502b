"""The three evaluation tables of notebook 4 (utils/eval_helpers_intra.py:287-748) against what the live
reference prints for the half-resolution Taj scene (tests/golden/make_golden.py tables): the oracle's numbers on the
CPU, the package's drivers (tables and DataFrames, byte for byte) on the GPU."""
import contextlib
import io
import json
import os
import shutil

import numpy as np
import pytest

from conftest import DATA, GOLDEN, pkg


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "tables_golden.npz"))


@pytest.fixture(scope="module")
def scene(tmp_path_factory, g):
    """nb4 cell 3 directory layout with the half-resolution Taj scene (same construction as make_golden.build_table_scene)."""
    tmp = tmp_path_factory.mktemp("nb4")
    ag = np.load(os.path.join(GOLDEN, "aligner_golden.npz"))
    dg = np.load(os.path.join(GOLDEN, "deform_golden.npz"))
    grid = ag["grid"]
    deformed = np.zeros((int(np.prod(grid.shape[:3])), 3), np.uint8)
    deformed[dg["f64_grid_nz"]] = dg["f64_grid_rgb"]
    deformed = deformed.reshape(grid.shape)
    d = {k: str(tmp / k) for k in ("voxels", "deformed", "cams", "data")}
    for v in d.values():
        os.makedirs(v)
    os.makedirs(os.path.join(d["data"], "Taj", "masks"))
    np.savez_compressed(os.path.join(d["voxels"], "Taj_voxel_grid.npz"), voxel_grid=grid)
    np.savez_compressed(os.path.join(d["deformed"], "Taj_deformed_voxel_grid.npz"), voxel_grid=deformed)
    shutil.copy(os.path.join(DATA, "Taj", "masks", "Taj_front_mask.png"), os.path.join(d["data"], "Taj", "masks"))
    cams = json.loads(str(g["cams_json"]))
    for tag, cam in cams.items():
        json.dump(cam, open(os.path.join(d["cams"], f"Taj_camera_params_{tag}.json"), "w"))
    d.update(grid=grid, deformed_grid=deformed, cams=cams, cam_dir=d["cams"])
    return d


def cam32(c):
    c = c["front"]
    return {"cam_pos": np.array(c["cam_pos"], np.float32), "target": np.array(c["target"], np.float32), "f": float(c["f"]),
            "cx": float(c["cx"]), "cy": float(c["cy"])}


def resized_mask(grid):
    import cv2
    m = cv2.cvtColor(cv2.imread(os.path.join(DATA, "Taj", "masks", "Taj_front_mask.png")), cv2.COLOR_BGR2RGB)
    H, W = m.shape[:2]
    s = max(grid.shape[:3]) / max(H, W)
    return cv2.resize(m, (int(round(W * s)), int(round(H * s))), interpolation=cv2.INTER_NEAREST)


def cells(g, name):
    return json.loads(str(g[f"{name}_df"]))["TM"]


# ---- CPU: the oracle's numbers format to the reference's table cells -----------------------------------------------
def test_oracle_tables(oracle, g, scene):
    grid, mask = scene["grid"], resized_mask(scene["grid"])
    colours = [oracle.PART_COLORS["front_minarets"], oracle.PART_COLORS["back_minarets"]]
    cams = {tag: cam32(c) for tag, c in scene["cams"].items()}
    err = oracle.minaret_kp_errors(grid, mask, {"init": cams["init"], "rep": cams["kp"]}, colours, back_top_only=False)
    want = cells(g, "kp")
    for m in ("LM1", "RM1", "LM2", "RM2"):
        assert f"{err['init'][m]:.2f}→{err['rep'][m]:.2f}" == want[m]
    assert f"{np.mean(list(err['init'].values())):.2f}→{np.mean(list(err['rep'].values())):.2f}" == want["Average"]
    iou = oracle.minaret_visible_ious(grid, mask, {"init": cams["init"], "rep": cams["kp"], "final": cams["final"]}, colours)
    want = cells(g, "iou")
    for m in ("LM1", "RM1", "LM2", "RM2"):
        assert f"{iou[m]['init']:.3f}→{iou[m]['rep']:.3f}→{iou[m]['final']:.3f}" == want[m]
    part = oracle.part_minaret_binary_ious(grid, scene["deformed_grid"], mask, cams["final"], oracle.PART_COLORS)
    want = cells(g, "part")
    for row, v in part.items():
        assert ("--" if v is None else f"{v[0]:.3f}→{v[1]:.3f}") == want[row], row


# ---- GPU: the package's drivers print the reference's tables ---------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kp", "iou", "part"])
def test_table_drivers_gpu(oracle, g, scene, name):
    eh = pkg("utils.eval_helpers_intra")
    common = dict(monuments=["Taj"], view="front", root_voxels=scene["voxels"], root_masks=scene["data"],
                  cam_dir=scene["cam_dir"], part_colors=oracle.PART_COLORS, visualize=False)
    fn = {"kp": eh.run_minaret_kp_evaluation, "iou": eh.run_minaret_iou_evaluation, "part": eh.run_part_minaret_binary_iou}[name]
    if name == "part":
        common["deformed_voxels"] = scene["deformed"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        df = fn(**common)
    assert json.loads(df.to_json()) == json.loads(str(g[f"{name}_df"]))
    assert buf.getvalue() == str(g[f"{name}_log"])


@pytest.mark.gpu
def test_binary_gt_and_mixed_dtype_visibility(oracle, scene):
    eh = pkg("utils.eval_helpers_intra")
    ce = pkg("utils.camera_estimation")
    grid, mask = scene["grid"], resized_mask(scene["grid"])
    assert np.array_equal(eh.compute_binary_gt(mask, grid), oracle.compute_binary_gt(mask, grid))
    cam = cam32(scene["cams"]["final"])
    H, W = mask.shape[:2]
    zbuf = eh.compute_global_depth_buffer(grid, cam, H, W)
    colours = [oracle.PART_COLORS["front_minarets"], oracle.PART_COLORS["back_minarets"]]
    for name, coords in ce.extract_minaret_voxels_by_label(grid, colours).items():
        assert coords.dtype == np.int64                      # float64 projection through the float32 look-at
        assert np.array_equal(eh.project_part_visible(coords, cam, zbuf, H, W), oracle.project_part_visible(coords, cam, zbuf, H, W)), name
