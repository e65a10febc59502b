"""Drop-in boundary (SURVEY 8 b): with the package directory on sys.path in place of the reference checkout, every
`from utils.<module> import <name>` of notebooks 1-4 resolves (import cells only; no GPU needed)."""
import os
import subprocess
import sys

from conftest import ROOT, PKG

NOTEBOOK_IMPORTS = """
from utils.config import PART_COLORS, PART_COLORS_NP, INTERIOR_PARTS
from utils.mask_utils import load_and_prepare_masks, load_mask
from utils.voxel_carving_utils import global_carve, partwise_carve
from utils.voxel_utils import voxel_grid_to_points, meshify_colored_voxel_grid, get_voxel_points_by_parts
from utils.visualization import plot_voxel, visualize_mesh_plotly
from utils.camera_estimation import *
from utils.camera_geometry import project
from utils.projection_utils import project_colored_voxels, visualize_reprojection
from utils.deformation_estimation import launch_deform_viewer_fixed_camera
from utils.eval_helpers_intra import *
from utils.config import *
for name in ("extract_minaret_kps_for_view", "auto_compute_initial_params_matching_bbox", "optimize_camera_with_keypoints",
             "visualize_voxel_projection_iou", "launch_smart_aligner", "compute_partwise_iou", "run_minaret_kp_evaluation",
             "run_minaret_iou_evaluation", "run_part_minaret_binary_iou", "compute_global_depth_buffer",
             "project_part_visible", "load_camera_json", "MONUMENT_CONFIG", "ROOT_PATH", "MAX_DIM"):
    assert name in globals(), name
print("ok")
"""


def test_notebook_import_cells_resolve():
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, PKG))
    out = subprocess.run([sys.executable, "-c", NOTEBOOK_IMPORTS], capture_output=True, text=True, timeout=300, env=env,
                         cwd=os.path.join(ROOT, "tests"))
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-3000:]
