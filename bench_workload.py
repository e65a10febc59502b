"""NumPy-only definition of bench.py's synthetic workload, shared by both arms.

The GPU arm builds the same scene on the device with the package's `synthetic.py`; this module restates it with NumPy
so that the reference arm (`bench.py --impl reference`) never imports the product package.  tests/test_bench_workload.py
checks that both give identical grids, cameras and candidates.  Geometry: SURVEY.md 8(d) "synthetic monument"; candidate
perturbations: the reference's random-search step sizes (utils/camera_estimation.py:611-625), seed 20240607.
"""
from __future__ import annotations

import numpy as np

PART_COLORS = {                                                    # utils/config.py:29-40 of the reference
    "full_building": (253, 248, 96), "chhatris": (1, 220, 5), "plinth": (63, 138, 173), "dome": (190, 0, 255),
    "front_minarets": (0, 0, 255), "back_minarets": (5, 223, 223), "small_minarets": (255, 180, 80),
    "main_door": (180, 140, 255), "windows": (255, 120, 230), "background": (216, 224, 251),
}
PART_NAMES = [k for k in PART_COLORS if k != "background"]            # label = index + 1
LABEL = {name: i + 1 for i, name in enumerate(PART_NAMES)}
CANDIDATE_SEED = 20240607
TOTAL_CANDIDATES = 65536
HIDDEN_DELTA = np.array([3.0, -2.0, 5.0, 1.0, 2.0, -3.0, 4.0, 1.5, -2.5])
STEP_SIZES = np.array([50, 50, 100, 50, 50, 100, 50, 20, 20], dtype=np.float64)   # camera_estimation.py:611-617


def label_lut() -> np.ndarray:
    lut = np.zeros((256, 3), np.uint8)
    for name, lab in LABEL.items():
        lut[lab] = PART_COLORS[name]
    return lut


def monument_labels(N: int, chunk: int = 32) -> np.ndarray:
    """(N,N,N) uint8 label grid, axes (z,y,x): plinth, body, dome, chhatris, minarets, door, windows."""
    if N % 32:
        raise ValueError("N must be a multiple of 32")
    q = N // 32

    def u(v):                                     # 256-scale coordinate -> voxels
        return (v * q) // 8

    out = np.zeros((N, N, N), np.uint8)
    y = np.arange(N, dtype=np.int64).reshape(1, N, 1)
    x = np.arange(N, dtype=np.int64).reshape(1, 1, N)
    for z0 in range(0, N, chunk):
        z = np.arange(z0, min(N, z0 + chunk), dtype=np.int64).reshape(-1, 1, 1)
        lab = out[z0:z0 + z.shape[0]]

        def box(x0, x1, y0, y1, zz0, zz1):
            return (x >= u(x0)) & (x < u(x1)) & (y >= u(y0)) & (y < u(y1)) & (z >= u(zz0)) & (z < u(zz1))

        def cyl(cx, cz, r, y0, y1):
            return ((x - u(cx)) ** 2 + (z - u(cz)) ** 2 <= u(r) ** 2) & (y >= u(y0)) & (y < u(y1))

        lab[box(28, 228, 0, 24, 28, 228)] = LABEL["plinth"]
        lab[box(68, 188, 24, 120, 68, 188)] = LABEL["full_building"]
        lab[((x - u(128)) ** 2 + (y - u(116)) ** 2 + (z - u(128)) ** 2 <= u(44) ** 2) & (y >= u(120))] = LABEL["dome"]
        for cx, cz in ((84, 84), (172, 84), (84, 172), (172, 172)):
            lab[cyl(cx, cz, 10, 120, 150)] = LABEL["chhatris"]
        for cx, cz in ((72, 72), (184, 72), (72, 184), (184, 184)):
            lab[cyl(cx, cz, 4, 120, 168)] = LABEL["small_minarets"]
        for cx, cz, name in ((44, 44, "front_minarets"), (212, 44, "front_minarets"),
                             (44, 212, "back_minarets"), (212, 212, "back_minarets")):
            lab[cyl(cx, cz, 9, 24, 200)] = LABEL[name]
        lab[box(116, 140, 24, 72, 64, 68)] = LABEL["main_door"]
        for x0 in (80, 160):
            lab[box(x0, x0 + 16, 56, 88, 64, 68)] = LABEL["windows"]
    return out


def points_of(labels: np.ndarray, names):
    """get_voxel_points_by_parts (utils/voxel_utils.py:7-21) of the label grid: (pts f32 (n,3) [x,y,z], colours u8 (n,3))
    in ascending flat index."""
    keep = np.isin(labels, [LABEL[n] for n in names])
    flat = np.flatnonzero(keep)
    a0, a1, a2 = np.unravel_index(flat, labels.shape)
    return np.stack([a2, a1, a0], axis=1).astype(np.float32), label_lut()[labels.reshape(-1)[flat]]


def base_camera(N: int, H: int, W: int, view: str = "front") -> np.ndarray:
    s = N / 256.0
    if view == "front":
        cam = (128 * s + 7.3, 70 * s + 3.1, -330 * s)
        tgt = (128 * s, 80 * s, 128 * s)
        f = 0.62 * 458.0 / 200.0 * W
    elif view == "aerial":
        cam = (300 * s, 330 * s, -250 * s)
        tgt = (128 * s, 70 * s, 128 * s)
        f = 0.5 * 560.0 / 200.0 * W
    else:
        raise ValueError(view)
    return np.array([*cam, *tgt, f, W / 2.0 + 1.7, H * 0.62], dtype=np.float64)


def candidates(base_row: np.ndarray, K: int, seed: int = CANDIDATE_SEED) -> np.ndarray:
    """Base camera at index 0 plus K-1 perturbations base + U(-1,1) * step sizes (order cam_pos, target, f, cx, cy)."""
    rng = np.random.default_rng(seed)
    base_row = np.asarray(base_row, dtype=np.float64)
    out = np.empty((K, 9), np.float64)
    if K > 0:
        out[0] = base_row
    out[1:] = base_row + rng.uniform(-1.0, 1.0, size=(max(K - 1, 0), 9)) * STEP_SIZES
    return out


def bench_config(N: int, H: int, W: int, parts, n_points: int, cands_per_step: int) -> dict:
    """The `config` object of bench.py's JSON line -- the same dict in both arms."""
    return {"workload": f"synthetic {N}^3 semantic monument, {len(parts)} parts ({n_points} points), {H}x{W} label mask, "
                        f"{TOTAL_CANDIDATES} candidate cameras (base + U(-1,1) * reference step sizes, seed {CANDIDATE_SEED})",
            "grid": N, "mask": [H, W], "parts": len(parts), "points": int(n_points),
            "candidates_total": TOTAL_CANDIDATES, "candidates_per_gpu_per_step": int(cands_per_step),
            "l2": "inputs larger than L2 (point list %.0f MB streamed by every launch); fresh candidates each step"
                  % (n_points * 13 / 1e6)}
