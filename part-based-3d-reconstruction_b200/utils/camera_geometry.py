"""Drop-in for the reference's utils/camera_geometry.py."""
from __future__ import annotations

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv

_DEFAULT_UP = np.array([0, 1, 0], dtype=np.float32)


def _np_dtype(a) -> np.dtype:
    if isinstance(a, torch.Tensor):
        return np.dtype(str(a.dtype).split(".")[-1])
    return a.dtype if isinstance(a, np.ndarray) else np.asarray(a).dtype


def working_dtype(*arrays) -> np.dtype:
    """NumPy promotion of `pts3d - cam_pos` / `target - eye`: float32 only if everything is
    float32, otherwise float64 (projection_utils.py:7, camera_geometry.py:4)."""
    t = np.result_type(*[_np_dtype(a) for a in arrays])
    return np.dtype(np.float32) if t == np.float32 else np.dtype(np.float64)


def candidate_row(cam_pos, target, f, cx, cy, dtype) -> np.ndarray:
    row = np.empty(9, dtype=dtype)
    row[0:3] = np.asarray(cam_pos, dtype=dtype).reshape(3)
    row[3:6] = np.asarray(target, dtype=dtype).reshape(3)
    row[6], row[7], row[8] = f, cx, cy
    return row


@nv.on_device
def look_at_rotation(eye, target, up=_DEFAULT_UP, device=None):
    """camera_geometry.py:3-14, evaluated on the GPU with the reference's operation order.
    Returns a (3,3) ndarray with rows [x; y; z] in the promoted dtype of (eye, target)."""
    if not np.array_equal(np.asarray(up), _DEFAULT_UP):
        raise NotImplementedError("only the reference's default up=[0,1,0] is supported")
    dev = nv.require_cuda(device)
    dt = working_dtype(eye, target)
    row = candidate_row(eye, target, 0.0, 0.0, 0.0, dt)
    cams = eng.setup_cameras(torch.from_numpy(row[None]).to(dev))
    return cams[0, 3:12].reshape(3, 3).cpu().numpy()


def project(pt3d, cam_pos, target, f, cx, cy):
    """camera_geometry.py:17-27: un-rounded (u, v) of one point.  Host-side helper used by the
    keypoint optimiser and the reprojection viewer (not on the sweep); the rotation comes from the
    device look_at so that it matches the one the sweep uses."""
    R = look_at_rotation(cam_pos, target)
    X, Y, Z = (np.asarray(pt3d) - np.asarray(cam_pos)) @ R.T
    Z = max(Z, 1e-8)
    return np.array([(X / Z) * f + cx, -(Y / Z) * f + cy])
