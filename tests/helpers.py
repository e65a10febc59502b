import hashlib

import numpy as np

GROUP_JOBS = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
              (["small_minarets"], 90), (["dome"], 90)]                       # notebook 1, cell 7
PART_SYMMETRY = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
EXTRUSION_DEPTHS = {"main_door": 20, "windows": 10}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape)


def row_to_args(row, dtype=np.float64):
    """9-vector -> (cam_pos, target, f, cx, cy) typed the way the reference receives them."""
    cp, tg = np.asarray(row[0:3], dtype=dtype), np.asarray(row[3:6], dtype=dtype)
    if dtype == np.float32:
        return cp, tg, np.float32(row[6]), np.float32(row[7]), np.float32(row[8])
    return cp, tg, float(row[6]), float(row[7]), float(row[8])
