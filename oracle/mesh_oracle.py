"""CPU oracle of meshify_colored_voxel_grid -- TEST INFRASTRUCTURE ONLY (tests/ and smoke() may import it).

PARITY UNPINNED for the marching-cubes part.  The reference (utils/voxel_utils.py:53-95) calls
``skimage.measure.marching_cubes(mask, level=0.5)`` -- scikit-image 0.21 (requirements.txt), Lewiner et al.'s tables and
ambiguity tests -- and scikit-image is neither under /root/reference nor installed in this image, so no output of the
reference itself could be recorded.  What is restated here is the PUBLISHED algorithm family (Lorensen & Cline's marching
cubes with a face-consistent ambiguity rule) in a canonical vertex / face order of this project:

  * level 0.5 on a 0/1 volume: every vertex is the midpoint of a grid edge whose two voxels differ in occupancy;
  * vertices ordered by (flat index of the edge's lower voxel, axis); faces by (flat index of the cell's origin voxel, the
    cell's loops by lowest edge id, fan triangles from the first loop vertex whose fan puts no triangle inside a cube face);
  * ambiguous faces separate the two occupied corners (occupancy is 6-connected); triangles are counter-clockwise seen from
    the empty side; normals are the negated central-difference gradient of the volume (replicated border), averaged over the
    edge's two voxels and normalised -- the fallback for a vanishing gradient is the edge direction from the occupied to the
    empty voxel;
  * no padding: like skimage, a surface that reaches the volume border stays open there.

The mesh is therefore the same SURFACE the reference shows (identical vertex set for any correct marching cubes on binary
data; topology differs from Lewiner's only inside cells with ambiguous faces), in another vertex / face order.  It is pinned
on closed-form properties instead (tests/test_mesh.py): watertightness, Euler characteristic, enclosed volume, outward
orientation, vertex set == set of occupancy-changing edges.  Everything around the marching cubes -- stride, axis swap,
the mirror of :79, the nearest-voxel colouring with sklearn (:82-90, the reference's own call) -- follows the reference
line by line.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------------------------------
# the 256-case table, derived from geometry (independently of tools/gen_mc_table.py, which ships the kernel's copy)
# ---------------------------------------------------------------------------------------------------------------------
_AX = np.eye(3, dtype=int)


def _edges():
    """[(lower corner (d0,d1,d2), axis)] in the id order e = 4*axis + 2*u + w."""
    out = []
    for axis in range(3):
        o1, o2 = [a for a in range(3) if a != axis]
        for u in (0, 1):
            for w in (0, 1):
                d = np.zeros(3, int)
                d[o1], d[o2] = u, w
                out.append((tuple(int(v) for v in d), axis))
    return out


EDGES = _edges()
_EDGE_ID = {e: i for i, e in enumerate(EDGES)}


def _mid(e):
    d, axis = EDGES[e]
    return np.asarray(d, float) + 0.5 * _AX[axis]


def _case_loops(case):
    """Closed, directed loops of cut-edge ids of one cell."""
    occ = lambda d: (case >> (4 * d[0] + 2 * d[1] + d[2])) & 1
    succ = {}
    for axis in range(3):
        o1, o2 = [a for a in range(3) if a != axis]
        for side in (0, 1):
            normal = (1.0 if side else -1.0) * _AX[axis]
            corner = lambda u, w: tuple(int(v) for v in (side * _AX[axis] + u * _AX[o1] + w * _AX[o2]))
            ring = [corner(0, 0), corner(1, 0), corner(1, 1), corner(0, 1)]
            cuts = []                                        # (edge id, the occupied corner at one of its ends)
            for k in range(4):
                a, b = ring[k], ring[(k + 1) % 4]
                if occ(a) != occ(b):
                    low = a if sum(a) < sum(b) else b
                    ax = int(np.flatnonzero(np.asarray(a) != np.asarray(b))[0])
                    cuts.append((_EDGE_ID[(low, ax)], a if occ(a) else b))
            if len(cuts) == 2:
                pairs = [(cuts[0], cuts[1])]
            elif len(cuts) == 4:                             # separate the occupied corners: pair the cuts that share one
                pairs = []
                for i in range(4):
                    for j in range(i + 1, 4):
                        if cuts[i][1] == cuts[j][1]:
                            pairs.append((cuts[i], cuts[j]))
                assert len(pairs) == 2
            else:
                assert not cuts
                continue
            for (ea, ca), (eb, _) in pairs:
                P, Q, C = _mid(ea), _mid(eb), np.asarray(ca, float)
                left = float(np.dot(np.cross(Q - P, C - P), normal))       # > 0: occupied corner on the left, seen from outside
                a, b = (ea, eb) if left > 0 else (eb, ea)
                assert a not in succ
                succ[a] = b
    loops, todo = [], set(succ)
    while todo:
        start = min(todo)
        loop, e = [], start
        while True:
            loop.append(e)
            todo.discard(e)
            e = succ[e]
            if e == start:
                break
        loops.append(loop)
    return loops


def _orientation_sign():
    """+1 when a loop's own direction already faces the empty side (checked on the one-corner case)."""
    (loop,) = _case_loops(1)
    A, B, C = (_mid(e) for e in loop)
    return 1 if np.dot(np.cross(B - A, C - A), np.ones(3)) > 0 else -1


def _on_common_face(tri):
    """True when the three cut edges lie on one cube face (the triangle would lie inside that face)."""
    sets = []
    for e in tri:
        d, axis = EDGES[e]
        sets.append({(a, d[a]) for a in range(3) if a != axis})
    return bool(sets[0] & sets[1] & sets[2])


def _fan(loop):
    """Fan from the first vertex of the loop (starting at its lowest edge id) that gives no in-face triangle: the
    neighbouring cell would mirror such a triangle into a zero-thickness fin."""
    k = len(loop)
    for r in range(k):
        rot = loop[r:] + loop[:r]
        tris = [(rot[0], rot[i], rot[i + 1]) for i in range(1, k - 1)]
        if not any(_on_common_face(t) for t in tris):
            return tris
    raise AssertionError(loop)


_TABLE = None


def mc_table():
    """case -> list of triangles (edge-id triples), counter-clockwise seen from the empty side."""
    global _TABLE
    if _TABLE is None:
        sign = _orientation_sign()
        tab = []
        for case in range(256):
            tris = []
            for loop in _case_loops(case):
                for t in _fan(loop):
                    tris.append(t if sign > 0 else (t[0], t[2], t[1]))
            tab.append(tris)
        _TABLE = tab
    return _TABLE


# ---------------------------------------------------------------------------------------------------------------------
# marching cubes on a 0/1 volume at level 0.5
# ---------------------------------------------------------------------------------------------------------------------
def marching_cubes_binary(mask):
    """mask (B0,B1,B2) of 0/1 -> verts (V,3) float32 in (a0,a1,a2), faces (F,3) int32, normals (V,3) float32."""
    m = (np.asarray(mask) != 0).astype(np.int8)
    B = m.shape
    n = m.size
    flat = np.arange(n).reshape(B)
    # vertices: per (voxel, axis) whether the edge to the next voxel along `axis` changes occupancy
    flags = np.zeros(B + (3,), bool)
    for axis in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis], hi[axis] = slice(0, B[axis] - 1), slice(1, B[axis])
        flags[tuple(lo) + (axis,)] = m[tuple(lo)] != m[tuple(hi)]
    vid = np.cumsum(flags.reshape(-1)) - 1                   # vertex id of (voxel, axis) in (flat index, axis) order
    owner, axis_of = np.divmod(np.flatnonzero(flags.reshape(-1)), 3)
    coords = np.stack(np.unravel_index(owner, B), axis=1)
    verts = coords.astype(np.float32)
    verts[np.arange(len(owner)), axis_of] += np.float32(0.5)
    # normals: -(g(p) + g(p + e)) normalised, g = central differences with a replicated border, all in float32
    pad = np.pad(m.astype(np.float32), 1, mode="edge")
    grad = np.stack([(pad[tuple(slice(2, None) if a == k else slice(1, -1) for a in range(3))]
                      - pad[tuple(slice(0, -2) if a == k else slice(1, -1) for a in range(3))]) * np.float32(0.5)
                     for k in range(3)], axis=-1)            # (B0,B1,B2,3)
    upper = coords.copy()
    upper[np.arange(len(owner)), axis_of] += 1
    g = -(grad[tuple(coords.T)] + grad[tuple(upper.T)])
    norm = np.sqrt(g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1] + g[:, 2] * g[:, 2], dtype=np.float32)
    normals = np.zeros_like(g)
    ok = norm > 0
    normals[ok] = g[ok] / norm[ok, None]
    # vanishing gradient: the edge direction from the occupied to the empty voxel
    bad = np.flatnonzero(~ok)
    if len(bad):
        sign = np.where(m[tuple(coords[bad].T)] != 0, 1.0, -1.0).astype(np.float32)
        normals[bad, axis_of[bad]] = sign
    # faces
    tab = mc_table()
    case = np.zeros([b - 1 for b in B], np.int32) if min(B) > 1 else np.zeros((0, 0, 0), np.int32)
    if case.size:
        for c in range(8):
            d = ((c >> 2) & 1, (c >> 1) & 1, c & 1)
            case |= m[d[0]:B[0] - 1 + d[0], d[1]:B[1] - 1 + d[1], d[2]:B[2] - 1 + d[2]].astype(np.int32) << c
    faces = []
    strides = (B[1] * B[2], B[2], 1)
    for i, j, k in zip(*np.nonzero((case != 0) & (case != 255))):
        for tri in tab[int(case[i, j, k])]:
            ids = []
            for e in tri:
                d, axis = EDGES[e]
                v = (i + d[0]) * strides[0] + (j + d[1]) * strides[1] + (k + d[2])
                assert flags.reshape(-1)[3 * v + axis]
                ids.append(vid[3 * v + axis])
            faces.append(ids)
    faces = np.asarray(faces, np.int32).reshape(-1, 3)
    return verts, faces, normals.astype(np.float32)


def meshify_colored_voxel_grid(colored_voxel_grid, stride=1):
    """voxel_utils.py:53-95 with marching_cubes_binary in place of skimage's marching_cubes (see the module header)."""
    from sklearn.neighbors import NearestNeighbors
    grid = colored_voxel_grid[::stride, ::stride, ::stride] if stride > 1 else colored_voxel_grid          # :60-63
    voxel_mask = np.any(grid > 0, axis=-1)                                                                  # :66
    verts, faces, normals = marching_cubes_binary(voxel_mask.astype(np.uint8))                              # :69-72
    verts = verts * stride                                                                                  # :75
    verts = verts[:, [2, 1, 0]]                                                                             # :78
    verts[:, 2] = colored_voxel_grid.shape[2] - verts[:, 2]                                                 # :82
    filled_coords = np.argwhere(voxel_mask)                                                                 # :85
    filled_colors = grid[voxel_mask]                                                                        # :86
    nbrs = NearestNeighbors(n_neighbors=1).fit(filled_coords)                                               # :88
    _, idx = nbrs.kneighbors(verts[:, [2, 1, 0]] / stride)                                                  # :89
    vertex_colors = filled_colors[idx[:, 0]]                                                                # :90
    if vertex_colors.max() > 1:                                                                             # :92-93
        vertex_colors = vertex_colors / 255.0
    return verts, faces, vertex_colors, normals
