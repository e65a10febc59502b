"""meshify_colored_voxel_grid (reference utils/voxel_utils.py:53-95).  The marching-cubes part cannot be pinned on the
reference's output (scikit-image is not available; oracle/mesh_oracle.py says PARITY UNPINNED), so the oracle is pinned
on closed-form properties here and the CUDA path must equal the oracle bit for bit."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, pkg


@pytest.fixture(scope="module")
def mo():
    from oracle import mesh_oracle
    return mesh_oracle


def _edge_counts(faces):
    """(directed edge -> count, undirected edge -> count) as dicts of int tuples."""
    from collections import Counter
    f = np.asarray(faces, np.int64)
    d = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    directed = Counter(map(tuple, d.tolist()))
    undirected = Counter(map(tuple, np.sort(d, axis=1).tolist()))
    return directed, undirected


def _watertight(faces):
    directed, undirected = _edge_counts(faces)
    return all(k == 2 for k in undirected.values()) and all(directed.get((b, a), 0) == 1 for (a, b) in directed)


def _volume(verts, faces):
    p = np.asarray(verts, np.float64)
    a, b, c = p[faces[:, 0]], p[faces[:, 1]], p[faces[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def _mesh_components(n_verts, faces):
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    f = np.asarray(faces, np.int64)
    rows = np.concatenate([f[:, 0], f[:, 1]])
    cols = np.concatenate([f[:, 1], f[:, 2]])
    g = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n_verts, n_verts))
    return cg.connected_components(g, directed=False)[0]


def _padded_random(rng, shape, p):
    m = np.zeros(shape, np.uint8)
    m[1:-1, 1:-1, 1:-1] = rng.random(tuple(s - 2 for s in shape)) < p
    return m


def test_shipped_table_equals_the_oracles(mo):
    """csrc/p3d_mc_table.inc (tools/gen_mc_table.py) against the table the oracle derives from geometry on its own."""
    src = open(os.path.join(ROOT, "part-based-3d-reconstruction_b200", "csrc", "p3d_mc_table.inc")).read()
    width = int(re.search(r"kMcMaxTris = (\d+);", src).group(1))
    info = re.search(r"kMcEdgeInfo\[12\]\[4\] = \{(.*?)\n\};", src, re.S).group(1)
    edges = [tuple(int(v) for v in row) for row in re.findall(r"\{(\d+), (\d+), (\d+), (\d+)\}", info)]
    assert edges == [d + (axis,) for d, axis in mo.EDGES]
    counts = [int(v) for v in re.search(r"kMcTriCount\[256\] = \{(.*?)\};", src, re.S).group(1).replace("\n", " ").split(",") if v.strip()]
    rows = re.search(r"kMcTris\[256\]\[\d+\] = \{(.*)\n\};", src, re.S).group(1)
    tris = [[int(v) for v in row.split(",")] for row in re.findall(r"\{([-\d, ]+)\}", rows)]
    tab = mo.mc_table()
    assert len(counts) == 256 and len(tris) == 256 and width == max(len(t) for t in tab) == 5
    for case in range(256):
        flat = [e for t in tab[case] for e in t]
        assert counts[case] == len(tab[case]) and tris[case] == flat + [-1] * (3 * width - len(flat)), case
    assert not tab[0] and not tab[255]
    for case in range(256):                                   # complementary cases cut the same edges
        assert {e for t in tab[case] for e in t} == {e for t in tab[255 - case] for e in t}


def test_oracle_closed_forms(mo):
    m = np.zeros((3, 3, 3), np.uint8)
    m[1, 1, 1] = 1                                            # one voxel: the octahedron of its six edge midpoints
    v, f, n = mo.marching_cubes_binary(m)
    assert len(v) == 6 and len(f) == 8 and _watertight(f) and abs(_volume(v, f) - 1 / 6) < 1e-12
    assert np.array_equal(n, np.sign(v - 1.0))                # normals = the axis directions away from the voxel
    for k in (2, 3, 5):                                       # k^3 block: cube + face slabs + edge prisms + corner tetrahedra
        m = np.zeros((k + 2,) * 3, np.uint8)
        m[1:-1, 1:-1, 1:-1] = 1
        v, f, n = mo.marching_cubes_binary(m)
        want = (k - 1) ** 3 + 6 * (k - 1) ** 2 * 0.5 + 12 * (k - 1) * 0.125 + 8 / 48
        _, und = _edge_counts(f)
        assert _watertight(f) and abs(_volume(v, f) - want) < 1e-9 and len(v) - len(und) + len(f) == 2
        fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        centre = (k + 1) / 2
        assert np.all(np.einsum("ij,ij->i", fn, v[f].mean(1) - centre) > 0)         # faces look away from the block
        assert np.all(np.einsum("ij,ij->i", n, v - centre) > 0) and np.allclose(np.linalg.norm(n, axis=1), 1)


def test_oracle_on_random_volumes(mo):
    """Closed surfaces: every edge in exactly two triangles with opposite directions; one surface per 6-connected
    occupied component plus one per enclosed 26-connected empty component; vertices = all occupancy-changing edges."""
    import scipy.ndimage as ndi
    rng = np.random.default_rng(11)
    for t in range(12):
        shape = tuple(int(s) for s in rng.integers(4, 12, 3))
        m = _padded_random(rng, shape, rng.uniform(0.15, 0.85))
        v, f, n = mo.marching_cubes_binary(m)
        assert _watertight(f) and _volume(v, f) > 0
        fg = ndi.label(m)[1]
        bg = ndi.label(1 - m, structure=np.ones((3, 3, 3)))[1]
        assert _mesh_components(len(v), f) == fg + bg - 1
        want = []
        for axis in range(3):
            idx = np.argwhere(np.diff(m.astype(np.int8), axis=axis) != 0).astype(np.float32)
            idx[:, axis] += 0.5
            want.append(idx)
        want = np.concatenate(want)
        assert sorted(map(tuple, v.tolist())) == sorted(map(tuple, want.tolist()))
        assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-6)
    # a volume that touches the border stays open there (no padding, as skimage), degenerate shapes give no faces
    m = np.ones((4, 4, 4), np.uint8)
    m[2:, :, :] = 0
    v, f, n = mo.marching_cubes_binary(m)
    assert len(v) == 16 and len(f) == 18 and not _watertight(f) and np.array_equal(n, np.tile([1, 0, 0], (16, 1)))
    v, f, n = mo.marching_cubes_binary(np.array([[[1, 0, 1, 1]]], np.uint8))
    assert len(v) == 2 and len(f) == 0


@pytest.mark.gpu
def test_cuda_marching_cubes_equals_the_oracle(mo):
    vu = pkg("utils.voxel_utils")
    rng = np.random.default_rng(5)
    shapes = [(3, 3, 3), (2, 2, 2), (1, 5, 4), (5, 1, 1), (9, 8, 10), (17, 16, 33), (40, 7, 23), (6, 300, 5)]
    for shape in shapes:
        for p in (0.1, 0.5, 0.9):
            m = (rng.random(shape) < p).astype(np.uint8)
            v, f, n = (a.cpu().numpy() for a in vu.marching_cubes_binary(m))
            wv, wf, wn = mo.marching_cubes_binary(m)
            assert v.dtype == np.float32 and f.dtype == np.int32 and n.dtype == np.float32
            assert np.array_equal(v, wv) and np.array_equal(f, wf) and np.array_equal(n, wn), (shape, p)
    for m in (np.zeros((4, 4, 4), np.uint8), np.ones((4, 4, 4), np.uint8)):
        v, f, n = vu.marching_cubes_binary(m)
        assert v.shape == (0, 3) and f.shape == (0, 3)


@pytest.mark.gpu
def test_meshify_matches_the_oracle_and_is_closed(mo, carve_golden):
    """The whole function on a carved monument (Bibi@64 partwise grid, live-reference golden): vertices after the
    reference's stride / axis swap / mirror, faces, the sklearn nearest-voxel colours and the normals equal the oracle's;
    the surface of the zero-padded grid is closed."""
    vu = pkg("utils.voxel_utils")
    grid = carve_golden["real_Bibi_64_partwise"]
    for stride in (1, 2):
        v, f, c, n = vu.meshify_colored_voxel_grid(grid, stride=stride)
        wv, wf, wc, wn = mo.meshify_colored_voxel_grid(grid, stride=stride)
        assert np.array_equal(v, wv) and np.array_equal(f, wf) and np.array_equal(n, wn)
        assert c.dtype == np.float64 and np.array_equal(c, wc) and c.max() <= 1.0
        assert v.dtype == np.float32 and v[:, 2].max() <= grid.shape[2]
    padded = np.pad(grid, ((1, 1), (1, 1), (1, 1), (0, 0)))
    v, f, c, n = vu.meshify_colored_voxel_grid(padded)
    assert _watertight(f) and len(v) > 1000
    with pytest.raises(ValueError):
        vu.meshify_colored_voxel_grid(np.zeros((4, 4, 4, 3), np.uint8))
