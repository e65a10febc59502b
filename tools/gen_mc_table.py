"""Generate csrc/p3d_mc_table.inc: the 256-case triangle table of the marching-cubes kernel (p3d_mesh.cu).

    python tools/gen_mc_table.py > part-based-3d-reconstruction_b200/csrc/p3d_mc_table.inc

Conventions (shared with the kernel and restated independently in oracle/mesh_oracle.py):
  corner c = 4*d0 + 2*d1 + d2, d_k in {0,1} the offset along volume axis k; case = sum(occupied(corner c) << c)
  edge   e = 4*axis + 2*u + w: the edge along `axis` whose lower corner has offsets (u, w) on the two other axes (in
             rising axis order); its vertex is owned by the voxel at that lower corner
  ambiguous faces (two diagonal corners occupied) SEPARATE the occupied corners: occupancy is 6-connected, like the
             scipy.ndimage.label calls of the carving stage
  orientation: counter-clockwise seen from the empty side (normals point from occupied to empty)
Every case is a set of closed loops over the cut edges (one segment per face crossing pair), loops in order of their
lowest edge id, each fan-triangulated from the first vertex (from the lowest edge id on, in loop order) whose fan has no
triangle lying inside a cube face.
"""
import itertools
import sys


def corner_offsets(c):
    return ((c >> 2) & 1, (c >> 1) & 1, c & 1)


def corner_index(d):
    return 4 * d[0] + 2 * d[1] + d[2]


def edge_table():
    """edge id -> (lower corner offsets, axis)."""
    edges = []
    for axis in range(3):
        others = [a for a in range(3) if a != axis]
        for u in range(2):
            for w in range(2):
                d = [0, 0, 0]
                d[others[0]], d[others[1]] = u, w
                edges.append((tuple(d), axis))
    return edges


EDGES = edge_table()


def edge_between(ca, cb):
    da, db = corner_offsets(ca), corner_offsets(cb)
    diff = [k for k in range(3) if da[k] != db[k]]
    assert len(diff) == 1
    low = da if da[diff[0]] == 0 else db
    return EDGES.index((low, diff[0]))


def edge_midpoint(e):
    d, axis = EDGES[e]
    p = [float(v) for v in d]
    p[axis] += 0.5
    return p


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def sub(a, b):
    return tuple(x - y for x, y in zip(a, b))


def faces():
    """(outward normal, the 4 corners in cyclic order) of the 6 cube faces."""
    out = []
    for axis in range(3):
        others = [a for a in range(3) if a != axis]
        for side in range(2):
            cyc = []
            for u, w in ((0, 0), (1, 0), (1, 1), (0, 1)):
                d = [0, 0, 0]
                d[axis], d[others[0]], d[others[1]] = side, u, w
                cyc.append(corner_index(d))
            n = [0.0, 0.0, 0.0]
            n[axis] = 1.0 if side else -1.0
            out.append((tuple(n), cyc))
    return out


FACES = faces()


def directed(p_edge, q_edge, inside_corner, normal):
    """Order the segment's two edges so that, seen from outside the cube, the occupied corner lies on its left."""
    P, Q, C = edge_midpoint(p_edge), edge_midpoint(q_edge), [float(v) for v in corner_offsets(inside_corner)]
    s = dot(cross(sub(Q, P), sub(C, P)), normal)
    assert s != 0
    return (p_edge, q_edge) if s > 0 else (q_edge, p_edge)


def case_segments(case):
    occ = [(case >> c) & 1 for c in range(8)]
    segs = []
    for normal, cyc in FACES:
        o = [occ[c] for c in cyc]
        n_in = sum(o)
        if n_in in (0, 4):
            continue
        cut = [k for k in range(4) if o[k] != o[(k + 1) % 4]]            # edge k joins cyc[k] and cyc[k+1]
        e_of = lambda k: edge_between(cyc[k], cyc[(k + 1) % 4])
        if len(cut) == 2:
            inside = next(cyc[k] for k in range(4) if o[k])
            segs.append(directed(e_of(cut[0]), e_of(cut[1]), inside, normal))
        else:                                                            # two diagonal corners occupied: cut each one off
            assert len(cut) == 4 and n_in == 2
            for k in range(4):
                if o[k]:
                    segs.append(directed(e_of((k - 1) % 4), e_of(k), cyc[k], normal))
    return segs


def case_triangles(case, flip):
    nxt = {}
    for a, b in case_segments(case):
        assert a not in nxt
        nxt[a] = b
    tris, left = [], set(nxt)
    while left:
        start = min(left)
        loop, e = [], start
        while True:
            loop.append(e)
            left.discard(e)
            e = nxt[e]
            if e == start:
                break
        for t in fan(loop):
            tris.append((t[0], t[2], t[1]) if flip else t)
    return tris


def edge_faces(e):
    d, axis = EDGES[e]
    return {(a, d[a]) for a in range(3) if a != axis}


def fan(loop):
    """Fan triangulation of a loop from the first of its vertices (in loop order, starting at the lowest edge id) for which
    no triangle has all three vertices on one cube face: such a triangle would lie IN the face, and the neighbouring cell
    would add its mirror image -- a zero-thickness fin."""
    k = len(loop)
    for r in range(k):
        rot = loop[r:] + loop[:r]
        tris = [(rot[0], rot[i], rot[i + 1]) for i in range(1, k - 1)]
        if not any(edge_faces(a) & edge_faces(b) & edge_faces(c) for a, b, c in tris):
            return tris
    raise AssertionError(f"no fin-free fan for loop {loop}")


def needs_flip():
    """Corner 0 alone occupied: the one triangle must face away from it (normal . (1,1,1) > 0)."""
    (a, b, c), = case_triangles(1, False)
    A, B, C = edge_midpoint(a), edge_midpoint(b), edge_midpoint(c)
    return dot(cross(sub(B, A), sub(C, A)), (1.0, 1.0, 1.0)) < 0


def table():
    flip = needs_flip()
    return [case_triangles(case, flip) for case in range(256)]


def main():
    tab = table()
    width = max(len(t) for t in tab)
    print("// generated by tools/gen_mc_table.py -- do not edit.  Marching-cubes triangle table: corner c = 4*d0 + 2*d1 + d2,")
    print("// edge e = 4*axis + 2*u + w, ambiguous faces separate the occupied corners, counter-clockwise seen from the empty side.")
    print(f"constexpr int kMcMaxTris = {width};")
    print("__device__ const unsigned char kMcEdgeInfo[12][4] = {   // lower-corner offsets (d0, d1, d2), axis")
    for d, axis in EDGES:
        print(f"    {{{d[0]}, {d[1]}, {d[2]}, {axis}}},")
    print("};")
    print("__device__ const unsigned char kMcTriCount[256] = {")
    for r in range(0, 256, 32):
        print("    " + ", ".join(str(len(t)) for t in tab[r:r + 32]) + ",")
    print("};")
    print(f"__device__ const signed char kMcTris[256][{3 * width}] = {{")
    for t in tab:
        flat = list(itertools.chain.from_iterable(t)) + [-1] * (3 * (width - len(t)))
        print("    {" + ", ".join(str(v) for v in flat) + "},")
    print("};")


if __name__ == "__main__":
    main()
