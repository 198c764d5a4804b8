"""Marching-cubes case table for csrc/marching_cubes.cu and oracle/mesh.py, GENERATED (not transcribed) from the rule it
implements, so that the rule is the specification:

* cube corner c in 0..7 sits at offset (c & 1, (c >> 1) & 1, (c >> 2) & 1) = (dx, dy, dz) from the cell origin, the x axis
  being the slowest array axis of the volume [Nx, Ny, Nz];
* cube edge e = 4 a + b0 + 2 b1 runs along axis a between the corners whose other two coordinates (in increasing axis
  order) are (b0, b1);
* a corner is INSIDE when value < level; an edge carries a vertex when exactly one endpoint is inside;
* on every cube face the crossing points are joined in pairs; a face with four of them (the two inside corners on a
  diagonal) joins them so that each inside corner is cut off on its own.  The rule looks only at the face's four corners,
  so the two cells sharing a face always agree and the mesh has no cracks (the classic 15-case table completed by
  complement symmetry does not have this property);
* the segments of the six faces close into loops; each loop becomes a triangle fan;
* orientation: the right-hand normal of every triangle points towards DECREASING values (skimage's default
  gradient_direction='descent'); the kernels flip the winding for 'ascent'.

`tables()` returns (tri_table int8 [256, 32]: up to 10 triangles as edge-id triples, -1 terminated; n_tris uint8 [256])."""
import itertools

import numpy as np

MAX_TRIS = 10
_CACHE = None


def corner_offset(c):
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1)


def edge_id(axis, b0, b1):
    return 4 * axis + b0 + 2 * b1


def edge_corners(e):
    """(corner at coordinate 0 of the edge's axis, corner at coordinate 1)"""
    axis, b = divmod(e, 4)
    others = [a for a in range(3) if a != axis]
    off = [0, 0, 0]
    off[others[0]], off[others[1]] = b & 1, b >> 1
    c0 = off[0] + 2 * off[1] + 4 * off[2]
    off[axis] = 1
    c1 = off[0] + 2 * off[1] + 4 * off[2]
    return c0, c1


_EDGE_OF = {frozenset(edge_corners(e)): e for e in range(12)}


def _edge_between(ca, cb):
    return _EDGE_OF[frozenset((ca, cb))]


def _faces():
    """the 6 faces as cyclic corner quadruples"""
    out = []
    for axis, side in itertools.product(range(3), (0, 1)):
        others = [a for a in range(3) if a != axis]
        quad = []
        for u, v in ((0, 0), (1, 0), (1, 1), (0, 1)):
            off = [0, 0, 0]
            off[axis], off[others[0]], off[others[1]] = side, u, v
            quad.append(off[0] + 2 * off[1] + 4 * off[2])
        out.append(quad)
    return out


def _case(index, faces):
    inside = [(index >> c) & 1 for c in range(8)]          # bit c of the case index: corner c inside
    adj = {}

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)

    for quad in faces:
        fe = [_edge_between(quad[i], quad[(i + 1) % 4]) for i in range(4)]      # edge i joins quad[i], quad[i + 1]
        cross = [i for i in range(4) if inside[quad[i]] != inside[quad[(i + 1) % 4]]]
        if len(cross) == 2:
            link(fe[cross[0]], fe[cross[1]])
        elif len(cross) == 4:
            for i in range(4):                           # cut off each inside corner: its two face edges are i - 1 and i
                if inside[quad[i]]:
                    link(fe[(i - 1) % 4], fe[i])
    loops, seen = [], set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            assert len(adj[cur]) == 2, (index, cur, adj[cur])
            nxt = adj[cur][0] if adj[cur][0] != prev else adj[cur][1]
            if adj[cur][0] == adj[cur][1]:
                nxt = adj[cur][0]
            if nxt == start:
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        loops.append(loop)
    tris = []
    for loop in loops:
        assert len(loop) >= 3
        mid, across = [], np.zeros(3)
        for e in loop:
            c0, c1 = edge_corners(e)
            p0, p1 = np.array(corner_offset(c0), float), np.array(corner_offset(c1), float)
            mid.append(0.5 * (p0 + p1))
            across += (p1 - p0) if inside[c0] else (p0 - p1)       # from the inside endpoint to the outside one
        normal = np.zeros(3)
        for i in range(len(loop)):
            normal += np.cross(mid[i], mid[(i + 1) % len(loop)])   # Newell
        s = float(normal @ across)
        assert abs(s) > 1e-9, (index, loop)
        if s > 0:                                        # normal towards increasing values: reverse ('descent' convention)
            loop = loop[::-1]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def tables():
    global _CACHE
    if _CACHE is None:
        faces = _faces()
        tri = -np.ones((256, 32), dtype=np.int8)
        cnt = np.zeros(256, dtype=np.uint8)
        for index in range(256):
            t = _case(index, faces)
            assert len(t) <= MAX_TRIS
            cnt[index] = len(t)
            if t:
                tri[index, : 3 * len(t)] = np.asarray(t, dtype=np.int8).reshape(-1)
        _CACHE = (tri, cnt)
    return _CACHE
