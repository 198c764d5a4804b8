"""Oracle restatement (numpy, CPU) of the iso-surface extraction behind ``utils/mesh_util.py`` (TEST INFRASTRUCTURE ONLY).

The reference calls ``skimage.measure.marching_cubes`` (Lewiner's variant, utils/mesh_util.py:33-35) and writes the mesh
with ``plyfile`` (:57-72).  Neither package is in this image and the reference pins no version of either (docs/usage.md:42), so
**parity with skimage's triangulation is unpinned**: this module restates the algorithm the product implements (classic marching
cubes with the face-consistent case table of neurecon_b200/mc_tables.py), vectorised, in the same un-fused fp32 arithmetic, so
that csrc/marching_cubes.cu can be checked bit for bit; what ties it to the reference are mesh-level properties any correct
marching-cubes output shares (vertices = the linear edge crossings, closed surface, area / volume, orientation convention,
``spacing`` and origin handling of mesh_util.py:33-42) and the PLY byte layout of plyfile (``read_ply`` below parses it back).
"""
import numpy as np

from neurecon_b200 import mc_tables


def marching_cubes(vol, level=0.0, spacing=(1.0, 1.0, 1.0), gradient_direction="descent"):
    """vol [Nx, Ny, Nz] float32 -> (verts [V, 3] float32 in units of ``spacing``, faces [F, 3] int32).
    Vertex order: (owner lattice point, axis); face order: (cell, table order)."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    Nx, Ny, Nz = vol.shape
    level = np.float32(level)
    inside = vol < level
    act = np.zeros((3,) + vol.shape, dtype=bool)
    act[0, :-1] = inside[:-1] != inside[1:]
    act[1, :, :-1] = inside[:, :-1] != inside[:, 1:]
    act[2, :, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    cnt = act.sum(0).reshape(-1)
    vbase = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    rank = np.stack([np.zeros_like(act[0], dtype=np.int64), act[0].astype(np.int64), act[0].astype(np.int64) + act[1]], 0)
    vid = vbase.reshape(vol.shape)[None] + rank                       # vertex id of (axis, point) where act
    V = int(cnt.sum())
    verts = np.zeros((V, 3), dtype=np.float32)
    sp = np.asarray(spacing, dtype=np.float32)
    shifted = [vol[1:], vol[:, 1:], vol[:, :, 1:]]
    for a in range(3):
        idx = np.nonzero(act[a])
        v0 = vol[idx]
        sl = [slice(None)] * 3
        sl[a] = slice(0, vol.shape[a] - 1)
        v1 = shifted[a][tuple(np.asarray(ix) for ix in idx)]
        t = ((level - v0).astype(np.float32) / (v1 - v0).astype(np.float32)).astype(np.float32)
        ids = vid[a][idx]
        for c in range(3):
            coord = idx[c].astype(np.float32)
            if c == a:
                coord = (coord + t).astype(np.float32)
            verts[ids, c] = (coord * sp[c]).astype(np.float32)
    tri, ntri = mc_tables.tables()
    case = np.zeros((Nx - 1, Ny - 1, Nz - 1), dtype=np.int64)
    for c in range(8):
        dx, dy, dz = mc_tables.corner_offset(c)
        case |= inside[dx:Nx - 1 + dx, dy:Ny - 1 + dy, dz:Nz - 1 + dz].astype(np.int64) << c
    nt = ntri[case].astype(np.int64).reshape(-1)
    fbase = np.concatenate([[0], np.cumsum(nt)[:-1]])
    F = int(nt.sum())
    faces = np.zeros((F, 3), dtype=np.int32)
    ci, cj, ck = np.meshgrid(np.arange(Nx - 1), np.arange(Ny - 1), np.arange(Nz - 1), indexing="ij")
    ci, cj, ck, cflat = ci.reshape(-1), cj.reshape(-1), ck.reshape(-1), case.reshape(-1)
    for t in range(mc_tables.MAX_TRIS):
        m = nt > t
        if not m.any():
            break
        pi, pj, pk, cs = ci[m], cj[m], ck[m], cflat[m]
        for c in range(3):
            e = tri[cs, 3 * t + c].astype(np.int64)
            a, b0, b1 = e >> 2, e & 1, (e >> 1) & 1
            q = [pi.copy(), pj.copy(), pk.copy()]
            for axis in range(3):
                others = [x for x in range(3) if x != axis]
                sel = a == axis
                q[others[0]][sel] += b0[sel]
                q[others[1]][sel] += b1[sel]
            faces[fbase[m] + t, c] = vid[a, q[0], q[1], q[2]]
    if gradient_direction == "ascent":
        faces = faces[:, [0, 2, 1]]
    elif gradient_direction != "descent":
        raise ValueError("gradient_direction must be 'descent' or 'ascent'")
    return verts, np.ascontiguousarray(faces)


def mesh_stats(verts, faces):
    """(closed: every edge in exactly two faces with opposite directions, Euler characteristic, area, signed volume)"""
    v = verts.astype(np.float64)
    f = faces.astype(np.int64)
    he = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    key = np.sort(he, 1)
    uniq, inv, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    sign = np.where(he[:, 0] < he[:, 1], 1, -1)
    balance = np.bincount(inv.reshape(-1), weights=sign, minlength=len(uniq))
    closed = bool((counts == 2).all() and (balance == 0).all())
    used = np.unique(f)
    euler = len(used) - len(uniq) + len(f)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cr = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(cr, axis=1).sum()
    volume = (a * cr).sum() / 6.0
    return closed, int(euler), float(area), float(volume)


def read_ply(path):
    """Parse the binary little-endian PLY that plyfile writes for the reference (mesh_util.py:57-72): vertex x y z float32,
    face = uchar count + int32 indices."""
    with open(path, "rb") as fh:
        header = []
        while True:
            line = fh.readline().decode("ascii").rstrip("\n")
            header.append(line)
            if line == "end_header":
                break
        nv = int([l for l in header if l.startswith("element vertex")][0].split()[-1])
        nf = int([l for l in header if l.startswith("element face")][0].split()[-1])
        verts = np.frombuffer(fh.read(12 * nv), dtype="<f4").reshape(nv, 3)
        rec = np.frombuffer(fh.read(13 * nf), dtype=np.dtype([("n", "u1"), ("idx", "<i4", (3,))]))
        assert fh.read() == b""
    return header, verts, rec["n"], rec["idx"]
