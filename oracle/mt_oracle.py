"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the reference's marching tetrahedra, `DMTet.__call__`
(nvdiffrec/lib/geometry/dmtet.py:105-163; tables :34-54, sort_edges :60-68, map_uv :70-99), of the vertex placement
`get_deformed` (dmtet.py:293-304) and of the grid -> tet-vertex gather (nvdiffrec/eval.py:389-419).
Pinned against the reference class itself (exec'd from its source with 'cuda' -> 'cpu') by oracle/make_golden.py.
"""
import numpy as np

TRIANGLE_TABLE = np.array([
    [-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
    [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
    [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
    [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1, -1, -1, -1, -1, -1]], dtype=np.int64)
NUM_TRIANGLES = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=np.int64)
BASE_TET_EDGES = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=np.int64)


def marching_tets(pos, sdf, tets):
    """pos [Nv,3] f32, sdf [Nv] f32, tets [F,4] int -> (verts, faces, uvs, uv_idx, face_to_valid_tet, valid_vert_idx)."""
    pos = np.asarray(pos, np.float32)
    sdf = np.asarray(sdf, np.float32)
    tets = np.asarray(tets, np.int64)
    occ = sdf > 0
    occ4 = occ[tets.reshape(-1)].reshape(-1, 4)
    occ_sum = occ4.sum(-1)
    valid = (occ_sum > 0) & (occ_sum < 4)

    edges = tets[valid][:, BASE_TET_EDGES].reshape(-1, 2)
    edges = np.stack([edges.min(1), edges.max(1)], -1)
    uniq, inverse = np.unique(edges, axis=0, return_inverse=True)  # lexicographically sorted rows
    inverse = inverse.reshape(-1)
    crossing = occ[uniq.reshape(-1)].reshape(-1, 2).sum(-1) == 1
    mapping = -np.ones(uniq.shape[0], np.int64)
    mapping[crossing] = np.arange(crossing.sum(), dtype=np.int64)
    idx_map = mapping[inverse].reshape(-1, 6)

    ev = uniq[crossing]
    p = pos[ev.reshape(-1)].reshape(-1, 2, 3)
    s = sdf[ev.reshape(-1)].reshape(-1, 2, 1).copy()
    s[:, -1] *= -1
    denom = s.sum(1, keepdims=True)
    w = s[:, ::-1] / denom
    verts = (p * w).sum(1).astype(np.float32)

    tetindex = (occ4[valid] * (2 ** np.arange(4, dtype=np.int64))[None]).sum(-1)
    ntri = NUM_TRIANGLES[tetindex]
    one, two = ntri == 1, ntri == 2
    f1 = np.take_along_axis(idx_map[one], TRIANGLE_TABLE[tetindex[one]][:, :3], 1).reshape(-1, 3)
    f2 = np.take_along_axis(idx_map[two], TRIANGLE_TABLE[tetindex[two]][:, :6], 1).reshape(-1, 3)
    faces = np.concatenate([f1, f2], 0)

    F = tets.shape[0]
    gidx = np.arange(F, dtype=np.int64)[valid]
    face_gidx = np.concatenate([gidx[one] * 2, np.stack([gidx[two] * 2, gidx[two] * 2 + 1], -1).reshape(-1)])
    uvs, uv_idx = map_uv(face_gidx, F * 2)
    face_to_valid_tet = np.concatenate([gidx[one], np.stack([gidx[two], gidx[two]], -1).reshape(-1)])
    valid_vert_idx = np.unique(tets[gidx[ntri > 0]])
    return verts, faces, uvs, uv_idx, face_to_valid_tet, valid_vert_idx


def marching_tets_vertex_grad(pos, sdf, tets, grad_verts):
    """What torch autograd returns for d(sum(verts * grad_verts)) / d(pos, sdf) through dmtet.py:125-132 (float64 inside):
    v = (p_a * (-s_b) + p_b * s_a) / (s_a - s_b) on every crossing edge (a < b) of the sorted unique edge list."""
    pos = np.asarray(pos, np.float64)
    sdf = np.asarray(sdf, np.float64)
    tets = np.asarray(tets, np.int64)
    g = np.asarray(grad_verts, np.float64)
    occ = np.asarray(sdf, np.float32) > 0
    occ4 = occ[tets.reshape(-1)].reshape(-1, 4)
    occ_sum = occ4.sum(-1)
    valid = (occ_sum > 0) & (occ_sum < 4)
    edges = tets[valid][:, BASE_TET_EDGES].reshape(-1, 2)
    edges = np.stack([edges.min(1), edges.max(1)], -1)
    uniq = np.unique(edges, axis=0)
    ev = uniq[occ[uniq.reshape(-1)].reshape(-1, 2).sum(-1) == 1]
    a, b = ev[:, 0], ev[:, 1]
    sa, t = sdf[a], -sdf[b]
    den = sa + t
    wa, wb = t / den, sa / den
    gpos = np.zeros_like(pos)
    np.add.at(gpos, a, g * wa[:, None])
    np.add.at(gpos, b, g * wb[:, None])
    dot = (g * (pos[b] - pos[a])).sum(-1)
    gsdf = np.zeros_like(sdf)
    np.add.at(gsdf, a, dot * t / den ** 2)
    np.add.at(gsdf, b, dot * sa / den ** 2)
    return gpos.astype(np.float32), gsdf.astype(np.float32)


def uv_grid_n(max_idx):
    return int(np.ceil(np.sqrt((max_idx + 1) // 2)))


def map_uv(face_gidx, max_idx):
    N = uv_grid_n(max_idx)
    lin = np.linspace(0, 1 - (1 / N), N, dtype=np.float32)
    tex_y, tex_x = np.meshgrid(lin, lin, indexing="ij")
    pad = np.float32(0.9 / N)
    uvs = np.stack([tex_x, tex_y, tex_x + pad, tex_y, tex_x + pad, tex_y + pad, tex_x, tex_y + pad], -1).reshape(-1, 2)
    tet_idx = face_gidx // 2
    tet_idx = (tet_idx // N) * N + tet_idx % N
    tri = face_gidx % 2
    uv_idx = np.stack([tet_idx * 4, tet_idx * 4 + tri + 1, tet_idx * 4 + tri + 2], -1).reshape(-1, 3)
    return uvs.astype(np.float32), uv_idx


def grid_coords_of_tet_vertices(vertices):
    """eval.py:391-397 / evaler.py:187-195: integer grid coordinate of every tet vertex."""
    vertices = np.asarray(vertices, np.float32)
    u = np.unique(vertices)
    dx = u[1] - u[0]
    return np.round((vertices - vertices.min()) / dx).astype(np.int64)


def grid_to_tet_inputs(grid, coords, vertices, grid_res, mesh_scale=1.0, deform_scale=1.0):
    """eval.py:412-419 + dmtet.py:303: (sdf [Nv], deformed positions [Nv,3]) from one sample grid [4,R,R,R]."""
    grid = np.asarray(grid, np.float32)
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    sdf = np.sign(grid[0, x, y, z]).astype(np.float32)
    deform = np.clip(grid[1:, x, y, z].T, -1.0, 1.0).astype(np.float32)
    verts = np.asarray(vertices, np.float32) * np.float32(mesh_scale)
    pos = verts + np.float32(2 / (grid_res * 2)) * deform * np.float32(deform_scale)
    return sdf, pos.astype(np.float32)
