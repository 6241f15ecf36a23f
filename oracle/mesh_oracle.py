"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the mesh post-ops that follow marching tetrahedra in the reference's `getMesh`
(nvdiffrec/lib/geometry/dmtet.py:283-289):

  * auto_normals      nvdiffrec/lib/render/mesh.py:200-227  (dot / safe_normalize: render/util.py:20-35)
  * compute_tangents  nvdiffrec/lib/render/mesh.py:233-277
  * write_obj         nvdiffrec/lib/render/obj.py:165-216   (positions + faces only, as the reference writes them)

Pinned against the reference functions (exec'd from their source, device pin removed) by oracle/make_golden.py.
"""
import numpy as np


def _safe_normalize(x, eps=1e-20):
    d = np.sum(x * x, -1, keepdims=True, dtype=np.float32)
    return x / np.sqrt(np.maximum(d, np.float32(eps)))


def auto_normals(v_pos, t_pos_idx):
    v = np.asarray(v_pos, np.float32)
    f = np.asarray(t_pos_idx, np.int64)
    v0, v1, v2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    fn = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    acc = np.zeros(v.shape, np.float64)  # the reference accumulates in fp32; order effects are below the test tolerance
    for k in range(3):
        np.add.at(acc, f[:, k], fn)
    vn = acc.astype(np.float32)
    d = np.sum(vn * vn, -1, keepdims=True)
    vn = np.where(d > 1e-20, vn, np.array([0.0, 0.0, 1.0], np.float32))
    return _safe_normalize(vn).astype(np.float32), fn


def compute_tangents(v_pos, t_pos_idx, v_tex, t_tex_idx, v_nrm, t_nrm_idx):
    v = np.asarray(v_pos, np.float32); uv = np.asarray(v_tex, np.float32); n = np.asarray(v_nrm, np.float32)
    tp, tt, tn = (np.asarray(a, np.int64) for a in (t_pos_idx, t_tex_idx, t_nrm_idx))
    pos = [v[tp[:, i]] for i in range(3)]
    tex = [uv[tt[:, i]] for i in range(3)]
    uve1, uve2 = tex[1] - tex[0], tex[2] - tex[0]
    pe1, pe2 = pos[1] - pos[0], pos[2] - pos[0]
    nom = pe1 * uve2[:, 1:2] - pe2 * uve1[:, 1:2]
    denom = uve1[:, 0:1] * uve2[:, 1:2] - uve1[:, 1:2] * uve2[:, 0:1]
    tang = nom / np.where(denom > 0.0, np.maximum(denom, np.float32(1e-6)), np.minimum(denom, np.float32(-1e-6)))
    acc = np.zeros(n.shape, np.float64)
    cnt = np.zeros((n.shape[0], 1), np.float64)
    for i in range(3):
        np.add.at(acc, tn[:, i], tang)
        np.add.at(cnt, tn[:, i], 1.0)
    t = (acc / cnt).astype(np.float32)
    t = _safe_normalize(t)
    t = _safe_normalize(t - np.sum(t * n, -1, keepdims=True) * n)
    return t.astype(np.float32)


def obj_text(v_pos, t_pos_idx):
    """Exactly the characters obj.write_obj emits for a mesh without texture coordinates / normals."""
    v = np.asarray(v_pos)
    f = np.asarray(t_pos_idx)
    out = ["g default\n"]
    for p in v:
        out.append("v {} {} {} \n".format(p[0], p[1], p[2]))
    out += ["s 1 \n", "g pMesh1\n", "usemtl defaultMat\n"]
    for tri in f:
        out.append("f " + "".join(" %s/%s/%s" % (str(tri[j] + 1), "", "") for j in range(3)) + "\n")
    return "".join(out)
