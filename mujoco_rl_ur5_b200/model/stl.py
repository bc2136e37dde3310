"""Binary STL loading, solid mass properties and convex hulls for mesh geoms.

Replaces what MuJoCo's compiler does for `<mesh file=...>` assets (reference scenes:
UR5+gripper/UR5gripper_2_finger.xml:54-71): collision uses the convex hull of the mesh, inertia uses the
mesh volume at the geom density.
"""
import struct

import numpy as np


def load_stl(path):
    """Return triangles as float64 array [ntri, 3, 3]."""
    with open(path, "rb") as f:
        data = f.read()
    (ntri,) = struct.unpack_from("<I", data, 80)
    if 84 + 50 * ntri != len(data):
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                        count=ntri, offset=84)
    return rec["v"].astype(np.float64)


def mesh_mass_properties(tris):
    """Volume, centre of mass and inertia-about-COM (unit density) of the solid bounded by `tris`,
    by signed tetrahedra against the origin (exact for a closed, consistently oriented surface)."""
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))  # 6 * signed tet volume
    vol = det.sum() / 6.0
    if vol < 0:  # inward-facing winding
        det, vol = -det, -vol
    com = ((a + b + c) * det[:, None]).sum(axis=0) / (24.0 * vol)
    # second moments: integral of x_i x_j over each tet = det/120 * (sum over vertex pairs ...)
    S = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            t = (a[:, i] * a[:, j] + b[:, i] * b[:, j] + c[:, i] * c[:, j]) * 2.0
            t += a[:, i] * b[:, j] + b[:, i] * a[:, j] + a[:, i] * c[:, j] + c[:, i] * a[:, j]
            t += b[:, i] * c[:, j] + c[:, i] * b[:, j]
            S[i, j] = (t * det).sum() / 120.0
    I0 = np.trace(S) * np.eye(3) - S  # about the origin
    I = I0 - vol * (com @ com * np.eye(3) - np.outer(com, com))
    return float(vol), com, I


def convex_hull(points):
    """Hull vertices [nh,3] and outward-oriented triangular faces [nf,3] (indices into the hull vertices)."""
    from scipy.spatial import ConvexHull

    pts = np.unique(np.round(points, 9), axis=0)
    hull = ConvexHull(pts)
    vid = hull.vertices
    remap = -np.ones(len(pts), np.int64)
    remap[vid] = np.arange(len(vid))
    faces = remap[hull.simplices]
    hv = pts[vid]
    # orient outward using the facet plane normals qhull returns
    n = hull.equations[:, :3]
    fa, fb, fc = hv[faces[:, 0]], hv[faces[:, 1]], hv[faces[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(fb - fa, fc - fa), n) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    return hv, faces.astype(np.int32)
