"""Small numpy rigid-body toolbox used by the model compiler (compile-time only, never on the hot path).

`Kinematics` evaluates forward kinematics, body Jacobians and the joint-space mass matrix by the
Jacobian-sum formula  M = sum_b  m Jp^T Jp + Jr^T I Jr  (+ armature).  It is deliberately a different
formulation from the composite-rigid-body recursion used by the oracle and the CUDA kernels, so the
tests can cross-check them against each other.
"""
import numpy as np

J_FREE, J_BALL, J_SLIDE, J_HINGE = 0, 1, 2, 3


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_axis_angle(axis, angle):
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def inertia6_to_mat(v):
    return np.array([[v[0], v[3], v[4]], [v[3], v[1], v[5]], [v[4], v[5], v[2]]])


class Kinematics:
    def __init__(self, M):
        self.M = M

    def forward(self, qpos):
        M = self.M
        nb = int(np.asarray(M["nbody"]).ravel()[0])
        xpos = np.zeros((nb, 3))
        xquat = np.zeros((nb, 4))
        xquat[0] = [1, 0, 0, 0]
        xmat = np.zeros((nb, 3, 3))
        xmat[0] = np.eye(3)
        nj = int(np.asarray(M["njnt"]).ravel()[0])
        xanchor = np.zeros((nj, 3))
        xaxis = np.zeros((nj, 3))
        for b in range(1, nb):
            p = M["body_parentid"][b]
            pos = xpos[p] + xmat[p] @ M["body_pos"][b]
            quat = quat_mul(xquat[p], M["body_quat"][b])
            for k in range(M["body_jntnum"][b]):
                j = M["body_jntadr"][b] + k
                t, qa = M["jnt_type"][j], M["jnt_qposadr"][j]
                if t == J_FREE:
                    pos = qpos[qa:qa + 3].copy()
                    quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                    xanchor[j] = pos
                    continue
                R = quat_to_mat(quat)
                xanchor[j] = pos + R @ M["jnt_pos"][j]
                xaxis[j] = R @ M["jnt_axis"][j]
                if t == J_SLIDE:
                    pos = pos + xaxis[j] * (qpos[qa] - M["qpos0"][qa])
                elif t == J_HINGE:
                    quat = quat_mul(quat, quat_axis_angle(M["jnt_axis"][j], qpos[qa] - M["qpos0"][qa]))
                    pos = xanchor[j] - quat_to_mat(quat) @ M["jnt_pos"][j]
                elif t == J_BALL:
                    quat = quat_mul(quat, qpos[qa:qa + 4] / np.linalg.norm(qpos[qa:qa + 4]))
                    pos = xanchor[j] - quat_to_mat(quat) @ M["jnt_pos"][j]
            xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
            xmat[b] = quat_to_mat(xquat[b])
        xipos = np.array([xpos[b] + xmat[b] @ M["body_ipos"][b] for b in range(nb)])
        return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, xanchor=xanchor, xaxis=xaxis)

    def jacobian(self, st, body, point):
        """3 x nv translational (of world `point` attached to `body`) and rotational Jacobians."""
        M = self.M
        nv = int(np.asarray(M["nv"]).ravel()[0])
        Jp, Jr = np.zeros((3, nv)), np.zeros((3, nv))
        d = M["body_lastdof"][body]
        while d >= 0:
            j = M["dof_jntid"][d]
            t = M["jnt_type"][j]
            k = d - M["jnt_dofadr"][j]
            b = M["dof_bodyid"][d]
            if t == J_SLIDE:
                Jp[:, d] = st["xaxis"][j]
            elif t == J_HINGE:
                Jr[:, d] = st["xaxis"][j]
                Jp[:, d] = np.cross(st["xaxis"][j], point - st["xanchor"][j])
            elif t == J_BALL:
                ax = st["xmat"][b][:, k]
                Jr[:, d] = ax
                Jp[:, d] = np.cross(ax, point - st["xanchor"][j])
            elif t == J_FREE:
                if k < 3:
                    Jp[k, d] = 1.0
                else:
                    ax = st["xmat"][b][:, k - 3]
                    Jr[:, d] = ax
                    Jp[:, d] = np.cross(ax, point - st["xpos"][b])
            d = M["dof_parentid"][d]
        return Jp, Jr

    def mass_matrix(self, qpos):
        M = self.M
        nv = int(np.asarray(M["nv"]).ravel()[0])
        st = self.forward(qpos)
        H = np.diag(np.asarray(M["dof_armature"], dtype=np.float64).copy()) if nv else np.zeros((0, 0))
        for b in range(1, int(np.asarray(M["nbody"]).ravel()[0])):
            if M["body_lastdof"][b] < 0 or M["body_mass"][b] <= 0:
                continue
            Jp, Jr = self.jacobian(st, b, st["xipos"][b])
            Iw = st["xmat"][b] @ inertia6_to_mat(M["body_inertia"][b]) @ st["xmat"][b].T
            H += M["body_mass"][b] * Jp.T @ Jp + Jr.T @ Iw @ Jr
        return H
