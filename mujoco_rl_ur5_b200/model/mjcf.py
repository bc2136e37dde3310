"""MJCF subset compiler: scene XML + binary STL meshes -> flat model arrays.

Stands in for `mujoco_py.load_model_from_path` (reference: MujocoController.py:33; the scenes are
UR5+gripper/UR5gripper_2_finger.xml and UR5gripper_2_finger_many_objects.xml + objects.xml).
Supported subset = exactly what those two scenes use: compiler(angle=radian, inertiafromgeom, meshdir),
option, default classes, include, body tree, hinge/slide/ball/free joints, plane/box/sphere/capsule/
cylinder/mesh geoms, torque motors, joint equality, contact excludes, fixed cameras, visual/map.

Everything MuJoCo's own compiler derives and the XML does not state (geom-derived inertia at density
1000, mesh volume integrals, convex hulls, invweight0, mean inertia) is re-derived here from the
MuJoCo documentation; see DESIGN.md "model compiler decisions".
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import rigid
from .stl import load_stl, mesh_mass_properties, convex_hull

# geom type codes (same numbering as MuJoCo's mjtGeom, for readability)
G_PLANE, G_SPHERE, G_CAPSULE, G_CYLINDER, G_BOX, G_MESH = 0, 2, 3, 5, 6, 7
GEOM_TYPES = {"plane": G_PLANE, "sphere": G_SPHERE, "capsule": G_CAPSULE, "cylinder": G_CYLINDER,
              "box": G_BOX, "mesh": G_MESH}
# joint type codes (mjtJoint)
J_FREE, J_BALL, J_SLIDE, J_HINGE = 0, 1, 2, 3
JOINT_TYPES = {"free": J_FREE, "ball": J_BALL, "slide": J_SLIDE, "hinge": J_HINGE}

# MuJoCo documented defaults (XML reference chapter)
DEF_GEOM = dict(type="sphere", size="0 0 0", pos="0 0 0", friction="1 0.005 0.0001", condim="3", margin="0",
                gap="0", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", density="1000", contype="1",
                conaffinity="1", rgba="0.5 0.5 0.5 1")
DEF_JOINT = dict(type="hinge", pos="0 0 0", axis="0 0 1", limited="false", range="0 0", damping="0",
                 armature="0", stiffness="0", ref="0", margin="0", solreflimit="0.02 1",
                 solimplimit="0.9 0.95 0.001 0.5 2")
DEF_MOTOR = dict(gear="1", ctrllimited="false", ctrlrange="0 0")
DEF_EQ = dict(solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", polycoef="0 1 0 0 0")


def _f(s):
    return np.array([float(x) for x in s.split()], dtype=np.float64)


def _solimp5(v):
    v = list(_f(v) if isinstance(v, str) else v)
    dflt = [0.9, 0.95, 0.001, 0.5, 2.0]
    return np.array(v + dflt[len(v):], dtype=np.float64)


def _orientation(attrib):
    """quat (w,x,y,z) from quat / euler (intrinsic xyz, radians) / axisangle attributes."""
    if "quat" in attrib:
        q = _f(attrib["quat"])
        return q / np.linalg.norm(q)
    if "euler" in attrib:
        e = _f(attrib["euler"])
        q = np.array([1.0, 0, 0, 0])
        for i in range(3):  # eulerseq "xyz": R = Rx * Ry * Rz (rotating frame)
            ax = np.zeros(3)
            ax[i] = 1.0
            q = rigid.quat_mul(q, rigid.quat_axis_angle(ax, e[i]))
        return q
    if "axisangle" in attrib:
        a = _f(attrib["axisangle"])
        n = np.linalg.norm(a[:3])
        return rigid.quat_axis_angle(a[:3] / n, a[3])
    return np.array([1.0, 0, 0, 0])


class _Defaults:
    """Nested <default> classes; lookup(tag, cls) returns the merged attribute dict."""

    def __init__(self, root):
        self.cls = {}
        top = root.find("default")
        self._walk(top, None, "main") if top is not None else None
        self.cls.setdefault("main", {})

    def _walk(self, node, parent, name):
        d = {k: dict(v) for k, v in self.cls.get(parent, {}).items()} if parent else {}
        for ch in node:
            if ch.tag != "default":
                d.setdefault(ch.tag, {}).update(ch.attrib)
        self.cls[name] = d
        for ch in node:
            if ch.tag == "default":
                self._walk(ch, name, ch.attrib["class"])

    def lookup(self, tag, cls):
        return dict(self.cls.get(cls or "main", self.cls["main"]).get(tag, {}))


def _expand_includes(node, base_dir):
    for i, ch in enumerate(list(node)):
        if ch.tag == "include":
            sub = ET.parse(os.path.join(base_dir, ch.attrib["file"])).getroot()
            idx = list(node).index(ch)
            node.remove(ch)
            for k, s in enumerate(list(sub)):
                node.insert(idx + k, s)
        else:
            _expand_includes(ch, base_dir)


def compile_mjcf(path, transform=None):
    """Parse `path` and return a dict of numpy arrays (the model).  `transform(root)`: optional edit of the parsed element tree
    (after include expansion) before compilation - used for scene variants that the reference realised by editing the XML."""
    base_dir = os.path.dirname(os.path.abspath(path))
    root = ET.parse(path).getroot()
    _expand_includes(root, base_dir)
    if transform is not None:
        transform(root)
    comp = root.find("compiler").attrib if root.find("compiler") is not None else {}
    assert comp.get("angle", "degree") == "radian", "only angle=radian scenes are supported"
    inertiafromgeom = comp.get("inertiafromgeom", "auto") == "true"
    meshdir = os.path.join(base_dir, comp.get("meshdir", ""))
    opt = root.find("option").attrib if root.find("option") is not None else {}
    defaults = _Defaults(root)

    # ---- assets: meshes (loaded lazily, only those referenced by geoms)
    mesh_files = {}
    for m in root.find("asset").findall("mesh"):
        mesh_files[m.attrib["name"]] = os.path.join(meshdir, m.attrib["file"])
    meshes, mesh_index = [], {}

    def get_mesh(name):
        if name not in mesh_index:
            tris = load_stl(mesh_files[name])
            vol, com, inertia_com = mesh_mass_properties(tris)
            hv, hf = convex_hull(tris.reshape(-1, 3))
            mesh_index[name] = len(meshes)
            meshes.append(dict(name=name, tris=tris, volume=vol, com=com, inertia=inertia_com,
                               hull_vert=hv, hull_face=hf))
        return mesh_index[name]

    bodies, joints, geoms, cameras = [], [], [], []

    def add_body(node, parent, childclass):
        bid = len(bodies)
        a = node.attrib
        b = dict(name=a.get("name", "world" if parent < 0 else f"body{bid}"), parent=max(parent, 0),
                 pos=_f(a.get("pos", "0 0 0")), quat=_orientation(a), inertial=None, geoms=[], joints=[])
        childclass = a.get("childclass", childclass)
        bodies.append(b)
        for ch in node:
            if ch.tag == "inertial":
                ia = ch.attrib
                b["inertial"] = dict(pos=_f(ia.get("pos", "0 0 0")), quat=_orientation(ia), mass=float(ia["mass"]),
                                     diag=_f(ia.get("diaginertia", "0 0 0")))
            elif ch.tag in ("joint", "freejoint"):
                d = dict(DEF_JOINT)
                if ch.tag == "freejoint":
                    d["type"] = "free"  # freejoint ignores defaults
                else:
                    d.update(defaults.lookup("joint", ch.attrib.get("class", childclass)))
                d.update(ch.attrib)
                j = dict(name=d.get("name", f"joint{len(joints)}"), type=JOINT_TYPES[d["type"]], body=bid,
                         pos=_f(d["pos"]), axis=_f(d["axis"]), limited=d["limited"] == "true",
                         range=_f(d["range"]), damping=float(d["damping"]), armature=float(d["armature"]),
                         ref=float(d["ref"]), margin=float(d["margin"]), solref=_f(d["solreflimit"]),
                         solimp=_solimp5(d["solimplimit"]))
                assert float(d["stiffness"]) == 0.0, "joint stiffness not supported"
                if j["type"] in (J_HINGE, J_SLIDE):
                    j["axis"] = j["axis"] / np.linalg.norm(j["axis"])
                b["joints"].append(len(joints))
                joints.append(j)
            elif ch.tag == "geom":
                d = dict(DEF_GEOM)
                d.update(defaults.lookup("geom", ch.attrib.get("class", childclass)))
                d.update(ch.attrib)
                gtype = GEOM_TYPES[d["type"]]
                size = np.zeros(3)
                s = _f(d["size"])
                size[:len(s)] = s
                g = dict(name=d.get("name", f"geom{len(geoms)}"), type=gtype, body=bid, pos=_f(d["pos"]),
                         quat=_orientation(d), size=size, friction=_f(d["friction"]), condim=int(d["condim"]),
                         margin=float(d["margin"]), gap=float(d["gap"]), solref=_f(d["solref"]),
                         solimp=_solimp5(d["solimp"]), density=float(d["density"]),
                         mass=float(d["mass"]) if "mass" in d else None, contype=int(d["contype"]),
                         conaffinity=int(d["conaffinity"]), rgba=_f(d["rgba"]), mesh=-1,
                         material=d.get("material", ""))
                assert "fromto" not in d, "fromto not supported"
                if gtype == G_MESH:
                    g["mesh"] = get_mesh(d["mesh"])
                b["geoms"].append(len(geoms))
                geoms.append(g)
            elif ch.tag == "camera":
                ca = ch.attrib
                cameras.append(dict(name=ca.get("name", ""), body=bid, pos=_f(ca.get("pos", "0 0 0")),
                                    quat=_orientation(ca), fovy=float(ca.get("fovy", "45"))))
        for ch in node:
            if ch.tag == "body":
                add_body(ch, bid, childclass)

    add_body(root.find("worldbody"), -1, None)
    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)

    # ---- addresses
    QN = {J_FREE: 7, J_BALL: 4, J_SLIDE: 1, J_HINGE: 1}
    VN = {J_FREE: 6, J_BALL: 3, J_SLIDE: 1, J_HINGE: 1}
    nq = nv = 0
    for j in joints:
        j["qadr"], j["dadr"] = nq, nv
        nq += QN[j["type"]]
        nv += VN[j["type"]]
    qpos0 = np.zeros(nq)
    for j in joints:
        if j["type"] == J_FREE:
            b = bodies[j["body"]]
            qpos0[j["qadr"]:j["qadr"] + 3] = b["pos"]
            qpos0[j["qadr"] + 3:j["qadr"] + 7] = b["quat"]
        elif j["type"] == J_BALL:
            qpos0[j["qadr"]:j["qadr"] + 4] = [1, 0, 0, 0]
        else:
            qpos0[j["qadr"]] = j["ref"]

    # dof tables; dof_parentid follows the kinematic chain (previous dof of the same body, else last dof of
    # the nearest ancestor that has dofs)
    dof_body = np.zeros(nv, np.int32)
    dof_jnt = np.zeros(nv, np.int32)
    dof_parent = np.full(nv, -1, np.int32)
    dof_arm = np.zeros(nv)
    dof_damp = np.zeros(nv)
    body_dofadr = np.full(nbody, -1, np.int32)
    body_dofnum = np.zeros(nbody, np.int32)
    body_lastdof = np.full(nbody, -1, np.int32)  # last dof on the path world -> body (inclusive)
    for bid, b in enumerate(bodies):
        last = body_lastdof[b["parent"]] if bid > 0 else -1
        for ji in b["joints"]:
            j = joints[ji]
            for k in range(VN[j["type"]]):
                d = j["dadr"] + k
                dof_body[d], dof_jnt[d], dof_parent[d] = bid, ji, last
                dof_arm[d], dof_damp[d] = j["armature"], j["damping"]
                if body_dofadr[bid] < 0:
                    body_dofadr[bid] = d
                body_dofnum[bid] += 1
                last = d
        body_lastdof[bid] = last
    dof_madr = np.zeros(nv, np.int32)
    nM = 0
    for d in range(nv):
        dof_madr[d] = nM
        k = d
        while k >= 0:
            nM += 1
            k = dof_parent[k]

    # weld ids (body welded to nearest ancestor that has joints; bodies without any joint up to world -> 0)
    body_weld = np.zeros(nbody, np.int32)
    for bid, b in enumerate(bodies):
        if bid == 0:
            continue
        body_weld[bid] = bid if b["joints"] else body_weld[b["parent"]]
    # tree id: root dof's index of the kinematic tree each dof belongs to
    dof_tree = np.zeros(nv, np.int32)
    for d in range(nv):
        dof_tree[d] = d if dof_parent[d] < 0 else dof_tree[dof_parent[d]]

    # ---- inertia from geoms (density 1000 unless stated), composed in the body frame
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_inertia = np.zeros((nbody, 6))  # xx yy zz xy xz yz about the COM, body-frame axes
    for bid, b in enumerate(bodies):
        parts = []
        if inertiafromgeom or b["inertial"] is None:
            for gi in b["geoms"]:
                g = geoms[gi]
                if g["type"] == G_PLANE:
                    continue
                m, c_local, I_local = _geom_inertia(g, meshes)
                R = rigid.quat_to_mat(g["quat"])
                parts.append((m, g["pos"] + R @ c_local, R @ I_local @ R.T))
        if not parts and b["inertial"] is not None:
            ii = b["inertial"]
            R = rigid.quat_to_mat(ii["quat"])
            parts.append((ii["mass"], ii["pos"], R @ np.diag(ii["diag"]) @ R.T))
        if parts:
            m = sum(p[0] for p in parts)
            com = sum(p[0] * p[1] for p in parts) / m if m > 0 else np.zeros(3)
            I = np.zeros((3, 3))
            for pm, pc, pI in parts:
                r = pc - com
                I += pI + pm * (r @ r * np.eye(3) - np.outer(r, r))
            body_mass[bid], body_ipos[bid] = m, com
            body_inertia[bid] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]

    # ---- geom derived data: bounding sphere, oriented box (geom frame), mesh tables
    mesh_vertadr, mesh_vertnum, mesh_faceadr, mesh_facenum = [], [], [], []
    mesh_vert, mesh_face, mesh_center = [], [], []
    va = fa = 0
    for m in meshes:
        mesh_vertadr.append(va)
        mesh_vertnum.append(len(m["hull_vert"]))
        mesh_faceadr.append(fa)
        mesh_facenum.append(len(m["hull_face"]))
        mesh_vert.append(m["hull_vert"])
        mesh_face.append(m["hull_face"])
        mesh_center.append(m["hull_vert"].mean(axis=0))
        va += len(m["hull_vert"])
        fa += len(m["hull_face"])
    geom_obbc = np.zeros((ngeom, 3))
    geom_obbh = np.zeros((ngeom, 3))
    geom_rbound = np.zeros(ngeom)
    for gi, g in enumerate(geoms):
        t, s = g["type"], g["size"]
        if t == G_SPHERE:
            geom_obbh[gi] = s[0]
            geom_rbound[gi] = s[0]
        elif t == G_BOX:
            geom_obbh[gi] = s
            geom_rbound[gi] = np.linalg.norm(s)
        elif t in (G_CAPSULE, G_CYLINDER):
            geom_obbh[gi] = [s[0], s[0], s[1] + (s[0] if t == G_CAPSULE else 0.0)]
            geom_rbound[gi] = s[0] + s[1] if t == G_CAPSULE else np.hypot(s[0], s[1])
        elif t == G_MESH:
            hv = meshes[g["mesh"]]["hull_vert"]
            lo, hi = hv.min(axis=0), hv.max(axis=0)
            geom_obbc[gi] = 0.5 * (lo + hi)
            geom_obbh[gi] = 0.5 * (hi - lo)
            geom_rbound[gi] = np.linalg.norm(hv - geom_obbc[gi], axis=1).max()
        elif t == G_PLANE:
            geom_rbound[gi] = 0.0

    # ---- contact excludes and the static candidate pair list
    excl = set()
    names = {b["name"]: i for i, b in enumerate(bodies)}
    con = root.find("contact")
    if con is not None:
        for e in con.findall("exclude"):
            b1, b2 = names[e.attrib["body1"]], names[e.attrib["body2"]]
            excl.add((min(b1, b2), max(b1, b2)))
    pairs = []
    for g1 in range(ngeom):
        for g2 in range(g1 + 1, ngeom):
            a, b = geoms[g1], geoms[g2]
            b1, b2 = a["body"], b["body"]
            w1, w2 = body_weld[b1], body_weld[b2]
            if w1 == w2:
                continue  # same body / welded together (covers static-static)
            wp1, wp2 = body_weld[bodies[w1]["parent"]], body_weld[bodies[w2]["parent"]]
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue  # parent-child filter (world exempt)
            if (min(b1, b2), max(b1, b2)) in excl:
                continue
            if not ((a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"])):
                continue
            # order so that the lower geom type comes first (dispatch table is upper-triangular)
            p = (g1, g2) if a["type"] <= b["type"] else (g2, g1)
            pairs.append(p)
    npair = len(pairs)
    pair_geom = np.array(pairs, np.int32).reshape(npair, 2)
    pair_condim = np.zeros(npair, np.int32)
    pair_friction = np.zeros((npair, 3))
    pair_margin = np.zeros(npair)
    pair_solref = np.zeros((npair, 2))
    pair_solimp = np.zeros((npair, 5))
    for i, (g1, g2) in enumerate(pairs):
        a, b = geoms[g1], geoms[g2]
        pair_condim[i] = max(a["condim"], b["condim"])
        pair_friction[i] = np.maximum(a["friction"], b["friction"])
        pair_margin[i] = max(a["margin"], b["margin"]) - max(a["gap"], b["gap"])
        # solmix defaults to 1 for both geoms -> plain average; solref (positive) and solimp mix alike
        pair_solref[i] = 0.5 * (a["solref"] + b["solref"])
        pair_solimp[i] = 0.5 * (a["solimp"] + b["solimp"])

    # compact broad-phase records: (geom1, geom2, type1, type2) and the bounding-sphere reach of the pair
    pair_rec = np.zeros((npair, 4), np.int32)
    pair_rsum = np.zeros(npair)
    for i, (g1, g2) in enumerate(pairs):
        pair_rec[i] = [g1, g2, geoms[g1]["type"], geoms[g2]["type"]]
        pair_rsum[i] = (0.0 if geoms[g1]["type"] == G_PLANE else geom_rbound[g1]) + geom_rbound[g2] + pair_margin[i]

    # ---- actuators, equality
    act = root.find("actuator")
    jname = {j["name"]: i for i, j in enumerate(joints)}
    motors = []
    if act is not None:
        for m in act.findall("motor"):
            d = dict(DEF_MOTOR)
            d.update(defaults.lookup("motor", m.attrib.get("class")))
            d.update(m.attrib)
            motors.append(dict(name=d.get("name", ""), jnt=jname[d["joint"]], gear=_f(d["gear"])[0],
                               limited=d["ctrllimited"] == "true", range=_f(d["ctrlrange"])))
    eqs = []
    eqn = root.find("equality")
    if eqn is not None:
        for e in eqn.findall("joint"):
            d = dict(DEF_EQ)
            d.update(e.attrib)
            eqs.append(dict(j1=jname[d["joint1"]], j2=jname[d["joint2"]] if "joint2" in d else -1,
                            poly=_f(d["polycoef"]), solref=_f(d["solref"]), solimp=_solimp5(d["solimp"])))

    vis = root.find("visual")
    znear, zfar = 0.01, 50.0
    if vis is not None and vis.find("map") is not None:
        znear = float(vis.find("map").attrib.get("znear", znear))
        zfar = float(vis.find("map").attrib.get("zfar", zfar))

    M = dict(
        nbody=nbody, njnt=njnt, nq=nq, nv=nv, nu=len(motors), ngeom=ngeom, neq=len(eqs), npair=npair,
        nmesh=len(meshes), nM=nM, ncam=len(cameras),
        opt_timestep=float(opt.get("timestep", 0.002)), opt_gravity=_f(opt.get("gravity", "0 0 -9.81")),
        opt_iterations=int(opt.get("iterations", 100)), opt_tolerance=float(opt.get("tolerance", 1e-8)),
        opt_impratio=float(opt.get("impratio", 1.0)),
        opt_mpr_tolerance=float(opt.get("mpr_tolerance", 1e-6)), opt_mpr_iterations=int(opt.get("mpr_iterations", 50)),
        vis_znear=znear, vis_zfar=zfar,
        qpos0=qpos0,
        body_parentid=np.array([b["parent"] for b in bodies], np.int32), body_weldid=body_weld,
        body_pos=np.array([b["pos"] for b in bodies]), body_quat=np.array([b["quat"] for b in bodies]),
        body_mass=body_mass, body_ipos=body_ipos, body_inertia=body_inertia,
        body_jntadr=np.array([b["joints"][0] if b["joints"] else -1 for b in bodies], np.int32),
        body_jntnum=np.array([len(b["joints"]) for b in bodies], np.int32),
        body_dofadr=body_dofadr, body_dofnum=body_dofnum, body_lastdof=body_lastdof,
        jnt_type=np.array([j["type"] for j in joints], np.int32),
        jnt_bodyid=np.array([j["body"] for j in joints], np.int32),
        jnt_qposadr=np.array([j["qadr"] for j in joints], np.int32),
        jnt_dofadr=np.array([j["dadr"] for j in joints], np.int32),
        jnt_pos=np.array([j["pos"] for j in joints]).reshape(njnt, 3),
        jnt_axis=np.array([j["axis"] for j in joints]).reshape(njnt, 3),
        jnt_limited=np.array([j["limited"] and j["type"] in (J_HINGE, J_SLIDE) for j in joints], np.int32),
        jnt_range=np.array([j["range"] for j in joints]).reshape(njnt, 2),
        jnt_margin=np.array([j["margin"] for j in joints]),
        jnt_solref=np.array([j["solref"] for j in joints]).reshape(njnt, 2),
        jnt_solimp=np.array([j["solimp"] for j in joints]).reshape(njnt, 5),
        dof_bodyid=dof_body, dof_jntid=dof_jnt, dof_parentid=dof_parent, dof_Madr=dof_madr, dof_treeid=dof_tree,
        dof_armature=dof_arm, dof_damping=dof_damp,
        geom_type=np.array([g["type"] for g in geoms], np.int32),
        geom_bodyid=np.array([g["body"] for g in geoms], np.int32),
        geom_pos=np.array([g["pos"] for g in geoms]), geom_quat=np.array([g["quat"] for g in geoms]),
        geom_size=np.array([g["size"] for g in geoms]), geom_rbound=geom_rbound,
        geom_obbcenter=geom_obbc, geom_obbhalf=geom_obbh,
        geom_meshid=np.array([g["mesh"] for g in geoms], np.int32),
        geom_rgba=np.array([g["rgba"] for g in geoms]),
        mesh_vertadr=np.array(mesh_vertadr, np.int32), mesh_vertnum=np.array(mesh_vertnum, np.int32),
        mesh_faceadr=np.array(mesh_faceadr, np.int32), mesh_facenum=np.array(mesh_facenum, np.int32),
        mesh_vert=np.concatenate(mesh_vert) if mesh_vert else np.zeros((0, 3)),
        mesh_face=np.concatenate(mesh_face).astype(np.int32) if mesh_face else np.zeros((0, 3), np.int32),
        mesh_center=np.array(mesh_center).reshape(len(meshes), 3),
        pair_geom=pair_geom, pair_condim=pair_condim, pair_rec=pair_rec, pair_rsum=pair_rsum, pair_friction=pair_friction, pair_margin=pair_margin,
        pair_solref=pair_solref, pair_solimp=pair_solimp,
        actuator_jntid=np.array([m["jnt"] for m in motors], np.int32),
        actuator_gear=np.array([m["gear"] for m in motors]),
        actuator_ctrlrange=np.array([m["range"] if m["limited"] else [-1e30, 1e30] for m in motors]).reshape(-1, 2),
        eq_jnt1=np.array([e["j1"] for e in eqs], np.int32), eq_jnt2=np.array([e["j2"] for e in eqs], np.int32),
        eq_polycoef=np.array([e["poly"] for e in eqs]).reshape(-1, 5),
        eq_solref=np.array([e["solref"] for e in eqs]).reshape(-1, 2),
        eq_solimp=np.array([e["solimp"] for e in eqs]).reshape(-1, 5),
    )
    # cameras: fixed in the world (all three hang off worldbody in both scenes); cam_mat0 is the rotation
    # matrix whose columns are the camera axes in world coordinates (camera looks along its -z)
    for c in cameras:
        assert c["body"] == 0, "only world-fixed cameras are supported"
    M["cam_pos0"] = np.array([c["pos"] for c in cameras]).reshape(-1, 3)
    M["cam_mat0"] = np.array([rigid.quat_to_mat(c["quat"]).reshape(9) for c in cameras]).reshape(-1, 9)
    M["cam_fovy"] = np.array([c["fovy"] for c in cameras])

    # ---- traversal tables for the warp-parallel kernels (bodies and dofs are numbered depth-first, so every
    # subtree is a contiguous index range)
    body_subtreenum = np.ones(nbody, np.int32)
    for bid in range(nbody - 1, 0, -1):
        body_subtreenum[bodies[bid]["parent"]] += body_subtreenum[bid]
    dof_subtreenum = np.ones(nv, np.int32)
    dof_depth = np.zeros(nv, np.int32)
    for d in range(nv - 1, -1, -1):
        if dof_parent[d] >= 0:
            dof_subtreenum[dof_parent[d]] += dof_subtreenum[d]
    for d in range(nv):
        dof_depth[d] = 0 if dof_parent[d] < 0 else dof_depth[dof_parent[d]] + 1
    tree_roots = [d for d in range(nv) if dof_parent[d] < 0]
    M["body_subtreenum"] = body_subtreenum
    M["dof_subtreenum"] = dof_subtreenum
    M["dof_depth"] = dof_depth
    M["dof_treeindex"] = np.array([tree_roots.index(int(dof_tree[d])) for d in range(nv)], np.int32)
    # per body: bit l set <=> tree-local dof l (dof - root of its tree) lies on the path world -> body (trees of <= 32 dofs)
    chainmask = np.zeros(nbody, np.int64)
    for bid in range(nbody):
        d = body_lastdof[bid]
        while d >= 0:
            loc = d - int(dof_tree[d])
            if loc < 31:
                chainmask[bid] |= 1 << loc
            d = dof_parent[d]
    M["body_chainmask"] = chainmask.astype(np.int32)
    M["tree_dofadr"] = np.array(tree_roots, np.int32)
    M["tree_dofnum"] = np.array([dof_subtreenum[d] for d in tree_roots], np.int32)
    M["ntree"] = len(tree_roots)
    M["geom_lmat"] = np.array([rigid.quat_to_mat(g["quat"]).reshape(9) for g in geoms]).reshape(ngeom, 9)
    # outward face planes of every hull (normal, offset), geom-local: used by the ray-caster
    planes = []
    for m_ in meshes:
        hv, hf = m_["hull_vert"], m_["hull_face"]
        nrm = np.cross(hv[hf[:, 1]] - hv[hf[:, 0]], hv[hf[:, 2]] - hv[hf[:, 0]])
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
        planes.append(np.concatenate([nrm, np.einsum("ij,ij->i", nrm, hv[hf[:, 0]])[:, None]], axis=1))
    M["mesh_faceplane"] = np.concatenate(planes) if planes else np.zeros((0, 4))

    # ---- quantities derived at qpos0: invweight0, mean inertia, extent
    kin = rigid.Kinematics(M)
    Mq = kin.mass_matrix(qpos0)
    Minv = np.linalg.inv(Mq)
    dof_inv = np.diag(Minv).copy()
    for j in joints:
        d = j["dadr"]
        if j["type"] == J_BALL:
            dof_inv[d:d + 3] = dof_inv[d:d + 3].mean()
        elif j["type"] == J_FREE:
            dof_inv[d:d + 3] = dof_inv[d:d + 3].mean()
            dof_inv[d + 3:d + 6] = dof_inv[d + 3:d + 6].mean()
    body_inv = np.zeros((nbody, 2))
    st = kin.forward(qpos0)
    for bid in range(1, nbody):
        if body_lastdof[bid] < 0:
            continue
        Jp, Jr = kin.jacobian(st, bid, st["xipos"][bid])
        A = Jp @ Minv @ Jp.T
        B = Jr @ Minv @ Jr.T
        body_inv[bid] = [np.trace(A) / 3, np.trace(B) / 3]
    # "simple" trees: a single free-floating body (3 world slides + ball, or one free joint) whose centre of mass sits on the joint
    # anchor.  Their mass-matrix block does not depend on the configuration (m*1 (+) I_body, + armature), so (M)^-1 and
    # (M + h*damping)^-1 are model constants: the sub-step kernel multiplies by them instead of factorising (tree_Minv[t,0|1,6,6]).
    ntree = len(tree_roots)
    tree_simple = np.zeros(ntree, np.int32)
    tree_Minv = np.zeros((ntree, 2, 6, 6))
    hstep = float(opt.get("timestep", 0.002))
    rng_chk = np.random.RandomState(0)
    for t, r in enumerate(tree_roots):
        n = int(dof_subtreenum[r])
        bset = sorted(set(int(dof_body[d]) for d in range(r, r + n)))
        if n != 6 or len(bset) != 1:
            continue
        # numerical check of configuration independence: M block at qpos0 vs at a random configuration
        q2 = qpos0.copy()
        for j in bodies[bset[0]]["joints"]:
            jt, qa = joints[j]["type"], joints[j]["qadr"]
            if jt in (J_SLIDE, J_HINGE):
                q2[qa] += rng_chk.uniform(-0.3, 0.3)
            else:
                if jt == J_FREE:
                    q2[qa:qa + 3] += rng_chk.uniform(-0.3, 0.3, 3)
                    qa += 3
                qq = rng_chk.normal(size=4)
                q2[qa:qa + 4] = qq / np.linalg.norm(qq)
        B0 = Mq[r:r + 6, r:r + 6]
        B1 = kin.mass_matrix(q2)[r:r + 6, r:r + 6]
        if np.abs(B0 - B1).max() > 1e-13 * max(1.0, np.abs(B0).max()):
            continue
        tree_simple[t] = 1
        tree_Minv[t, 0] = np.linalg.inv(B0)
        tree_Minv[t, 1] = np.linalg.inv(B0 + hstep * np.diag(dof_damp[r:r + 6]))
    M["tree_simple"] = tree_simple
    M["tree_Minv"] = tree_Minv
    M["dof_invweight0"] = dof_inv
    M["body_invweight0"] = body_inv
    M["stat_meaninertia"] = float(np.trace(Mq) / max(nv, 1))
    # model extent: half-diagonal of the world AABB of geom bounding spheres at qpos0 (documented meaning of
    # mjModel.stat.extent); it cancels in depth_2_meters so only its positivity matters
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for gi, g in enumerate(geoms):
        if g["type"] == G_PLANE:
            continue
        c = st["xpos"][g["body"]] + st["xmat"][g["body"]] @ g["pos"]
        lo, hi = np.minimum(lo, c - geom_rbound[gi]), np.maximum(hi, c + geom_rbound[gi])
    M["stat_extent"] = float(0.5 * np.linalg.norm(hi - lo))
    M["_names"] = dict(body=[b["name"] for b in bodies], joint=[j["name"] for j in joints],
                       geom=[g["name"] for g in geoms], camera=[c["name"] for c in cameras],
                       actuator=[m["name"] for m in motors], mesh=[m["name"] for m in meshes])
    M["_render_meshes"] = [m["tris"] for m in meshes]
    M["_materials"] = [g["material"] for g in geoms]
    return M


def _geom_inertia(g, meshes):
    """mass, COM (geom frame), inertia about COM (geom frame, 3x3) of a solid geom at its density."""
    t, s, rho = g["type"], g["size"], g["density"]
    c = np.zeros(3)
    if t == G_SPHERE:
        v = 4.0 / 3.0 * np.pi * s[0] ** 3
        I = np.eye(3) * 0.4 * s[0] ** 2
    elif t == G_BOX:
        v = 8.0 * s[0] * s[1] * s[2]
        I = np.diag([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]) / 3.0
    elif t == G_CYLINDER:
        r, h = s[0], s[1]
        v = np.pi * r * r * 2 * h
        I = np.diag([(3 * r * r + 4 * h * h) / 12.0] * 2 + [r * r / 2.0])
    elif t == G_CAPSULE:
        r, h = s[0], s[1]
        vc, vs = np.pi * r * r * 2 * h, 4.0 / 3.0 * np.pi * r ** 3
        v = vc + vs
        # cylinder + two hemispheres (parallel axis for the caps); per unit volume below, scaled by v at the end
        ix = vc * (3 * r * r + 4 * h * h) / 12.0 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
        iz = vc * r * r / 2.0 + vs * 0.4 * r * r
        I = np.diag([ix, ix, iz]) / v
    elif t == G_MESH:
        m = meshes[g["mesh"]]
        v, c = m["volume"], m["com"]
        I = m["inertia"] / v
    else:
        raise ValueError(t)
    mass = g["mass"] if g["mass"] is not None else rho * v
    return mass, c, I * mass


def model_to_blob_arrays(M):
    """Drop the python-only entries (names, render meshes) and return what goes into the blob."""
    return {k: np.asarray(v) for k, v in M.items() if not k.startswith("_")}
