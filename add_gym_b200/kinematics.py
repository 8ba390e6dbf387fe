"""Kinematic description of the character (host side, construction time only).

Mirrors the part of the reference's ``KinCharModel`` the hot path depends on
(reference add_gym/anim/kin_char_model.py:72-224): bodies are enumerated breadth-first over the
MJCF ``<worldbody>`` tree (kin_char_model.py:116-161, "to match Genesis ordering"), body 0 is the
floating root, every other body carries one hinge (1 dof) or is fixed (0 dof), and dof indices
are assigned in body order.  The device kernels only need three flat arrays out of this:
the hinge axis per dof, the dof -> motion-file column permutation, and the joint limits.

Two sources are accepted: the MJCF ``.xml`` itself, or the ``.json`` digest of it that ships in
``assets/`` (written by oracle/make_assets.py) so that nothing needs /root/reference at run time.
"""
import json
import os
import xml.etree.ElementTree as ET

import numpy as np

JOINT_ROOT, JOINT_HINGE, JOINT_SPHERICAL, JOINT_FIXED = 0, 1, 2, 3


class Joint:
    def __init__(self, name, joint_type, axis, limit=None):
        self.name = name
        self.joint_type = joint_type
        self.axis = None if axis is None else np.asarray(axis, dtype=np.float32)
        self.limit = limit  # (lo, hi) or None
        self.dof_idx = -1

    def get_dof_dim(self):
        return {JOINT_ROOT: 0, JOINT_HINGE: 1, JOINT_SPHERICAL: 3, JOINT_FIXED: 0}[self.joint_type]


class KinCharModel:
    def __init__(self, device="cpu"):
        self._device = device
        self._body_names = []
        self._parent_indices = []
        self._local_translation = []
        self._local_rotation = []
        self._joints = []
        self._dof_size = 0

    # ---- loading -------------------------------------------------------------------------
    def load_char_file(self, char_file):
        ext = os.path.splitext(char_file)[1]
        if ext == ".json":
            with open(char_file, "r") as f:
                bodies = json.load(f)["bodies"]
        elif ext == ".xml":
            bodies = parse_mjcf_bodies(char_file)
        else:
            raise AssertionError("Unsupported character file format: {:s}".format(ext))
        self._init_from_bodies(bodies)

    def _init_from_bodies(self, bodies):
        for b in bodies:
            self._body_names.append(b["name"])
            self._parent_indices.append(b["parent"])
            self._local_translation.append(b["pos"])
            self._local_rotation.append(b["quat_xyzw"])
            j = b["joint"]
            jt = {"root": JOINT_ROOT, "hinge": JOINT_HINGE, "fixed": JOINT_FIXED}[j["type"]]
            if jt == JOINT_SPHERICAL:
                raise AssertionError("spherical joints are not supported on the B200 path")
            self._joints.append(Joint(j["name"], jt, j.get("axis"), j.get("range")))
        dof = 0
        for j in self._joints:
            d = j.get_dof_dim()
            if d > 0:
                j.dof_idx = dof
                dof += d
        self._dof_size = dof

    # ---- queries (same names as the reference) -----------------------------------------------
    def get_body_names(self):
        return self._body_names

    def get_num_joints(self):
        return len(self._joints)

    def get_joint(self, j):
        assert j > 0
        return self._joints[j]

    def get_dof_size(self):
        return self._dof_size

    def get_joint_dof_idx(self, j):
        return self.get_joint(j).dof_idx

    def get_joint_dof_dim(self, j):
        return self.get_joint(j).get_dof_dim()

    def get_joint_order(self):
        return [j.name for j in self._joints]

    def get_body_id(self, body_name):
        return self._body_names.index(body_name)

    # ---- flat arrays for the kernels ----------------------------------------------------------
    def dof_axes(self):
        """[dof_size, 3] float32 hinge axis of every dof, in dof (BFS) order."""
        ax = np.zeros((self._dof_size, 3), dtype=np.float32)
        for j in self._joints[1:]:
            if j.joint_type == JOINT_HINGE:
                ax[j.dof_idx] = j.axis
        return ax

    def dof_limits(self):
        lim = np.zeros((self._dof_size, 2), dtype=np.float32)
        for j in self._joints[1:]:
            if j.joint_type == JOINT_HINGE:
                lim[j.dof_idx] = j.limit if j.limit is not None else (-np.inf, np.inf)
        return lim

    def motion_column_of_dof(self, motion_order):
        """For every dof (BFS order) the column of the motion file that holds it
        (reference motion_lib.py:102-111)."""
        names = [j.name for j in self._joints[1:] if j.get_dof_dim() > 0]
        return np.asarray([motion_order.index(n) for n in names], dtype=np.int32)


def parse_mjcf_bodies(xml_file):
    """Breadth-first body list of an MJCF file (one hinge or no joint per non-root body)."""
    root = ET.parse(xml_file).getroot()
    world = root.find("worldbody")
    assert world is not None
    body_root = world.find("body")
    assert body_root is not None
    out = []
    queue = [(body_root, -1, True)]
    while queue:
        node, parent, is_root = queue.pop(0)
        pos = node.attrib.get("pos")
        pos = [0.0, 0.0, 0.0] if pos is None else [float(v) for v in pos.split()]
        quat = node.attrib.get("quat")
        if quat is None:
            q = [0.0, 0.0, 0.0, 1.0]
        else:
            w, x, y, z = [float(v) for v in quat.split()]
            q = [x, y, z, w]
        if is_root:
            joint = {"name": "root", "type": "root"}
        else:
            jts = node.findall("joint")
            if len(jts) == 0:
                joint = {"name": node.attrib.get("name"), "type": "fixed"}
            elif len(jts) == 1:
                jt = jts[0]
                assert jt.attrib.get("type", "hinge") == "hinge", "only hinge joints are supported"
                jpos = jt.attrib.get("pos")
                if jpos is not None:
                    assert not any(float(v) != 0.0 for v in jpos.split()), "joint offsets unsupported"
                rng = jt.attrib.get("range")
                joint = {
                    "name": jt.attrib.get("name"),
                    "type": "hinge",
                    "axis": [float(v) for v in jt.attrib.get("axis").split()],
                    "range": None if rng is None else [float(v) for v in rng.split()],
                }
            else:
                raise AssertionError("series / spherical joints are not supported")
        idx = len(out)
        out.append({"name": node.attrib.get("name"), "parent": parent, "pos": pos,
                    "quat_xyzw": q, "joint": joint})
        for child in node.findall("body"):
            queue.append((child, idx, False))
    return out
