"""Physics-engine boundary: the abstract interface the hot path talks to, and the
physics-free synthetic-state engine used by the tests and by bench.py.

The ABCs keep the names and tensor conventions of the reference's engine layer
(add_gym/engine/base_engine.py:93-511): getters return row-major fp32 tensors on
``engine.device`` -- ``get_pos [N,3]``, ``get_quat [N,4]`` (wxyz), ``get_vel/get_ang [N,3]``,
``get_dofs_position/velocity [N,6+D]`` (6 floating-base dofs first, hinge dofs in BFS order) --
writers take ``envs_idx``.  Genesis / MuJoCo-Warp stay behind this interface (out of scope here).

`SyntheticEngine` serves seeded random-walk humanoid states (BASELINE config 5).  Its
state stream does not depend on the applied action, only on the seed and on ``set_qpos`` /
``set_dofs_velocity`` calls, so a CPU oracle and the CUDA path driven with the same seed see
bit-identical physics.  One extension over the reference interface: ``set_state_masked`` lets the
agent overwrite the state of "done" environments without first compacting their indices on the
host (the reference's ``nonzero`` + ``len`` sync, base_agent.py:449-453).
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import kinematics


class BaseLink(ABC):
    @property
    @abstractmethod
    def idx(self):
        ...

    @property
    @abstractmethod
    def idx_local(self):
        ...

    @property
    @abstractmethod
    def name(self):
        ...


class BaseJoint(ABC):
    @property
    @abstractmethod
    def dofs_idx(self):
        ...

    @property
    @abstractmethod
    def dofs_idx_local(self):
        ...

    @property
    @abstractmethod
    def dofs_limit(self):
        ...

    @property
    @abstractmethod
    def name(self):
        ...


class BaseEntity(ABC):
    """Subset of the reference BaseEntity the ADD hot path calls (base_engine.py:93-376)."""

    @abstractmethod
    def get_pos(self):
        ...

    @abstractmethod
    def get_quat(self):
        ...

    @abstractmethod
    def get_vel(self):
        ...

    @abstractmethod
    def get_ang(self):
        ...

    @abstractmethod
    def get_dofs_position(self):
        ...

    @abstractmethod
    def get_dofs_velocity(self):
        ...

    @abstractmethod
    def set_qpos(self, qpos, envs_idx=None):
        ...

    @abstractmethod
    def set_dofs_velocity(self, velocity, envs_idx=None, dofs_idx_local=None):
        ...

    @abstractmethod
    def control_dofs_position(self, position, dofs_idx_local=None):
        ...

    @abstractmethod
    def get_contacts(self, with_entity=None, exclude_self_contact=False):
        ...


class BaseScene(ABC):
    @abstractmethod
    def add_entity(self, morph_type, **kwargs):
        ...

    @abstractmethod
    def build(self, n_envs, env_spacing):
        ...

    @abstractmethod
    def step(self):
        ...

    @property
    @abstractmethod
    def t(self):
        ...


class BaseEngine(ABC):
    @abstractmethod
    def init(self, backend, precision):
        ...

    @abstractmethod
    def create_scene(self, show_viewer=False, **options):
        ...

    @property
    @abstractmethod
    def device(self):
        ...

    @property
    @abstractmethod
    def tc_float(self):
        ...


# ------------------------------------------------------------------------------------------------
# Synthetic engine
# ------------------------------------------------------------------------------------------------
class _Link(BaseLink):
    def __init__(self, name, idx, idx_local):
        self._name, self._idx, self._idx_local = name, idx, idx_local

    idx = property(lambda s: s._idx)
    idx_local = property(lambda s: s._idx_local)
    name = property(lambda s: s._name)


class _Joint(BaseJoint):
    def __init__(self, name, dofs_idx, dofs_limit):
        self._name, self._dofs, self._lim = name, list(dofs_idx), list(dofs_limit)

    dofs_idx = property(lambda s: s._dofs)
    dofs_idx_local = property(lambda s: s._dofs)
    dofs_limit = property(lambda s: s._lim)
    name = property(lambda s: s._name)


class SyntheticPlane:
    def __init__(self):
        self.links = [_Link("plane", 0, 0)]


class SyntheticEntity(BaseEntity):
    N_CONTACT_SLOTS = 4
    persistent_state_tensors = True     # getters return the same tensors every step (CUDA-graph friendly)

    def __init__(self, scene, char_file, link_offset):
        self._scene = scene
        kin = kinematics.KinCharModel()
        kin.load_char_file(char_file)
        self._kin = kin
        names = kin.get_body_names()
        self._links = [_Link(n, link_offset + i, i) for i, n in enumerate(names)]
        inf = float("inf")
        self._joints = [_Joint("floating_base_joint", range(6), [(-inf, inf)] * 6)]
        lim = kin.dof_limits()
        for j in range(1, kin.get_num_joints()):
            jt = kin.get_joint(j)
            if jt.get_dof_dim() == 1:
                d = 6 + jt.dof_idx
                self._joints.append(_Joint(jt.name, [d], [(float(lim[jt.dof_idx, 0]), float(lim[jt.dof_idx, 1]))]))
        self._n_dofs = 6 + kin.get_dof_size()
        self._foot_links = [l.idx for l in self._links if "ankle_roll" in l.name]

    # -- built once the env count is known
    def _build(self, n_envs, device, dtype):
        D = self._n_dofs
        self._n = n_envs
        self._dofs_pos = torch.zeros(n_envs, D, device=device, dtype=dtype)
        self._dofs_pos[:, 2] = 0.793
        self._quat = torch.zeros(n_envs, 4, device=device, dtype=dtype)
        self._quat[:, 0] = 1.0
        self._dofs_vel = torch.zeros(n_envs, D, device=device, dtype=dtype)
        self._target = torch.zeros(n_envs, D - 6, device=device, dtype=dtype)
        lim = torch.tensor(self._kin.dof_limits(), device=device, dtype=dtype)
        self._lim_lo, self._lim_hi = lim[:, 0].contiguous(), lim[:, 1].contiguous()
        C = self.N_CONTACT_SLOTS
        self._link_a = torch.zeros(n_envs, C, device=device, dtype=torch.int32)
        self._link_b = torch.zeros(n_envs, C, device=device, dtype=torch.int32)
        self._valid = torch.zeros(n_envs, C, device=device, dtype=torch.bool)
        feet = (self._foot_links + [1, 1])[:2]
        self._link_b[:, 0] = feet[0]
        self._link_b[:, 1] = feet[1]
        self._valid[:, 0:2] = True
        self._kp = None
        self._kv = None

    # -- BaseEntity getters: persistent tensors / views, never reallocated
    def get_pos(self):
        return self._dofs_pos[:, 0:3]

    def get_quat(self):
        return self._quat

    def get_vel(self):
        return self._dofs_vel[:, 0:3]

    def get_ang(self):
        return self._dofs_vel[:, 3:6]

    def get_dofs_position(self):
        return self._dofs_pos

    def get_dofs_velocity(self):
        return self._dofs_vel

    def get_links_pos(self):
        raise NotImplementedError("synthetic engine has no link kinematics")

    def get_AABB(self):
        aabb = torch.zeros(self._n, 2, 3, device=self._dofs_pos.device)
        aabb[:, 1, 2] = 1.3
        return aabb

    def get_joint(self, name):
        for j in self._joints:
            if j.name == name:
                return j
        raise KeyError(name)

    def get_link(self, name):
        for l in self._links:
            if l.name == name:
                return l
        raise KeyError(name)

    joints = property(lambda s: s._joints)
    links = property(lambda s: s._links)
    n_dofs = property(lambda s: s._n_dofs)

    # -- writers
    def _rows(self, envs_idx):
        return slice(None) if envs_idx is None else envs_idx

    def set_pos(self, pos, envs_idx=None):
        self._dofs_pos[self._rows(envs_idx), 0:3] = pos

    def set_quat(self, quat, envs_idx=None):
        self._quat[self._rows(envs_idx)] = quat

    def set_dofs_position(self, position, envs_idx=None, dofs_idx_local=None):
        r = self._rows(envs_idx)
        if dofs_idx_local is None:
            self._dofs_pos[r] = position
        else:
            cols = torch.as_tensor(dofs_idx_local, device=self._dofs_pos.device)
            if envs_idx is None:
                self._dofs_pos[:, cols] = position
            else:
                self._dofs_pos[envs_idx.unsqueeze(-1), cols.unsqueeze(0)] = position

    def set_qpos(self, qpos, envs_idx=None):
        r = self._rows(envs_idx)
        self._dofs_pos[r, 0:3] = qpos[:, 0:3]
        self._quat[r] = qpos[:, 3:7]
        self._dofs_pos[r, 6:] = qpos[:, 7:]

    def set_dofs_velocity(self, velocity, envs_idx=None, dofs_idx_local=None):
        assert dofs_idx_local is None
        self._dofs_vel[self._rows(envs_idx)] = velocity

    def zero_all_dofs_velocity(self, envs_idx=None):
        self._dofs_vel[self._rows(envs_idx)] = 0.0

    def set_dofs_kp(self, kp):
        self._kp = kp

    def set_dofs_kv(self, kv):
        self._kv = kv

    def control_dofs_position(self, position, dofs_idx_local=None):
        self._target = position

    def get_contacts(self, with_entity=None, exclude_self_contact=False):
        return {"link_a": self._link_a, "link_b": self._link_b, "valid_mask": self._valid}

    # -- extension: masked state write without an index list (no host sync)
    def set_state_masked(self, mask, qpos, qvel):
        """``mask`` bool[N]; ``qpos`` [N,7+D] (pos, quat wxyz, dof); ``qvel`` [N,6+D]."""
        m = mask.unsqueeze(-1)
        self._dofs_pos[:, 0:3] = torch.where(m, qpos[:, 0:3], self._dofs_pos[:, 0:3])
        self._quat[:] = torch.where(m, qpos[:, 3:7], self._quat)
        self._dofs_pos[:, 6:] = torch.where(m, qpos[:, 7:], self._dofs_pos[:, 6:])
        self._dofs_vel[:] = torch.where(m, qvel, self._dofs_vel)


class SyntheticScene(BaseScene):
    def __init__(self, engine):
        self._engine = engine
        self._entities = []
        self._robots = []
        self._t = 0
        self._n_links = 0

    def add_entity(self, morph_type, morph_file=None, **kwargs):
        if morph_type == "plane":
            ent = SyntheticPlane()
            self._n_links += 1
        else:
            ent = SyntheticEntity(self, morph_file, self._n_links)
            self._n_links += len(ent.links)
            self._robots.append(ent)
        self._entities.append(ent)
        return ent

    def add_camera(self, **kwargs):
        raise NotImplementedError("no rendering in the synthetic engine")

    def build(self, n_envs, env_spacing=None):
        for r in self._robots:
            r._build(n_envs, self._engine.device, self._engine.tc_float)

    def step(self):
        self._engine._advance(self._robots[0])
        self._t += 1

    t = property(lambda s: s._t)


class SyntheticEngine(BaseEngine):
    """Random-walk humanoid states; see module docstring.

    noise_device "cpu": draws come from a CPU ``torch.Generator`` and are copied to the device
    (bit-identical stream for a CPU oracle and the CUDA path; used by the parity tests).
    noise_device "device": draws come from a generator on ``device`` (bench).
    """

    SIGMA_POS, SIGMA_QUAT, SIGMA_VEL, SIGMA_ANG = 0.004, 0.01, 0.5, 0.5
    SIGMA_DOF, SIGMA_DOFVEL = 0.01, 1.0

    def __init__(self, num_envs=4, ctrl_dt=0.01, seed=1234, noise_device="cpu", fall_prob=0.002,
                 device=None, **unused):
        self._num_envs = num_envs
        self._seed = seed
        self._noise_device = noise_device
        self._fall_prob = fall_prob
        self._device = torch.device(device) if device is not None else None
        self._gen = None

    def init(self, backend="cpu", precision="32"):
        if self._device is None:
            self._device = torch.device("cuda:0" if backend == "gpu" else "cpu")
        gdev = "cpu" if self._noise_device == "cpu" else self._device
        self._gen = torch.Generator(device=gdev)
        self._gen.manual_seed(self._seed)
        self._gdev = gdev

    def create_scene(self, show_viewer=False, **options):
        return SyntheticScene(self)

    device = property(lambda s: s._device)
    tc_float = property(lambda s: torch.float32)

    def _advance(self, ent):
        n, D = ent._n, ent._n_dofs - 6
        z = torch.randn(n, 13 + 2 * D, generator=self._gen, device=self._gdev).to(self._device)
        u = torch.rand(n, 2, generator=self._gen, device=self._gdev).to(self._device)
        ent._dofs_pos[:, 0:3] += self.SIGMA_POS * z[:, 0:3]
        q = ent._quat + self.SIGMA_QUAT * z[:, 3:7]
        s = q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3]
        ent._quat[:] = q / torch.sqrt(s).unsqueeze(-1)
        ent._dofs_vel[:, 0:3] = self.SIGMA_VEL * z[:, 7:10]
        ent._dofs_vel[:, 3:6] = self.SIGMA_ANG * z[:, 10:13]
        dof = ent._dofs_pos[:, 6:] + self.SIGMA_DOF * z[:, 13:13 + D]
        ent._dofs_pos[:, 6:] = torch.minimum(torch.maximum(dof, ent._lim_lo), ent._lim_hi)
        ent._dofs_vel[:, 6:] = self.SIGMA_DOFVEL * z[:, 13 + D:13 + 2 * D]
        # slot 2: a random robot link touches the plane with probability fall_prob
        n_links = len(ent.links)
        ent._valid[:, 2] = u[:, 0] < self._fall_prob
        ent._link_a[:, 2] = ent.links[0].idx + torch.clamp((u[:, 1] * n_links).to(torch.int32), max=n_links - 1)
        ent._link_b[:, 2] = 0


# ------------------------------------------------------------------------------------------------
# An engine whose contact list changes its WIDTH from step to step, like the real backends do: MuJoCo-Warp's
# get_contacts returns [nworld, 0] while nothing touches (the state at build time, mjwarp_engine.py:896-986), Genesis pads
# to the per-step maximum.  Same random walk; the one synthetic contact lands in a different column every step.
# ------------------------------------------------------------------------------------------------
class DynamicContactEntity(SyntheticEntity):
    persistent_state_tensors = False
    WIDTHS = (0, 3, 40, 1, 33, 0, 7)

    def get_contacts(self, with_entity=None, exclude_self_contact=False):
        k = int(self._scene.t)
        w = self.WIDTHS[k % len(self.WIDTHS)]
        n, dev = self._n, self._link_a.device
        la = torch.zeros(n, w, dtype=self._link_a.dtype, device=dev)
        lb = torch.zeros(n, w, dtype=self._link_b.dtype, device=dev)
        va = torch.zeros(n, w, dtype=self._valid.dtype, device=dev)
        if w > 0:
            col = (7 * k + 5) % w
            la[:, col], lb[:, col], va[:, col] = self._link_a[:, 2], self._link_b[:, 2], self._valid[:, 2]
        return {"link_a": la, "link_b": lb, "valid_mask": va}


class DynamicContactScene(SyntheticScene):
    def add_entity(self, morph_type, morph_file=None, **kwargs):
        if morph_type == "plane":
            return super().add_entity(morph_type, morph_file, **kwargs)
        ent = DynamicContactEntity(self, morph_file, self._n_links)
        self._n_links += len(ent.links)
        self._robots.append(ent)
        self._entities.append(ent)
        return ent


class DynamicContactEngine(SyntheticEngine):
    def create_scene(self, show_viewer=False, **options):
        return DynamicContactScene(self)


# ------------------------------------------------------------------------------------------------
# Host-boundary engine (bench.py "e2e"): a simulator whose state lives in pinned HOST memory
# ------------------------------------------------------------------------------------------------
class HostBoundaryEntity(SyntheticEntity):
    """Same random-walk physics, but the hot path only ever sees what crossed the host boundary: after every
    ``scene.step()`` the state is parked in pinned host memory (engine side), and the first getter call of the
    step brings it to the device with ONE packed host->device copy per dtype; ``control_dofs_position`` copies
    the action device->host.  Byte counters feed bench.py's ``h2d_bytes_per_step`` / ``d2h_bytes_per_step``."""

    def _build(self, n_envs, device, dtype):
        super()._build(n_envs, device, dtype)
        D, Cn = self._n_dofs, self.N_CONTACT_SLOTS
        self._wf = 2 * D + 4
        self._host_f = torch.zeros(n_envs, self._wf, dtype=torch.float32).pin_memory()
        self._host_i = torch.zeros(n_envs, 2 * Cn, dtype=torch.int32).pin_memory()
        self._host_v = torch.zeros(n_envs, Cn, dtype=torch.uint8).pin_memory()
        self._host_action = torch.zeros(n_envs, D - 6, dtype=torch.float32).pin_memory()
        self._stage_f = torch.zeros(n_envs, self._wf, dtype=torch.float32, device=device)
        self._stage_i = torch.zeros(n_envs, 2 * Cn, dtype=torch.int32, device=device)
        self._stage_v = torch.zeros(n_envs, Cn, dtype=torch.uint8, device=device)
        self._dirty = True
        self.h2d_bytes_per_env_step = n_envs * (self._wf * 4 + 2 * Cn * 4 + Cn)
        self.d2h_bytes_per_env_step = n_envs * (D - 6) * 4
        self._park()

    def _park(self):
        """engine side: device simulator state -> pinned host memory"""
        D, Cn = self._n_dofs, self.N_CONTACT_SLOTS
        self._host_f.copy_(torch.cat([self._dofs_pos, self._quat, self._dofs_vel], dim=1), non_blocking=True)
        self._host_i.copy_(torch.cat([self._link_a, self._link_b], dim=1), non_blocking=True)
        self._host_v.copy_(self._valid.to(torch.uint8), non_blocking=True)
        self._dirty = True

    def _fetch(self):
        if self._dirty:   # hot-path side: pinned host memory -> device staging tensors
            self._stage_f.copy_(self._host_f, non_blocking=True)
            self._stage_i.copy_(self._host_i, non_blocking=True)
            self._stage_v.copy_(self._host_v, non_blocking=True)
            self._dirty = False

    def fetch_from_host(self):
        """Explicit form of the lazy fetch (the agent calls it before replaying a captured step)."""
        self._fetch()

    def get_pos(self):
        self._fetch()
        return self._stage_f[:, 0:3]

    def get_quat(self):
        self._fetch()
        D = self._n_dofs
        return self._stage_f[:, D:D + 4]

    def get_vel(self):
        self._fetch()
        D = self._n_dofs
        return self._stage_f[:, D + 4:D + 7]

    def get_ang(self):
        self._fetch()
        D = self._n_dofs
        return self._stage_f[:, D + 7:D + 10]

    def get_dofs_position(self):
        self._fetch()
        return self._stage_f[:, 0:self._n_dofs]

    def get_dofs_velocity(self):
        self._fetch()
        D = self._n_dofs
        return self._stage_f[:, D + 4:]

    def get_contacts(self, with_entity=None, exclude_self_contact=False):
        self._fetch()
        Cn = self.N_CONTACT_SLOTS
        return {"link_a": self._stage_i[:, :Cn], "link_b": self._stage_i[:, Cn:], "valid_mask": self._stage_v}

    def control_dofs_position(self, position, dofs_idx_local=None):
        self._target = position
        self._host_action.copy_(position, non_blocking=True)

    def set_state_masked(self, mask, qpos, qvel):
        super().set_state_masked(mask, qpos, qvel)
        self._fetch()
        D = self._n_dofs
        m = mask.unsqueeze(-1)
        s = self._stage_f
        s[:, 0:3] = torch.where(m, qpos[:, 0:3], s[:, 0:3])
        s[:, 6:D] = torch.where(m, qpos[:, 7:], s[:, 6:D])
        s[:, D:D + 4] = torch.where(m, qpos[:, 3:7], s[:, D:D + 4])
        s[:, D + 4:] = torch.where(m, qvel, s[:, D + 4:])

    def set_qpos(self, qpos, envs_idx=None):
        super().set_qpos(qpos, envs_idx)
        self._park()

    def set_dofs_velocity(self, velocity, envs_idx=None, dofs_idx_local=None):
        super().set_dofs_velocity(velocity, envs_idx, dofs_idx_local)
        self._park()


class HostBoundaryScene(SyntheticScene):
    def add_entity(self, morph_type, morph_file=None, **kwargs):
        if morph_type == "plane":
            return super().add_entity(morph_type, morph_file=morph_file, **kwargs)
        ent = HostBoundaryEntity(self, morph_file, self._n_links)
        self._n_links += len(ent.links)
        self._robots.append(ent)
        self._entities.append(ent)
        return ent

    def step(self):
        super().step()
        self._robots[0]._park()


class HostBoundaryEngine(SyntheticEngine):
    def create_scene(self, show_viewer=False, **options):
        return HostBoundaryScene(self)
