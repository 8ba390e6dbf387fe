"""Environment shell and robot adaptor (the boundary the hot path reads its tensors through).

Restates what the ADD plugins consume from the reference's ``Environment`` /
``ImitationEnvironment`` (add_gym/envs/env.py:8-190) and ``Manipulator`` (add_gym/robot.py:13-330):
``env.num_envs``, ``env.ctrl_dt``, ``env.time_buf`` (fp32 [N], advanced by ``+= ctrl_dt`` after every
physics step, env.py:150-155), ``env.reset(ids)`` (``time_buf[ids] = 0``, env.py:157-173),
``env.plane``, ``env.robot.{base_pos, base_quat, base_lin_vel, base_ang_vel, dof_pos, dof_vel}``
(robot.py:271-293), ``robot.get_action_space()`` (1.4 x joint-limit box, robot.py:183-212) and
``robot.get_ground_contact_forces_v2`` (robot.py:221-231).  Rendering, video and PD gains are not
part of the hot path and are left out.
"""
import importlib

import numpy as np
import torch

from . import kinematics


def _instantiate(engine_cfg):
    """Stand-in for hydra.utils.instantiate on the engine sub-config (env.py:35)."""
    target = engine_cfg["_target_"]
    mod, cls = target.rsplit(".", 1)
    kwargs = {k: v for k, v in engine_cfg.items() if k != "_target_"}
    return getattr(importlib.import_module(mod), cls)(**kwargs)


class Manipulator:
    def __init__(self, num_envs, scene, engine, robot_cfg, env_spacing=2.0, enable_ref=False, device="cpu"):
        self._device = device
        self._scene = scene
        self._engine = engine
        self._num_envs = num_envs
        self._args = robot_cfg
        morph_path = robot_cfg["urdf_path"]
        self._kin_char_model = kinematics.KinCharModel(device)
        self._kin_char_model.load_char_file(morph_path)
        self._robot_entity = scene.add_entity(
            morph_type="urdf" if morph_path.endswith(".urdf") else "mjcf",
            morph_file=morph_path, morph_pos=(0.0, 0.0, 0.0), morph_quat=(1.0, 0.0, 0.0, 0.0),
            material_type="rigid")
        self._ref_entity = None
        limits = []
        for joint in self._robot_entity.joints:
            limits.extend(joint.dofs_limit)
        self.joint_limits = torch.tensor(limits, device=engine.device, dtype=engine.tc_float)
        base = []
        for joint in self._robot_entity.joints:
            if joint.name in ("root_joint", "floating_base_joint"):
                base.extend(joint.dofs_idx)
        self.non_root_joints = [i for i in range(self._robot_entity.n_dofs) if i not in base]

    def on_build(self):
        pass

    def get_action_space(self):
        """[D,2] (low, high): joint-limit midpoint +- 1.4 x half-range (robot.py:183-212)."""
        lim = self.joint_limits[self.non_root_joints].to("cpu", torch.float32)
        lo, hi = lim[:, 0], lim[:, 1]
        mid = 0.5 * (hi + lo)
        scale = torch.maximum(torch.abs(hi - mid), torch.abs(lo - mid)) * 1.4
        return torch.stack([mid - scale, mid + scale], dim=1)

    def get_ground_contact_forces_v2(self, surface_plane, contact_idx):
        c = self._robot_entity.get_contacts(with_entity=surface_plane, exclude_self_contact=True)
        ia = torch.isin(c["link_a"], contact_idx) & c["valid_mask"]
        ib = torch.isin(c["link_b"], contact_idx) & c["valid_mask"]
        return ia.any(dim=1) | ib.any(dim=1)

    def apply_action(self, action, allowed_action_idx=None):
        self._robot_entity.control_dofs_position(
            position=action, dofs_idx_local=allowed_action_idx if allowed_action_idx else self.non_root_joints)

    base_pos = property(lambda s: s._robot_entity.get_pos())
    base_quat = property(lambda s: s._robot_entity.get_quat())
    base_lin_vel = property(lambda s: s._robot_entity.get_vel())
    base_ang_vel = property(lambda s: s._robot_entity.get_ang())
    dof_pos = property(lambda s: s._robot_entity.get_dofs_position()[:, 6:])
    dof_vel = property(lambda s: s._robot_entity.get_dofs_velocity()[:, 6:])
    entity = property(lambda s: s._robot_entity)
    ref_entity = property(lambda s: s._ref_entity)


class Environment:
    def __init__(self, config, device):
        self.device = device
        self.env_cfg = config
        self.engine_cfg = config["engine"]
        self.robot_cfg = config["robot"]
        self.task_cfg = config["task"]
        self.ctrl_dt = self.engine_cfg["ctrl_dt"]
        engine = self.engine_cfg.get("_instance_")
        self.engine = engine if engine is not None else _instantiate(
            {**{k: v for k, v in self.engine_cfg.items() if k in ("_target_", "num_envs", "ctrl_dt", "seed",
                                                                    "noise_device", "fall_prob")},
             "device": device})
        self.engine.init(backend="gpu" if str(device).startswith("cuda") else "cpu", precision="32")
        self.scene = self.engine.create_scene(show_viewer=False, sim_options={"dt": self.ctrl_dt})
        self.plane = self.scene.add_entity(morph_type="plane")
        self.robot = Manipulator(self.engine_cfg["num_envs"], self.scene, self.engine, self.robot_cfg,
                                 env_spacing=self.engine_cfg.get("env_spacing", 2.0), device=device)
        sp = self.engine_cfg.get("env_spacing", 2.0)
        self.scene.build(n_envs=self.engine_cfg["num_envs"], env_spacing=(sp, sp))
        self.robot.on_build()
        self.num_envs = self.engine_cfg["num_envs"]
        self.time_buf = torch.zeros(self.num_envs, device=self.engine.device, dtype=torch.float32)
        self.extras = dict()

    def set_mode(self, mode):
        # the reference shrinks num_envs to 1 in TEST mode (env.py:142-148); train mode restores it
        self.num_envs = self.engine_cfg["num_envs"] if getattr(mode, "value", mode) == 0 else 1

    def step(self, actions):
        self.robot.apply_action(actions)
        self.scene.step()
        self.time_buf += self.ctrl_dt

    def reset_idx(self, envs_idx):
        if len(envs_idx) == 0:
            return
        self.time_buf[envs_idx] = 0

    def reset(self, env_ids=None):
        if env_ids is None:
            env_ids = torch.arange(self.num_envs, device=self.device, dtype=torch.long)
        self.reset_idx(env_ids)


class ImitationEnvironment(Environment):
    def __init__(self, config, device):
        super().__init__(config, device)
        self._diagnostics = {}

    def get_reward_succ(self):
        return 0.0

    def get_reward_fail(self):
        return 0.0

    def get_diagnostics(self):
        return self._diagnostics
