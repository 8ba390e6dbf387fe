"""ADDObservation / ADDReward / ADDDone, drop-in for the reference task plugins
(add_gym/learning/add/add_observation.py:11-419, add_reward.py:5-89, add_done.py:12-93).

Same constructors, public methods and in-place mutated tensors (``obs_buf``, ``info["disc_obs"]``,
``info["disc_obs_demo"]``, ``_motion_ids``, ``_motion_time_offsets``, ``ref_*``, ``done_buf``); the work
is one fused CUDA kernel (csrc/step.cu) that the three classes share through ``StepCore``.  Calling the
plugins one by one, as the reference agent does, launches that kernel with the matching flag subset;
``ADDAgent`` instead calls ``StepCore.step`` once per env step.
"""
import ctypes as C
import enum

import torch

from . import _lib
from .add_motion import ADDMotion


class DoneFlags(enum.Enum):  # base_agent.py:16-20
    NULL = 0
    FAIL = 1
    SUCC = 2
    TIME = 3


def _obs_dims(cfg, D):
    vel = cfg.get("enable_vel_obs", False)
    char = (1 if cfg.get("root_height_obs", False) else 0) + 6 + D + ((6 + D) if vel else 0)
    phase = (1 + 2 * cfg.get("num_phase_encoding", 0)) if cfg.get("enable_phase_obs", True) else 0
    n_tar = len(cfg.get("tar_obs_steps", [1])) if cfg.get("enable_tar_obs", False) else 0
    tar = n_tar * ((3 if cfg.get("root_height_obs", False) else 2) + 6 + D)
    disc = cfg["num_disc_obs_steps"] * (9 + D + ((6 + D) if vel else 0))
    return char + phase + tar, disc


class StepCore:
    """Owns the persistent per-env tensors and the C structs handed to addk_env_step / addk_reset_done."""

    def __init__(self, config, env, motion, device):
        self.env, self.motion, self.device, self.config = env, motion, torch.device(device), config
        lib = motion.motion_lib
        N, D = env.num_envs, lib._num_dofs
        self.N, self.D = N, D
        dev = self.device
        obs_dim, disc_dim = _obs_dims(config, D)
        self.obs_dim, self.disc_dim = obs_dim, disc_dim
        dt = env.ctrl_dt
        tar_steps = list(config.get("tar_obs_steps", [1]))
        enable_tar = bool(config.get("enable_tar_obs", False))
        nH = int(config["num_disc_obs_steps"])
        # offsets computed with torch exactly as the reference does (add_observation.py:214-215, 362-370)
        tar_off = (dt * torch.tensor(tar_steps, dtype=torch.int)).to(torch.float32)
        disc_off = torch.flip(-dt * torch.arange(0, nH), dims=[0]).to(torch.float32)
        assert len(tar_steps) <= _lib.ADDK_MAX_TAR_STEPS and nH <= _lib.ADDK_MAX_DISC_STEPS
        t = _lib.AddkTask()
        t.num_dofs, t.num_tar_steps, t.num_disc_steps = D, (len(tar_steps) if enable_tar else 0), nH
        t.global_obs = int(config.get("global_obs", False))
        t.root_height_obs = int(config.get("root_height_obs", False))
        t.enable_vel_obs = int(config.get("enable_vel_obs", False))
        t.enable_phase_obs = int(config.get("enable_phase_obs", True))
        t.enable_tar_obs = int(enable_tar)
        t.num_phase_encoding = int(config.get("num_phase_encoding", 0))
        t.obs_dim, t.disc_obs_dim = obs_dim, disc_dim
        t.track_root = int(enable_tar and config.get("global_obs", False))   # _track_global_root()
        t.track_root_h = int(config.get("root_height_obs", False))
        t.enable_early_termination = int(config["enable_early_termination"])
        t.pose_termination = int(config.get("pose_termination", False))
        for i, v in enumerate(tar_off.tolist()):
            t.tar_offsets[i] = v
        for i, v in enumerate(disc_off.tolist()):
            t.disc_offsets[i] = v
        t.ctrl_dt = dt
        t.dt_inv = float(lib._dt_inv)
        for k in ("pose_w", "vel_w", "root_pose_w", "root_vel_w", "pose_scale", "vel_scale", "root_pose_scale",
                  "root_vel_scale"):
            v = config.get("reward_" + k)
            setattr(t, k, 0.0 if v is None else float(v))
        t.ep_len = float(config.get("max_episode_length", lib.get_total_length()))
        t.pose_termination_dist = float(config.get("pose_termination_dist", 1.0))
        self.task = t
        self.c_lib = lib._c_lib
        # persistent tensors
        f32 = dict(dtype=torch.float32, device=dev)
        self.motion_ids = torch.zeros(N, dtype=torch.int64, device=dev)
        self.motion_time_offsets = torch.zeros(N, **f32)
        self.ref_root_pos = torch.zeros(N, 3, **f32)
        self.ref_root_rot = torch.zeros(N, 4, **f32)
        self.ref_root_vel = torch.zeros(N, 3, **f32)
        self.ref_root_ang_vel = torch.zeros(N, 3, **f32)
        self.ref_dof_pos = torch.zeros(N, D, **f32)
        self.ref_dof_vel = torch.zeros(N, D, **f32)
        self.hist = torch.zeros(N, nH, lib._row_stride, **f32)
        self.hist_head = 0                      # CircularBuffer._head (shared by the six reference rings)
        self.obs_buf = torch.zeros(N, obs_dim, **f32)
        self.disc_obs = torch.zeros(N, disc_dim, **f32)
        self.disc_obs_demo = torch.zeros(N, disc_dim, **f32)
        self.reward = torch.zeros(N, **f32)
        self.done_buf = torch.zeros(N, dtype=torch.int32, device=dev)
        self.return_buf = torch.zeros(N, **f32)
        self.ep_len_buf = torch.zeros(N, dtype=torch.int64, device=dev)
        self.eps_per_env = torch.zeros(N, dtype=torch.int64, device=dev)
        self.tracker_sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self.tracker_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.dof_err_w = torch.ones(D, **f32)
        self.new_ids = torch.zeros(N, dtype=torch.int64, device=dev)
        self.new_times = torch.zeros(N, **f32)
        self.qpos_out = torch.zeros(N, 7 + D, **f32)
        self.qvel_out = torch.zeros(N, 6 + D, **f32)
        self.reset_mask = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._noncontact_mask = 0
        self.ground_plane = None
        self._env_struct(with_tracker=True)
        self._env_struct(with_tracker=False)

    # ---- struct assembly ------------------------------------------------------------------------------
    def _env_struct(self, with_tracker):
        e = _lib.AddkEnvBuffers()
        e.time_buf = self.env.time_buf.data_ptr()
        e.motion_ids = self.motion_ids.data_ptr()
        e.motion_time_offsets = self.motion_time_offsets.data_ptr()
        for k in ("ref_root_pos", "ref_root_rot", "ref_root_vel", "ref_root_ang_vel", "ref_dof_pos", "ref_dof_vel"):
            setattr(e, k, getattr(self, k).data_ptr())
        e.hist = self.hist.data_ptr()
        e.hist_stride = self.hist.shape[2]
        e.obs_buf = self.obs_buf.data_ptr()
        e.disc_obs = self.disc_obs.data_ptr()
        e.disc_obs_demo = self.disc_obs_demo.data_ptr()
        e.reward = self.reward.data_ptr()
        e.done = self.done_buf.data_ptr()
        if with_tracker:
            e.return_buf = self.return_buf.data_ptr()
            e.ep_len_buf = self.ep_len_buf.data_ptr()
            e.eps_per_env = self.eps_per_env.data_ptr()
            e.tracker_sums = self.tracker_sums.data_ptr()
            e.tracker_count = self.tracker_count.data_ptr()
            self.c_env = e
        else:
            self.c_env_notrack = e
        return e

    def sim_struct(self):
        """Pointers + leading dims of the engine getters (robot.py:271-293); re-read every call because an
        engine may hand out fresh tensors each step."""
        r = self.env.robot
        s = _lib.AddkSimState()

        def put(name, t):
            if t.dtype != torch.float32 or t.stride(-1) != 1:
                t = t.to(torch.float32).contiguous()
            self._keep.append(t)
            setattr(s, name, t.data_ptr())
            setattr(s, "ld_" + name, t.stride(0))

        self._keep = []
        put("root_pos", r.base_pos)
        put("root_rot", r.base_quat)
        put("root_vel", r.base_lin_vel)
        put("root_ang", r.base_ang_vel)
        put("dof_pos", r.dof_pos)
        put("dof_vel", r.dof_vel)
        # contacts are fetched EVERY step with the width the engine reports for this step: MJWarp returns [N, 0] while
        # nothing touches (and at build time), Genesis pads to the per-step maximum (robot.py:221-231)
        s.contact_slots, s.ld_contact = 0, 0
        link_masks = getattr(r.entity, "get_contact_link_masks", None)
        if self.ground_plane is not None and self.task.enable_early_termination and link_masks is not None:
            # engine extension: per-env {self, other} link bitmasks straight from the backend's flat contact arrays
            # (addk_contact_link_mask) -- no padded list, no host round trip
            m = link_masks(with_entity=self.ground_plane, exclude_self_contact=True)
            assert m.dtype == torch.int64 and m.shape == (self.N, 2) and m.is_contiguous()
            self._keep.append(m)
            s.contact_link_masks = m.data_ptr()
        elif self.ground_plane is not None and self.task.enable_early_termination:
            c = r.entity.get_contacts(with_entity=self.ground_plane, exclude_self_contact=True)
            la, lb, va = c["link_a"], c["link_b"], c["valid_mask"]
            width = int(la.shape[1]) if la.dim() == 2 else 0
            if width > 0:
                if la.dtype != torch.int32:
                    la, lb = la.to(torch.int32), lb.to(torch.int32)
                if va.dtype not in (torch.bool, torch.uint8):
                    va = va != 0
                la, lb, va = la.contiguous(), lb.contiguous(), va.contiguous()
                self._keep += [la, lb, va]
                s.link_a, s.link_b, s.valid = la.data_ptr(), lb.data_ptr(), va.data_ptr()
                s.contact_slots, s.ld_contact = width, width
        return s

    def set_contact_model(self, ground_plane, noncontact_link_ids):
        self.ground_plane = ground_plane
        mask = 0
        for l in noncontact_link_ids:
            assert 0 <= l < 64, "link ids above 63 are not supported by the contact bitmask"
            mask |= (1 << l)
        self._noncontact_mask = mask
        self.task.noncontact_link_mask = mask
        self.task.contact_slots = 0          # (per-step width: sim_struct)

    # ---- launches ---------------------------------------------------------------------------------------
    def step(self, flags, exp_row=None, env_mask=None, track_returns=True):
        sim = self.sim_struct()
        if (flags & _lib.F_MASKED) and env_mask is None:
            raise _lib.AddkError("masked step needs env_mask")
        env = self.c_env if track_returns else self.c_env_notrack
        rc = _lib.lib().addk_env_step(
            _lib.stream(), C.byref(self.task), C.byref(self.c_lib), C.byref(sim), C.byref(env),
            C.byref(exp_row) if exp_row is not None else None, _lib.ptr(self.dof_err_w), _lib.ptr(env_mask),
            C.c_int(self.N), C.c_int(self.hist_head), C.c_int(flags))
        _lib.check(rc, "addk_env_step")
        if flags & _lib.F_UPDATE_MOTION:
            self.hist_head = (self.hist_head + 1) % self.hist.shape[1]

    def reset(self, flags_tensor, reset_all, zero_time_done):
        """flags_tensor int32[N]: envs with a non-zero entry are reset.  Candidates come from new_ids/new_times."""
        env = _lib.AddkEnvBuffers.from_buffer_copy(self.c_env_notrack)
        env.done = flags_tensor.data_ptr()
        rc = _lib.lib().addk_reset_done(
            _lib.stream(), C.byref(self.task), C.byref(self.c_lib), C.byref(env), _lib.ptr(self.new_ids),
            _lib.ptr(self.new_times), C.c_int(self.N), C.c_int(self.hist_head), C.c_int(1 if reset_all else 0),
            C.c_int(1 if zero_time_done else 0), _lib.ptr(self.qpos_out), _lib.ptr(self.qvel_out),
            _lib.ptr(self.reset_mask))
        _lib.check(rc, "addk_reset_done")


class ADDObservation:
    def __init__(self, config, env, motion: ADDMotion, device):
        self.env, self.device, self.config, self.motion = env, device, config, motion
        self.dt = env.ctrl_dt
        self._enable_phase_obs = config.get("enable_phase_obs", True)
        self._enable_tar_obs = config.get("enable_tar_obs", False)
        self._enable_vel_obs = config.get("enable_vel_obs", False)
        self._global_obs = config.get("global_obs", False)
        self._root_height_obs = config.get("root_height_obs", False)
        self._num_disc_obs_steps = config["num_disc_obs_steps"]
        self._num_phase_encoding = config.get("num_phase_encoding", 0)
        self.core = StepCore(config, env, motion, device)
        c = self.core
        self._motion_ids = c.motion_ids
        self._motion_time_offsets = c.motion_time_offsets
        self.ref_root_pos, self.ref_root_rot = c.ref_root_pos, c.ref_root_rot
        self.ref_root_vel, self.ref_root_ang_vel = c.ref_root_vel, c.ref_root_ang_vel
        self.ref_dof_pos, self.ref_dof_vel = c.ref_dof_pos, c.ref_dof_vel
        self.obs_buf = c.obs_buf
        self._disc_obs_buf, self._disc_obs_demo_buf = c.disc_obs, c.disc_obs_demo
        self.info = {"disc_obs": c.disc_obs, "disc_obs_demo": c.disc_obs_demo}

    def get_obs_shape(self):
        return torch.Size([self.core.obs_dim])

    def get_disc_obs_shape(self):
        return torch.Size([self.core.disc_dim])

    def get_disc_obs_space(self):
        space = torch.zeros((2, self.core.disc_dim), dtype=torch.float32)
        space[0, ...] = -torch.inf
        space[1, ...] = torch.inf
        return space

    def update_motion(self):
        self.core.step(_lib.F_UPDATE_MOTION, track_returns=False)

    def compute_obs(self):
        self.core.step(0, track_returns=False)
        return self.obs_buf

    def get_observations(self):
        return self.obs_buf

    def reset_idx(self, env_ids):
        n = len(env_ids)
        if n == 0:
            return
        c = self.core
        motion_ids, motion_times = self.motion.sample_time(n)
        c.new_ids[env_ids] = motion_ids
        c.new_times[env_ids] = motion_times
        flags = torch.zeros(c.N, dtype=torch.int32, device=c.device)
        flags[env_ids] = 1
        c.reset(flags, reset_all=False, zero_time_done=False)
        ent = self.env.robot.entity
        ent.set_qpos(c.qpos_out[env_ids], envs_idx=env_ids)
        ent.set_dofs_velocity(c.qvel_out[env_ids], envs_idx=env_ids)

    def _track_global_root(self):
        return self._enable_tar_obs and self._global_obs

    def _get_motion_times(self):
        return self.env.time_buf + self._motion_time_offsets

    def fetch_disc_obs_demo(self, num_samples):
        """Demo discriminator observations at freshly sampled (clip, time) pairs (add_observation.py:158-161).
        Off the hot path (the reference only uses it to infer shapes); composed from the table gather."""
        ids, t0 = self.motion.sample_time(num_samples)
        nH, D = self._num_disc_obs_steps, self.core.D
        offs = torch.flip(-self.dt * torch.arange(0, nH, device=t0.device), dims=[0])
        times = (t0.unsqueeze(-1) + offs).view(-1)
        idr = torch.tile(ids.unsqueeze(-1), [1, nH]).view(-1)
        pos, rot, vel, ang, dof, dofv = self.motion.get_motion_step(idr, times)
        w, v = rot[:, 0:1], rot[:, 1:]

        def rotate(vec):
            t = 2 * torch.cross(v, vec, dim=-1)
            return vec + w * t + torch.cross(v, t, dim=-1)
        ex = torch.zeros_like(pos); ex[:, 0] = 1
        ez = torch.zeros_like(pos); ez[:, 2] = 1
        if not self._global_obs:
            pos = pos.clone(); pos[:, 0:2] = 0.0
        parts = [pos, rotate(ex), rotate(ez), dof]
        if self._enable_vel_obs:
            assert self._global_obs, "fetch_disc_obs_demo: local-frame velocity obs only exist in the fused kernel"
            parts += [vel, ang, dofv]
        return torch.cat(parts, dim=-1).reshape(num_samples, -1)


class ADDReward:
    def __init__(self, config, env, add_obs: ADDObservation, device):
        self.env, self.add_obs, self.device, self.config = env, add_obs, device, config
        self._root_height_obs = config.get("root_height_obs", False)
        kin = env.robot._kin_char_model
        num_joints = kin.get_num_joints()
        joint_err_w = config.get("joint_err_w", None)
        if joint_err_w is None:
            self._joint_err_w = torch.ones(num_joints - 1, dtype=torch.float32)
        else:
            self._joint_err_w = torch.tensor(joint_err_w, dtype=torch.float32)
        assert self._joint_err_w.shape[-1] == num_joints - 1
        w = torch.zeros(kin.get_dof_size(), dtype=torch.float32)
        for j in range(1, num_joints):
            d = kin.get_joint_dof_dim(j)
            if d > 0:
                i = kin.get_joint_dof_idx(j)
                w[i:i + d] = self._joint_err_w[j - 1]
        self._dof_err_w = w.to(device)
        add_obs.core.dof_err_w.copy_(self._dof_err_w)

    def compute_reward(self):
        self.add_obs.core.step(_lib.F_REWARD_DONE, track_returns=False)
        return self.add_obs.core.reward


class ADDDone:
    def __init__(self, config, env, add_obs: ADDObservation, add_motion: ADDMotion, ground_plane, device):
        self.env, self.add_obs, self.add_motion, self.device, self.config = env, add_obs, add_motion, device, config
        self.ground_plane = ground_plane
        self._max_episode_length = config.get("max_episode_length", add_motion.motion_lib.get_total_length())
        self._enable_early_termination = config["enable_early_termination"]
        self._termination_height = config["termination_height"]
        self._pose_termination = config.get("pose_termination", False)
        self._pose_termination_dist = config.get("pose_termination_dist", 1.0)
        ent = env.robot.entity
        contact_ids = [ent.get_link(name=n).idx for n in config.get("contact_bodies", [])]
        self._contact_body_ids = torch.tensor(contact_ids, device=device, dtype=torch.long)
        noncontact = [l.idx for l in ent.links if l.idx not in contact_ids]
        self._noncontact_body_ids = torch.tensor(noncontact, device=device, dtype=torch.long)
        add_obs.core.set_contact_model(ground_plane, noncontact)
        self.done_buf = add_obs.core.done_buf

    def compute_done(self):
        self.add_obs.core.step(_lib.F_REWARD_DONE, track_returns=False)
        return self.done_buf

    def reset_idx(self, env_ids):
        self.done_buf[env_ids] = DoneFlags.NULL.value
