"""ADDMotion and the adaptive start-time sampler, drop-in for the reference plugins
``add_gym.learning.add.add_motion.ADDMotion`` (add_motion.py:12-61) and
``add_gym.learning.sampler.AdaptiveSegmentSampler`` (sampler.py:5-92).

The public methods keep the reference's names, argument meaning and return types.  On top of them the
B200 path adds a masked, host-sync-free variant (``sample_time_masked``): candidates are drawn for the
environments whose done flag is set, straight into per-env arrays, instead of first compacting the done
indices on the host (base_agent.py:449-453).
"""
import ctypes as C

import torch
import torch.nn.functional as Fn

from . import _lib
from .motion_lib import MotionLib


class AdaptiveSegmentSampler:
    def __init__(self, clip_lengths, dt, num_segments=20, temperature=None, min_start_time=0.0, device=None):
        self.num_segments = num_segments
        self.dt = dt
        self.temperature = temperature
        self.min_start_time = min_start_time
        lengths = clip_lengths.detach().to("cpu", torch.float32)
        dev = clip_lengths.device if device is None else torch.device(device)
        # fp32 division, exactly torch.tensor([len / num_segments for len in clip_lengths]) (sampler.py:13-15)
        self.segment_sizes = (lengths / num_segments).to(dev)
        self.errors = torch.ones((lengths.shape[0], num_segments), device=dev)
        self._sums = torch.zeros(lengths.shape[0] * num_segments, dtype=torch.float64, device=dev)
        self._counts = torch.zeros(lengths.shape[0] * num_segments, dtype=torch.int32, device=dev)
        self._temp_bits = torch.zeros(1, dtype=torch.int32, device=dev)

    @torch.no_grad()
    def update_errors(self, clip_ids, timesteps, tracking_errors, disc_obs_demo=None):
        """EMA of the per-(clip, segment) mean tracking error.  With ``disc_obs_demo`` given,
        ``tracking_errors`` is the [n, dim] agent disc-obs and sum((a-b)^2) is formed in the kernel."""
        n = clip_ids.shape[0]
        dim = tracking_errors.shape[-1] if disc_obs_demo is not None else 1
        rc = _lib.lib().addk_sampler_update_errors(
            _lib.stream(), _lib.ptr(clip_ids.contiguous()), _lib.ptr(timesteps.contiguous()),
            _lib.ptr(tracking_errors.contiguous()), _lib.ptr(disc_obs_demo), C.c_int(dim), C.c_int(n),
            _lib.ptr(self.segment_sizes), C.c_int(self.errors.shape[0]), C.c_int(self.num_segments),
            _lib.ptr(self._sums), _lib.ptr(self._counts), _lib.ptr(self.errors))
        _lib.check(rc, "addk_sampler_update_errors")

    def get_probs(self, clip_ids=None):
        if clip_ids is None:
            clip_ids = torch.arange(self.errors.shape[0], device=self.errors.device)
        clip_errors = self.errors[clip_ids]
        temperature = torch.max(clip_errors) + 1e-6 if self.temperature is None else self.temperature
        return Fn.softmax(clip_errors / temperature, dim=-1)

    def sample_masked(self, motion_weights, rand_reset, done, uniforms, ids_out, times_out):
        """Draw (clip, start time) for every env with done != 0 (all envs when ``done`` is None)."""
        n = ids_out.shape[0]
        rc = _lib.lib().addk_sample_motion_time(
            _lib.stream(), _lib.ptr(motion_weights), C.c_int(self.errors.shape[0]), _lib.ptr(self.errors),
            C.c_int(self.num_segments), _lib.ptr(self.segment_sizes), C.c_float(self.dt),
            C.c_float(self.min_start_time), C.c_float(-1.0 if self.temperature is None else float(self.temperature)),
            C.c_int(1 if rand_reset else 0), _lib.ptr(done), _lib.ptr(uniforms), C.c_int(n), _lib.ptr(self._temp_bits),
            _lib.ptr(ids_out), _lib.ptr(times_out))
        _lib.check(rc, "addk_sample_motion_time")

    def start_times_from_draws(self, clip_ids, segments, uniforms):
        """The arithmetic of sample_start_frame AFTER its two draws (sampler.py:84-92), for callers that keep the
        reference's own torch.multinomial / torch.rand calls: t = seg * size + U * size -> (t // dt) * dt ->
        clamp(min = min_start_time).  Bit-exact with the reference on identical draws."""
        n = clip_ids.shape[0]
        times = torch.empty(n, dtype=torch.float32, device=clip_ids.device)
        rc = _lib.lib().addk_start_time_from_draws(
            _lib.stream(), _lib.ptr(self.segment_sizes), C.c_float(self.dt), C.c_float(self.min_start_time),
            _lib.ptr(clip_ids.to(torch.long).contiguous()), _lib.ptr(segments.to(torch.long).contiguous()),
            _lib.ptr(uniforms.to(torch.float32).contiguous()), C.c_int(n), _lib.ptr(times))
        _lib.check(rc, "addk_start_time_from_draws")
        return times

    def sample_start_frame(self, clip_ids=None):
        """Reference signature: start times for the given clips (sampler.py:75-92)."""
        n = clip_ids.shape[0]
        dev = clip_ids.device
        # one-hot "weights" per call are not expressible; draw segments with the kernel by fixing the clip:
        # the clip draw is skipped by passing uniforms whose first column selects nothing new.
        u = torch.rand(n, 3, device=dev)
        ids = clip_ids.to(torch.long).contiguous()
        times = torch.empty(n, dtype=torch.float32, device=dev)
        L = _lib.lib()
        # temperature over the clips of this batch (sampler.py:66-69)
        temp = (torch.max(self.errors[ids]) + 1e-6).item() if self.temperature is None else float(self.temperature)
        rc = L.addk_sample_start_time(
            _lib.stream(), _lib.ptr(self.errors), C.c_int(self.num_segments), _lib.ptr(self.segment_sizes),
            C.c_float(self.dt), C.c_float(self.min_start_time), C.c_float(temp), _lib.ptr(u), C.c_int(n), _lib.ptr(ids),
            _lib.ptr(times))
        _lib.check(rc, "addk_sample_start_time")
        return times


class ADDMotion:
    def __init__(self, config, env, device, fix_start_idx=False):
        self.env = env
        self.device = device
        self.config = config
        self.motion_lib = MotionLib(
            motion_file=config["motion_file"], motion_order=list(config["motion_joint_order"]),
            kin_char_model=env.robot._kin_char_model, dt=self.env.ctrl_dt, device=device,
            fix_start_idx=fix_start_idx)
        sampler_config = config.get("sampler", {})
        num_disc_obs_steps = config.get("num_disc_obs_steps", 1)
        self.sampler = AdaptiveSegmentSampler(
            clip_lengths=self.motion_lib.get_motion_lengths(), dt=self.env.ctrl_dt,
            num_segments=sampler_config.get("num_segments", 20), temperature=sampler_config.get("temperature", None),
            min_start_time=(num_disc_obs_steps - 1) * self.env.ctrl_dt)
        self._rand_reset = config.get("rand_reset", True)

    def get_motion_step(self, motion_ids, motion_times):
        return self.motion_lib.get_precomputed_motion_step(motion_ids, motion_times)

    def get_motion_phase(self, motion_ids, motion_times):
        return self.motion_lib.calc_motion_phase(motion_ids, motion_times)

    def get_motion_length(self, motion_ids):
        return self.motion_lib.get_motion_length(motion_ids)

    def get_motion_loop_mode(self, motion_ids):
        return self.motion_lib.get_motion_loop_mode(motion_ids)

    def sample_motions(self, n):
        return self.motion_lib.sample_motions(n)

    def sample_time(self, n):
        motion_ids = self.sample_motions(n)
        if self._rand_reset:
            motion_times = self.sampler.sample_start_frame(motion_ids)
        else:
            motion_times = torch.zeros(n, dtype=torch.float, device=self.device)
        return motion_ids, motion_times

    def sample_time_masked(self, done, ids_out, times_out, uniforms=None):
        """Fill ids_out/times_out where done != 0 (everywhere if done is None); no host sync."""
        n = ids_out.shape[0]
        if uniforms is None:
            uniforms = torch.rand(n, 3, device=ids_out.device)
        self.sampler.sample_masked(self.motion_lib.get_motion_weights(), self._rand_reset, done, uniforms, ids_out,
                                   times_out)
