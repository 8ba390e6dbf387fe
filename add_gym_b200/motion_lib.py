"""MotionLib: the reference-motion step table in HBM and its lookups.

Drop-in for the reference's ``add_gym.anim.motion_lib.MotionLib`` as far as the hot path uses it
(motion_lib.py:18-59,285-335): same constructor, same query methods, same index arithmetic -- but the
30 fps -> 100 Hz resampling (root lerp, root / joint slerp, twist-angle extraction, finite-difference
velocities) runs as CUDA kernels (csrc/motion.cu) and the result is ONE packed table
``[S_total, 72]`` instead of seven tensors.

Reference quirk Q2 is reproduced by default: the start row of clip ``m`` is the cumulative sum of the
30 fps *frame* counts (motion_lib.py:280-282) although the table is sampled at 100 Hz, so for every
clip but the first the lookup lands in an earlier clip's rows.  ``fix_start_idx=True`` uses the true row
offsets instead (opt-in; changes results versus the reference).
"""
import ctypes as C
import os

import numpy as np
import torch
from . import _lib, motion_io

ROW_STRIDE_ALIGN = 4


def arange_count(length_f32, dt):
    """Number of elements of torch.arange(0, tensor(length, float32), dt): ceil((end-start)/step) in double."""
    return int(np.ceil(float(np.float32(length_f32)) / dt))


class MotionLib:
    def __init__(self, motion_file, motion_order, kin_char_model, dt, device, fix_start_idx=False):
        self._device = torch.device(device)
        self._kin_char_model = kin_char_model
        self._dt = dt
        self._dt_inv = round(1 / dt)
        self._fix_start_idx = fix_start_idx
        self._load_motions(motion_file, list(motion_order))

    # ---- loading ----------------------------------------------------------------------------------------
    def _fetch_motion_files(self, motion_file):
        """YAML library (motion_lib.py:337-358), clip pack (all clips, weight 1.0) or a single clip."""
        return motion_io.fetch_motion_files(motion_file)

    def _load_baked(self, path):
        """A pre-baked step table (`save_table`): upload and go -- no text parse, no resampling kernels."""
        header, table = motion_io.load_step_table(path)
        kin, dev = self._kin_char_model, self._device
        D = kin.get_dof_size()
        if int(header["num_dofs"]) != D or int(header["row_stride"]) != 2 * ((7 + D + 3) & ~3):
            raise ValueError("%s was baked for %s dofs / row stride %s" % (path, header["num_dofs"], header["row_stride"]))
        if abs(float(header["dt"]) - self._dt) > 1e-12:
            raise ValueError("%s was baked at dt = %s, the task runs at %s" % (path, header["dt"], self._dt))
        self._num_dofs, self._row_stride = D, int(header["row_stride"])
        self._motion_files = list(header["files"])
        # already normalised at bake time (float32 values, exact through the JSON header): dividing again would move an ulp
        self._motion_weights = torch.tensor(header["weights"], dtype=torch.float32).to(dev)
        self._motion_fps = torch.tensor(header["fps"], dtype=torch.float32, device=dev)
        self._motion_num_frames = torch.tensor(header["num_frames"], dtype=torch.long, device=dev)
        lengths_f32 = torch.tensor(header["lengths"], dtype=torch.float32)
        self._motion_lengths_host = lengths_f32.clone()
        self._motion_lengths = lengths_f32.to(dev)
        self._motion_loop_modes = torch.tensor(header["loop_modes"], dtype=torch.int, device=dev)
        n_steps = [int(v) for v in header["num_steps"]]
        self._motion_num_steps = n_steps
        true_start = np.concatenate([[0], np.cumsum(n_steps)[:-1]]).astype(np.int64)
        quirk_start = np.concatenate([[0], np.cumsum(header["num_frames"])[:-1]]).astype(np.int64)
        self._true_start_idx = torch.tensor(true_start, dtype=torch.long, device=dev)
        self._motion_start_idx = self._true_start_idx if self._fix_start_idx else torch.tensor(
            quirk_start, dtype=torch.long, device=dev)
        self._s_total = int(header["s_total"])
        self._table = torch.from_numpy(np.array(table, copy=True)).to(dev)
        self._frame_idx = None          # build-time diagnostics do not exist for a baked table
        self._frame_joint_rot, self._frame_vel = [], []
        self._frames_cat = None         # a baked table carries no 30 fps source frames (calc_motion_frame refuses)
        self._c_lib = _lib.AddkMotionLib(
            table=self._table.data_ptr(), row_stride=self._row_stride, num_motions=len(n_steps), s_total=self._s_total,
            start_idx=self._motion_start_idx.data_ptr(), lengths=self._motion_lengths.data_ptr(),
            loop_modes=self._motion_loop_modes.data_ptr())

    def save_table(self, path):
        """Bake the library: the packed 100 Hz table as it sits in HBM + the per-clip metadata (motion_io.save_step_table).
        `MotionLib(path_to_addkt, ...)` then skips the text parse and the table build."""
        header = dict(row_stride=self._row_stride, num_dofs=self._num_dofs, dt=self._dt, s_total=self._s_total,
                      files=[os.path.basename(f) for f in self._motion_files],
                      weights=[float(v) for v in self._motion_weights.cpu().tolist()],
                      fps=[float(v) for v in self._motion_fps.cpu().tolist()],
                      num_frames=[int(v) for v in self._motion_num_frames.cpu().tolist()],
                      lengths=[float(v) for v in self._motion_lengths_host.tolist()],
                      loop_modes=[int(v) for v in self._motion_loop_modes.cpu().tolist()],
                      num_steps=[int(v) for v in self._motion_num_steps])
        motion_io.save_step_table(path, header, self._table.cpu().numpy())

    def _load_motions(self, motion_file, motion_order):
        if motion_file.endswith(".addkt"):
            return self._load_baked(motion_file)
        L = _lib.lib()
        kin = self._kin_char_model
        D = kin.get_dof_size()
        self._num_dofs = D
        half = (7 + D + 3) & ~3
        self._row_stride = 2 * half
        files, weights = self._fetch_motion_files(motion_file)
        clips = [motion_io.load_motion(f) for f in files]
        fps = [c.fps for c in clips]
        nframes = [c.frames.shape[0] for c in clips]
        lengths = [1.0 / c.fps * (c.frames.shape[0] - 1) for c in clips]
        dev = self._device
        self._motion_files = files
        self._motion_weights = torch.tensor(weights, dtype=torch.float32)
        self._motion_weights /= self._motion_weights.sum()
        self._motion_weights = self._motion_weights.to(dev)
        self._motion_fps = torch.tensor(fps, dtype=torch.float32, device=dev)
        self._motion_num_frames = torch.tensor(nframes, dtype=torch.long, device=dev)
        lengths_f32 = torch.tensor(lengths, dtype=torch.float32)
        self._motion_lengths_host = lengths_f32.clone()
        self._motion_lengths = lengths_f32.to(dev)
        self._motion_loop_modes = torch.tensor([c.loop_mode.value for c in clips], dtype=torch.int, device=dev)
        n_steps = [arange_count(l, self._dt) for l in lengths_f32.tolist()]
        self._motion_num_steps = n_steps
        true_start = np.concatenate([[0], np.cumsum(n_steps)[:-1]]).astype(np.int64)
        quirk_start = np.concatenate([[0], np.cumsum(nframes)[:-1]]).astype(np.int64)
        self._true_start_idx = torch.tensor(true_start, dtype=torch.long, device=dev)
        self._motion_start_idx = self._true_start_idx if self._fix_start_idx else torch.tensor(
            quirk_start, dtype=torch.long, device=dev)
        s_total = int(np.sum(n_steps))
        self._s_total = s_total
        self._table = torch.zeros(s_total, self._row_stride, dtype=torch.float32, device=dev)
        # kinematic constants for the kernels
        self._dof_axis = torch.tensor(kin.dof_axes(), dtype=torch.float32, device=dev).contiguous()
        self._col_of_dof = torch.tensor(kin.motion_column_of_dof(motion_order), dtype=torch.int32, device=dev)
        self._frame_idx = torch.zeros(s_total, 2, dtype=torch.long, device=dev)
        self._frame_joint_rot = []      # per clip [F, D, 4]: 30 fps joint rotations (reference `_frame_joint_rot`)
        self._frame_vel = []            # per clip [F, 6 + D]: root_vel, root_ang_vel, dof_vel of the source frames
        # the 30 fps source data of all clips, concatenated (calc_motion_frame at arbitrary times reads them)
        tot_f = int(np.sum(nframes))
        f_start = np.concatenate([[0], np.cumsum(nframes)[:-1]]).astype(np.int64)
        self._frames_cat = torch.empty(tot_f, 7 + D, dtype=torch.float32, device=dev)
        self._jrot_cat = torch.empty(tot_f, D, 4, dtype=torch.float32, device=dev)
        self._fvel_cat = torch.empty(tot_f, 6 + D, dtype=torch.float32, device=dev)
        self._frame_start = torch.tensor(f_start, dtype=torch.long, device=dev)
        for m, clip in enumerate(clips):
            F = clip.frames.shape[0]
            assert clip.frames.shape[1] == 7 + D
            a, b = int(f_start[m]), int(f_start[m]) + F
            frames = self._frames_cat[a:b]
            frames.copy_(torch.tensor(clip.frames, dtype=torch.float32))                     # fp64 -> fp32 rounding
            jrot = self._jrot_cat[a:b]
            fvel = self._fvel_cat[a:b]
            rc = L.addk_motion_table_build(
                _lib.stream(), _lib.ptr(frames), C.c_int(F), C.c_int(D), _lib.ptr(self._col_of_dof),
                _lib.ptr(self._dof_axis), C.c_float(float(clip.fps)), C.c_float(float(np.float32(1.0 / clip.fps))),
                C.c_int(n_steps[m]), C.c_double(self._dt), C.c_float(float(lengths_f32[m])),
                C.c_int(1 if clip.loop_mode == motion_io.LoopMode.WRAP else 0), _lib.ptr(jrot), _lib.ptr(fvel),
                _lib.ptr(self._table), C.c_int(self._row_stride), C.c_longlong(int(true_start[m])), C.c_void_p(0),
                _lib.ptr(self._frame_idx))
            _lib.check(rc, "addk_motion_table_build")
            self._frame_joint_rot.append(jrot)
            self._frame_vel.append(fvel)
        torch.cuda.synchronize(dev)
        self._c_lib = _lib.AddkMotionLib(
            table=self._table.data_ptr(), row_stride=self._row_stride, num_motions=len(clips), s_total=s_total,
            start_idx=self._motion_start_idx.data_ptr(), lengths=self._motion_lengths.data_ptr(),
            loop_modes=self._motion_loop_modes.data_ptr())

    # ---- queries (reference names) ----------------------------------------------------------------------
    def get_num_motions(self):
        return self._motion_lengths.shape[0]

    def get_total_length(self):
        return torch.sum(self._motion_lengths).item()

    def get_motion_length(self, motion_ids):
        return self._motion_lengths[motion_ids]

    def get_motion_lengths(self):
        return self._motion_lengths

    def get_motion_weights(self):
        return self._motion_weights

    def get_motion_loop_mode(self, motion_ids):
        return self._motion_loop_modes[motion_ids]

    def sample_motions(self, n):
        return torch.multinomial(self._motion_weights, num_samples=n, replacement=True)

    def sample_time(self, motion_ids):
        """Uniform start time per clip, floored to the control step (motion_lib.py:40-46).  The ADD path samples through
        ADDMotion.sample_time (segment sampler) instead; kept for callers of the plain library."""
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_time = phase * self._motion_lengths[motion_ids]
        return (motion_time // self._dt) * self._dt

    def calc_motion_phase(self, motion_ids, times):
        length = self._motion_lengths[motion_ids]
        phase = times / length
        wrap = self._motion_loop_modes[motion_ids] == motion_io.LoopMode.WRAP.value
        phase = torch.where(wrap, phase - torch.floor(phase), phase)
        return torch.clip(phase, 0.0, 1.0)

    def calc_motion_frame(self, motion_ids, motion_times):
        """Reference MotionLib.calc_motion_frame (motion_lib.py:61-88): interpolate the 30 fps clips at ARBITRARY times --
        root lerp, root / joint slerp, hinge angle = twist angle of the blended joint rotation, velocities of frame i0,
        loop offset for WRAP clips.  -> (root_pos [n,3], root_rot [n,4] wxyz, root_vel [n,3], root_ang_vel [n,3],
        joint_rot [n,D,4], dof_pos [n,D], dof_vel [n,D]).  The hot path itself never calls this (it reads the 100 Hz
        table, quirk Q1); the table builder applies the same arithmetic at the grid times."""
        if self._frames_cat is None:
            raise _lib.AddkError("calc_motion_frame needs the 30 fps source clips: this library was loaded from a baked step table")
        n, D, dev = int(motion_ids.shape[0]), self._num_dofs, self._device
        ids = motion_ids.to(torch.long).contiguous()
        times = motion_times.to(torch.float32).contiguous()
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        out = (f(n, 3), f(n, 4), f(n, 3), f(n, 3), f(n, D, 4), f(n, D), f(n, D))
        rc = _lib.lib().addk_motion_frame(
            _lib.stream(), _lib.ptr(self._frames_cat), _lib.ptr(self._jrot_cat), _lib.ptr(self._fvel_cat),
            _lib.ptr(self._frame_start), _lib.ptr(self._motion_num_frames), _lib.ptr(self._motion_lengths),
            _lib.ptr(self._motion_loop_modes), C.c_int(D), _lib.ptr(self._dof_axis), _lib.ptr(ids), _lib.ptr(times),
            C.c_int(n), *[_lib.ptr(t) if n > 0 else C.c_void_p(0) for t in out])
        _lib.check(rc, "addk_motion_frame")
        return out

    def get_precomputed_motion_step(self, motion_ids, motion_times, return_index=False):
        """(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel) rows of the step table."""
        n = motion_ids.shape[0]
        dev, D = self._device, self._num_dofs
        ids = motion_ids.to(torch.long).contiguous()
        times = motion_times.to(torch.float32).contiguous()
        out = [torch.empty(n, w, dtype=torch.float32, device=dev) for w in (3, 4, 3, 3, D, D)]
        idx = torch.empty(n, dtype=torch.long, device=dev) if return_index else None
        rc = _lib.lib().addk_motion_gather(
            _lib.stream(), _lib.ptr(self._table), C.c_int(self._row_stride), C.c_int(D), C.c_longlong(self._s_total),
            _lib.ptr(self._motion_start_idx), C.c_float(float(self._dt_inv)), _lib.ptr(ids), _lib.ptr(times), C.c_int(n),
            *[_lib.ptr(o) for o in out], _lib.ptr(idx))
        _lib.check(rc, "addk_motion_gather")
        return tuple(out) + ((idx,) if return_index else ())

    # ---- views of the packed table under the reference attribute names -----------------------------------
    @property
    def step_table(self):
        return self._table

    @property
    def _step_root_pos(self):
        return self._table[:, 0:3]

    @property
    def _step_root_rot(self):
        return self._table[:, 3:7]

    @property
    def _step_dof_pos(self):
        return self._table[:, 7:7 + self._num_dofs]

    @property
    def _step_root_vel(self):
        h = self._row_stride // 2
        return self._table[:, h:h + 3]

    @property
    def _step_root_ang_vel(self):
        h = self._row_stride // 2
        return self._table[:, h + 3:h + 6]

    @property
    def _step_dof_vel(self):
        h = self._row_stride // 2
        return self._table[:, h + 6:h + 6 + self._num_dofs]
