"""CPU oracle for the add-gym rollout + update hot path.  TEST INFRASTRUCTURE -- NOT THE PRODUCT.

A restatement, in plain torch-on-CPU fp32, of what the reference (rsamf/add-gym, Python/PyTorch) computes
on this path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import it, and only as the checker / the timed CPU baseline.  The product (add_gym_b200/) never does.

Pinning: the reference ships no tests or golden vectors, so this port is pinned against the
EXECUTED reference: tests/golden/make_golden.py runs the real `ADDAgent` from /root/reference (through
oracle/ref_harness.py) and stores its tensors in tests/golden/*.npz; tests/test_oracle_golden.py checks
this file against them on every machine, and tests/test_oracle_vs_reference.py re-runs the live
comparison where /root/reference exists.

Third-party arithmetic: torch (pinned torch==2.8.0 in the reference's uv.lock, 2.11.0 here) supplies the
reference's GEMMs, autograd, AdamW, BCE-with-logits and RNG; the oracle uses the same library calls for
those (`torch.nn.functional.linear`, `torch.autograd`), and restates AdamW explicitly (checked against
torch.optim.AdamW in tests).

Each function cites the reference lines it follows (paths relative to /root/reference/add_gym/).
"""
import math

import numpy as np
import torch

NULL, FAIL, SUCC, TIME = 0, 1, 2, 3          # learning/base_agent.py:16-20
CLAMP, WRAP = 0, 1                            # anim/motion.py:6-8


# ==================================================================================================
# quaternions (wxyz)                                                        util/torch_util.py:34-406
# ==================================================================================================
def q_rotate(q, v):                                                       # torch_util.py:65-71
    w, u = q[..., 0:1], q[..., 1:]
    t = 2 * torch.cross(u, v, dim=-1)
    return v + w * t + torch.cross(u, t, dim=-1)


def q_mul(a, b):                                                          # torch_util.py:48-62
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], dim=-1)


def q_conj(q):                                                            # torch_util.py:34-37
    return torch.cat([q[..., 0:1], -q[..., 1:]], dim=-1)


def q_pos(q):                                                             # torch_util.py:40-45
    return (1 - 2 * (q[..., 0:1] < 0).float()) * q


def v_unit(x, eps=1e-9):                                                  # torch_util.py:11-14
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def q_axis_angle(q):                                                      # torch_util.py:74-94
    q = q_pos(q)
    length = torch.norm(q[..., 1:], dim=-1, p=2)
    angle = 2.0 * torch.atan2(length, q[..., 0])
    axis = q[..., 1:] / length.unsqueeze(-1)
    ok = length > 1e-5
    dflt = torch.zeros_like(axis)
    dflt[..., -1] = 1
    return torch.where(ok.unsqueeze(-1), axis, dflt), torch.where(ok, angle, torch.zeros_like(angle))


def axis_angle_q(axis, angle):                                            # torch_util.py:176-182
    th = (angle / 2).unsqueeze(-1)
    return v_unit(torch.cat([th.cos(), v_unit(axis) * th.sin()], dim=-1))


def q_exp_map(q):                                                         # torch_util.py:203-208
    axis, angle = q_axis_angle(q)
    return angle.unsqueeze(-1) * axis


def q_tan_norm(q):                                                        # torch_util.py:230-242
    ex = torch.zeros_like(q[..., 1:]); ex[..., 0] = 1
    ez = torch.zeros_like(q[..., 1:]); ez[..., 2] = 1
    return torch.cat([q_rotate(q, ex), q_rotate(q, ez)], dim=-1)


def q_diff_angle(q0, q1):                                                 # torch_util.py:269-284
    return q_axis_angle(q_mul(q1, q_conj(q0)))[1]


def q_slerp(q0, q1, t):                                                   # torch_util.py:300-323
    c = torch.sum(q0 * q1, dim=-1)
    q1 = torch.where((c < 0).unsqueeze(-1), -q1, q1)
    c = torch.abs(c).unsqueeze(-1)
    half = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    t = t.unsqueeze(-1)
    out = torch.sin((1 - t) * half) / s * q0 + torch.sin(t * half) / s * q1
    out = torch.where(torch.abs(s) < 0.001, 0.5 * q0 + 0.5 * q1, out)
    return torch.where(torch.abs(c) >= 1, q0, out)


def q_heading_inv(q):                                                     # torch_util.py:326-356
    ex = torch.zeros_like(q[..., 1:]); ex[..., 0] = 1
    d = q_rotate(q, ex)
    heading = torch.atan2(d[..., 1], d[..., 0])
    ez = torch.zeros_like(q[..., 1:]); ez[..., 2] = 1
    return axis_angle_q(ez, -heading)


def q_twist_angle(q, axis):                                               # torch_util.py:385-406
    p = torch.sum(axis * q[..., 1:], dim=-1)
    tw = q.clone()
    tw[..., 1:] = p.unsqueeze(-1) * axis
    tw = v_unit(q_pos(tw))
    ax, ang = q_axis_angle(tw)
    ang = ang.clone()
    ang[torch.sum(axis * ax, dim=-1) < 0] *= -1
    return ang


# ==================================================================================================
# motion table                                           anim/motion_lib.py, anim/kin_char_model.py
# ==================================================================================================
def arange_times(n, dt):
    """Values of torch.arange(0, len, dt) (fp32, CPU): ATen fills 2x8-lane blocks as
    float(double(dt*block_start)) + lane*dt (double) and the n%16 tail as float(dt*i)
    (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp).  Checked against torch.arange in tests."""
    i = np.arange(n)
    base = ((i - i % 8).astype(np.float64) * dt).astype(np.float32)
    v = (base.astype(np.float64) + (i % 8).astype(np.float64) * dt).astype(np.float32)
    tail = n - n % 16
    v[tail:] = (i[tail:].astype(np.float64) * dt).astype(np.float32)
    return torch.from_numpy(v)


class OracleMotionLib:
    """Frames -> 100 Hz step table, and the truncated-index lookup (motion_lib.py:18-335)."""

    def __init__(self, clips, weights, dof_axis, col_of_dof, dt, fix_start_idx=False, jrot_override=None):
        """clips: list of (frames[F,36] float64 ndarray, fps, loop_mode); dof_axis [D,3]; col_of_dof [D].
        jrot_override: optional per-clip [F,D,4] joint rotations to use instead of cos/sin of the hinge angles
        (tests feed the device-computed ones so that the branchy resampling is compared on identical inputs)."""
        self.dt, self.dt_inv = dt, round(1 / dt)
        self.axis = torch.as_tensor(dof_axis, dtype=torch.float32)
        D = self.axis.shape[0]
        self.D = D
        col = torch.as_tensor(np.asarray(col_of_dof), dtype=torch.long)
        w = torch.tensor(weights, dtype=torch.float32)
        self.weights = w / w.sum()                                          # motion_lib.py:236-239
        self.lengths = torch.tensor([1.0 / fps * (f.shape[0] - 1) for f, fps, _ in clips], dtype=torch.float32)
        self.loop_modes = torch.tensor([lm for _, _, lm in clips], dtype=torch.int)
        nframes = torch.tensor([f.shape[0] for f, _, _ in clips], dtype=torch.long)
        self.num_frames = nframes
        rows, fidx = [], []
        self.clip_frames = []       # per clip (pos, rot, jrot, vel, ang, dofvel): the reference's _frame_* tensors
        for m, (frames, fps, loop) in enumerate(clips):
            fr = torch.tensor(frames, dtype=torch.float32)                  # motion_lib.py:108-110
            pos, rot = fr[:, 0:3], fr[:, [6, 3, 4, 5]]                      # xyzw -> wxyz, motion_lib.py:10-15
            dof = fr[:, 7:][:, col]                                         # file order -> BFS order, :102-111
            F = fr.shape[0]
            ax = self.axis.unsqueeze(0).expand(F, D, 3)
            jrot = q_pos(axis_angle_q(ax, dof))                             # kin_char_model.py:595-639, :113-114
            self.frame_joint_rot = getattr(self, "frame_joint_rot", []) + [jrot]
            if jrot_override is not None:
                jrot = torch.as_tensor(jrot_override[m], dtype=torch.float32).cpu()
            vel = torch.zeros_like(pos)                                     # motion_lib.py:203-205
            vel[:-1] = fps * (pos[1:] - pos[:-1]); vel[-1] = vel[-2]
            ang = torch.zeros_like(pos)                                     # motion_lib.py:207-212
            ang[:-1] = fps * q_exp_map(q_mul(rot[1:], q_conj(rot[:-1]))); ang[-1] = ang[-2]
            fdt = 1.0 / fps                                                 # kin_char_model.py:226-266
            drot = v_unit(q_pos(q_mul(q_conj(jrot[:-1]), jrot[1:])))
            dv = torch.sum(self.axis * (q_exp_map(drot) / fdt), dim=-1)
            dofvel = torch.cat([dv, dv[-1:]], dim=0)
            self.clip_frames.append((pos, rot, jrot, vel, ang, dofvel))
            # _precompute_motion_steps -> calc_motion_frame (motion_lib.py:285-320, 61-88)
            L = self.lengths[m]
            n = int(math.ceil(float(L) / dt))
            t = arange_times(n, dt)
            phase = t / L                                                   # calc_phase :361-372
            wraps = torch.floor(phase) if loop == WRAP else torch.zeros_like(phase)
            phase = torch.clip(phase - wraps, 0.0, 1.0)
            pf = phase * (F - 1)                                            # _calc_frame_blend :118-131
            i0 = pf.long()
            i1 = torch.clamp(i0 + 1, max=F - 1)
            blend = pf - i0
            b = blend.unsqueeze(-1)
            rp = (1.0 - b) * pos[i0] + b * pos[i1]
            if loop == WRAP:                                                # _calc_loop_offset :133-150
                delta = pos[-1] - pos[0]; delta[-1] = 0.0
                rp = rp + wraps.unsqueeze(-1) * delta
            rr = q_slerp(rot[i0], rot[i1], blend)
            jr = q_slerp(jrot[i0], jrot[i1], b)
            dp = q_twist_angle(jr, self.axis.unsqueeze(0).expand(n, D, 3))  # KinCharModel.rot_to_dof :210-224
            rows.append(torch.cat([rp, rr, dp, vel[i0], ang[i0], dofvel[i0]], dim=-1))
            fidx.append(torch.stack([i0, i1], dim=-1))
        self.table = torch.cat(rows, dim=0)          # [S, 7+D+6+D]: pos3 rot4 dof | vel3 ang3 dofvel
        self.frame_idx = torch.cat(fidx, dim=0)
        nsteps = torch.tensor([r.shape[0] for r in rows], dtype=torch.long)
        self.num_steps = nsteps
        true_start = torch.cumsum(nsteps, 0) - nsteps
        quirk_start = torch.cumsum(nframes, 0) - nframes                    # motion_lib.py:280-282 (Q2)
        self.start_idx = true_start if fix_start_idx else quirk_start
        self.s_total = self.table.shape[0]

    def to(self, device):
        """Move every tensor (bench.py's GPU-eager comparator runs this port on the B200 as the reference would run)."""
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
            elif isinstance(v, list) and v and torch.is_tensor(v[0]):
                setattr(self, k, [t.to(device) for t in v])
        return self

    def calc_motion_frame(self, ids, times):                                # motion_lib.py:61-88,118-150,361-372
        """Interpolation at arbitrary (clip, time) queries from the concatenated 30 fps frames ->
        (root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_pos, dof_vel)."""
        cat = [torch.cat([c[k] for c in self.clip_frames], dim=0) for k in range(6)]
        pos, rot, jrot, vel, ang, dofvel = cat
        start = torch.cumsum(self.num_frames, 0) - self.num_frames          # frame offsets (what _motion_start_idx is)
        L, nf = self.lengths[ids], self.num_frames[ids]
        wrap = self.loop_modes[ids] == WRAP
        phase = times / L                                                   # calc_phase
        phase = torch.where(wrap, phase - torch.floor(phase), phase)
        phase = torch.clip(phase, 0.0, 1.0)
        i0 = (phase * (nf - 1)).long()                                      # _calc_frame_blend
        i1 = torch.min(i0 + 1, nf - 1)
        blend = phase * (nf - 1) - i0
        i0, i1 = i0 + start[ids], i1 + start[ids]
        b = blend.unsqueeze(-1)
        rp = (1.0 - b) * pos[i0] + b * pos[i1]
        rr = q_slerp(rot[i0], rot[i1], blend)
        jr = q_slerp(jrot[i0], jrot[i1], b)
        D = self.D
        dp = q_twist_angle(jr, self.axis.unsqueeze(0).expand(ids.shape[0], D, 3))
        delta = torch.stack([c[0][-1] - c[0][0] for c in self.clip_frames], dim=0)    # _motion_root_pos_delta, z zeroed
        delta[:, -1] = 0.0
        off = torch.where(wrap.unsqueeze(-1), torch.floor(times / L).unsqueeze(-1) * delta[ids], torch.zeros_like(rp))
        return rp + off, rr, vel[i0], ang[i0], jr, dp, dofvel[i0]

    def rows(self, ids, times):                                             # motion_lib.py:322-326
        fr = (times * self.dt_inv).long()
        fr = torch.clip(fr, 0, self.s_total - 1)
        return fr + self.start_idx[ids]

    def step(self, ids, times):
        r = self.table[self.rows(ids, times)]
        D = self.D
        return r[:, 0:3], r[:, 3:7], r[:, 7 + D:10 + D], r[:, 10 + D:13 + D], r[:, 7:7 + D], r[:, 13 + D:]


# ==================================================================================================
# observations / reward / done                       learning/add/add_observation.py, add_reward.py, add_done.py
# ==================================================================================================
def disc_obs_rows(pos, rot, vel, ang, dof, dofvel, cfg):                   # add_observation.py:462-554
    """inputs [n, H, .] oldest -> newest; returns [n, H*step_dim]"""
    p = pos.clone()
    if not cfg["global_obs"]:
        p[..., 0:2] = 0.0
    parts = [p, q_tan_norm(rot), dof]
    if cfg["enable_vel_obs"]:
        if cfg["global_obs"]:
            parts += [vel, ang, dofvel]
        else:
            h = q_heading_inv(rot)
            parts += [q_rotate(h, vel), q_rotate(h, ang), dofvel]
    o = torch.cat(parts, dim=-1)
    return o.reshape(o.shape[0], -1)


def policy_obs(root_pos, root_rot, root_vel, root_ang, dof, dofvel, phase, tar_pos, tar_rot, tar_dof, cfg):
    """add_observation.py:422-459 (char), :557-575 (phase), :578-717 (target)."""
    glob = cfg["global_obs"]
    hinv = q_heading_inv(root_rot)
    rot_obs = q_tan_norm(root_rot if glob else q_mul(hinv, root_rot))
    parts = [rot_obs, dof]
    if cfg["enable_vel_obs"]:
        parts += [root_vel, root_ang, dofvel] if glob else [q_rotate(hinv, root_vel), q_rotate(hinv, root_ang), dofvel]
    if cfg["root_height_obs"]:
        parts = [root_pos[:, 2:3]] + parts
    if cfg["enable_phase_obs"]:
        ph = phase.unsqueeze(-1)
        k = cfg["num_phase_encoding"]
        if k > 0:
            sc = (2.0 * torch.pi * torch.pow(2.0, torch.arange(k, dtype=phase.dtype))).unsqueeze(0)
            ph = torch.cat((ph, torch.sin(phase.unsqueeze(-1) * sc), torch.cos(phase.unsqueeze(-1) * sc)), dim=-1)
        parts.append(ph)
    if cfg["enable_tar_obs"]:
        ref_pos = root_pos if glob else tar_pos[:, 0]
        dpos = tar_pos - ref_pos.unsqueeze(-2)
        trot = tar_rot
        if not glob:
            h = q_heading_inv(tar_rot[:, 0]).unsqueeze(-2).expand(-1, tar_pos.shape[1], -1)
            dpos = q_rotate(h, dpos)
            trot = q_mul(h, tar_rot)
        if cfg["root_height_obs"]:
            dpos = dpos.clone(); dpos[..., 2] = tar_pos[..., 2]
        else:
            dpos = dpos[..., :2]
        t = torch.cat([dpos, q_tan_norm(trot), tar_dof], dim=-1)
        parts.append(t.reshape(t.shape[0], -1))
    return torch.cat(parts, dim=-1)


def tracking_reward(sim, ref, w, cfg):                                       # add_reward.py:104-177
    root_pos, root_rot, root_vel, root_ang, dof, dofvel = sim
    t_pos, t_rot, t_vel, t_ang, t_dof, t_dofvel = ref
    pd = t_dof - dof
    pose_err = torch.sum(w * pd * pd, dim=-1)
    vd = t_dofvel - dofvel
    vel_err = torch.sum(w * vd * vd, dim=-1)
    dp = (t_pos - root_pos).clone()
    track_root = cfg["enable_tar_obs"] and cfg["global_obs"]
    if not track_root:
        dp[..., 0:2] = 0
    if not cfg["root_height_obs"]:
        dp[..., 2] = 0
    pos_err = torch.sum(dp * dp, dim=-1)
    if not track_root:
        h0, h1 = q_heading_inv(root_rot), q_heading_inv(t_rot)
        root_rot, root_vel, root_ang = q_mul(h0, root_rot), q_rotate(h0, root_vel), q_rotate(h0, root_ang)
        t_rot, t_vel, t_ang = q_mul(h1, t_rot), q_rotate(h1, t_vel), q_rotate(h1, t_ang)
    rot_err = q_diff_angle(root_rot, t_rot)
    rot_err = rot_err * rot_err
    dv = t_vel - root_vel
    da = t_ang - root_ang
    v_err, a_err = torch.sum(dv * dv, dim=-1), torch.sum(da * da, dim=-1)
    return (cfg["reward_pose_w"] * torch.exp(-cfg["reward_pose_scale"] * pose_err)
            + cfg["reward_vel_w"] * torch.exp(-cfg["reward_vel_scale"] * vel_err)
            + cfg["reward_root_pose_w"] * torch.exp(-cfg["reward_root_pose_scale"] * (pos_err + 0.1 * rot_err))
            + cfg["reward_root_vel_w"] * torch.exp(-cfg["reward_root_vel_scale"] * (v_err + 0.1 * a_err)))


def done_flags(time, ep_len, root_pos, dof, t_pos, t_dof, contact, motion_times, motion_len, len_term, cfg):
    """add_done.py:97-147 -- later assignments win."""
    done = torch.zeros(time.shape, dtype=torch.int32)
    done[time >= ep_len] = TIME
    done[(motion_times >= motion_len) & len_term] = SUCC
    if cfg["enable_early_termination"]:
        failed = contact.clone()
        if cfg.get("pose_termination", False):
            dist = cfg.get("pose_termination_dist", 1.0)
            pf = torch.mean((t_dof - dof) ** 2, dim=-1) > dist
            if cfg["enable_tar_obs"] and cfg["global_obs"]:
                pf = pf | (torch.sum((t_pos - root_pos) ** 2, dim=-1) > dist)
            failed = failed | pf
        done[failed & (time > 0.0)] = FAIL
    return done


def td_lambda_return(r, next_vals, done, discount, lam):                     # base_agent.py:624-647
    ret = torch.zeros_like(r)
    reset = (done != NULL).float()
    ret[-1] = r[-1] + discount * next_vals[-1]
    for i in reversed(range(r.shape[0] - 1)):
        cl = lam * (1.0 - reset[i])
        ret[i] = r[i] + discount * ((1.0 - cl) * next_vals[i] + cl * ret[i + 1])
    return ret


def adamw_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.AdamW single-tensor update (torch/optim/adam.py _single_tensor_adam), in place."""
    if wd != 0:
        p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# ==================================================================================================
# randomness
# ==================================================================================================
class TorchRandom:
    """The reference's own draws, in its call order, on the global CPU generator."""

    def action_noise(self, n, dim):                  # distribution_gaussian_diag.py:84-88
        return torch.normal(torch.zeros(n, dim), torch.ones(n, dim))

    def exp_mask(self, n, prob):                     # ppo_agent.py:80-88
        return torch.bernoulli(torch.full([n, 1], prob, dtype=torch.float)).squeeze(-1)

    def motions(self, weights, n):                   # motion_lib.py:35-39
        return torch.multinomial(weights, num_samples=n, replacement=True)

    def segments(self, probs):                       # sampler.py:78
        return torch.multinomial(probs, 1, True).squeeze(-1)

    def uniform(self, n):                            # sampler.py:84
        return torch.rand(n)

    def randperm(self, n):                           # experience_buffer.py:15,85
        return torch.randperm(n, dtype=torch.long)


# ==================================================================================================
# the agent: one rollout + ADD/PPO update                    learning/{base,ppo,amp}_agent.py, add/add_agent.py
# ==================================================================================================
def mlp_forward(x, layers, masks=None, rec=None):
    """layers: list of (W, b, relu).  Test hooks for the ReLU-boundary measurement (tests/test_gpu_parity.py):
    `rec` (a list) receives the boolean mask `pre-activation > 0` of every ReLU layer; `masks` (a list of boolean
    tensors, one per ReLU layer) replaces relu(z) by z * mask -- the same function wherever the mask agrees with
    z > 0, zero second derivative like ReLU, so autograd's double backward is unchanged."""
    k = 0
    for W, b, relu in layers:
        x = torch.nn.functional.linear(x, W, b)
        if relu:
            if rec is not None:
                rec.append((x > 0).detach())
            x = torch.relu(x) if masks is None else x * masks[k].to(x.dtype)
            k += 1
    return x


PARAM_ORDER = ["_actor_layers.0", "_actor_layers.2", "_actor_layers.4", "_action_dist._mean_net",
               "_critic_layers.0", "_critic_layers.2", "_critic_layers.4", "_critic_out",
               "_disc_layers.0", "_disc_layers.2", "_disc_logits"]


def init_params(obs_dim, act_dim, disc_dim, hidden=(1024, 1024, 512), disc_hidden=(1024, 512), out_scale=0.01):
    """Weights in the reference's construction order and init rules (ppo_model.py:36-60, add_model.py:32-46,
    nets/fc_*.py:11-14, distribution_gaussian_diag.py:18-23): consumes the global RNG like torch.nn.Linear."""
    p = {}

    def trunk(prefix, in_dim, sizes):
        for i, h in enumerate(sizes):
            lin = torch.nn.Linear(in_dim, h)
            p["%s.%d.weight" % (prefix, 2 * i)] = lin.weight.data
            p["%s.%d.bias" % (prefix, 2 * i)] = torch.zeros(h)
            in_dim = h
        return in_dim

    a = trunk("_actor_layers", obs_dim, hidden)
    lin = torch.nn.Linear(a, act_dim)
    torch.nn.init.uniform_(lin.weight, -out_scale, out_scale)
    p["_action_dist._mean_net.weight"], p["_action_dist._mean_net.bias"] = lin.weight.data, torch.zeros(act_dim)
    c = trunk("_critic_layers", obs_dim, hidden)
    lin = torch.nn.Linear(c, 1)
    p["_critic_out.weight"], p["_critic_out.bias"] = lin.weight.data, torch.zeros(1)
    d = trunk("_disc_layers", disc_dim, disc_hidden)
    lin = torch.nn.Linear(d, 1)
    torch.nn.init.uniform_(lin.weight, -1.0, 1.0)
    p["_disc_logits.weight"], p["_disc_logits.bias"] = lin.weight.data, torch.zeros(1)
    return p


class OracleAgent:
    """State + one training iteration of the reference agent, driven through an engine that offers the
    BaseEntity getters/setters (add_gym_b200.engine.SyntheticEngine on CPU)."""

    def __init__(self, cfg, env, lib, rng=None, mimic_reference_rng=True):
        self.cfg, self.task, self.acfg = cfg, cfg["task"], cfg["agent"]
        self.env, self.lib = env, lib
        self.rng = rng or TorchRandom()
        t, a = self.task, self.acfg
        self.t = {"global_obs": t.get("global_obs", False), "root_height_obs": t.get("root_height_obs", False),
                  "enable_vel_obs": t.get("enable_vel_obs", False), "enable_phase_obs": t.get("enable_phase_obs", True),
                  "enable_tar_obs": t.get("enable_tar_obs", False), "num_phase_encoding": t.get("num_phase_encoding", 0),
                  "enable_early_termination": t["enable_early_termination"],
                  "pose_termination": t.get("pose_termination", False),
                  "pose_termination_dist": t.get("pose_termination_dist", 1.0)}
        for k in ("reward_pose_w", "reward_vel_w", "reward_root_pose_w", "reward_root_vel_w", "reward_pose_scale",
                  "reward_vel_scale", "reward_root_pose_scale", "reward_root_vel_scale"):
            self.t[k] = t.get(k)
        N, D = env.num_envs, lib.D
        self.N, self.D = N, D
        self.dt = env.ctrl_dt
        self.H = t["num_disc_obs_steps"]
        self.tar_steps = torch.tensor(t.get("tar_obs_steps", [1]), dtype=torch.int)
        self.ep_len = t.get("max_episode_length", torch.sum(lib.lengths).item())
        self.rand_reset = t.get("rand_reset", True)
        # sampler (sampler.py:5-19)
        sc = t.get("sampler", {})
        self.num_segments = sc.get("num_segments", 20)
        self.temperature = sc.get("temperature", None)
        self.seg_sizes = torch.tensor([l / self.num_segments for l in lib.lengths])
        self.errors = torch.ones((lib.lengths.shape[0], self.num_segments))
        self.min_start = (self.H - 1) * self.dt
        # per-env state (add_observation.py:43-63)
        self.motion_ids = torch.zeros(N, dtype=torch.int64)
        self.offsets = torch.zeros(N, dtype=torch.float32)
        self.time_buf = torch.zeros(N, dtype=torch.float32)
        self.ref = [torch.zeros(N, k) for k in (3, 4, 3, 3, D, D)]
        self.hist = [torch.zeros(N, self.H, k) for k in (3, 4, 3, 3, D, D)]     # circular_buffer.py
        self.head = 0
        self.done_buf = torch.zeros(N, dtype=torch.int32)
        self.dof_err_w = torch.ones(D)
        ent = env.robot.entity
        contact = [ent.get_link(name=n).idx for n in t.get("contact_bodies", [])]
        self.noncontact = torch.tensor([l.idx for l in ent.links if l.idx not in contact], dtype=torch.long)
        # shapes; the reference infers them by sampling demo obs (5 RNG-consuming calls)
        if mimic_reference_rng:
            self.sample_time(N)                                   # ADDObservation._build_disc_obs_buffers
        self.obs_dim = self.compute_obs().shape[1]
        self.disc_dim = self.H * (9 + D + ((6 + D) if self.t["enable_vel_obs"] else 0))
        if mimic_reference_rng:
            self.sample_time(N)                                   # ADDAgent._build_normalizers
            self.sample_time(N)                                   # _build_model -> get_disc_obs_shape
        aspace = env.robot.get_action_space().to(lib.table.device)
        self.act_dim = aspace.shape[0]
        self.a_mean = 0.5 * (aspace[:, 1] + aspace[:, 0])
        self.a_std = 0.5 * (aspace[:, 1] - aspace[:, 0])
        self.params = init_params(self.obs_dim, self.act_dim, self.disc_dim, out_scale=a["model"]["actor_init_output_scale"])
        self.logstd = torch.full((self.act_dim,), float(np.log(a["model"]["action_std"])), dtype=torch.float32)
        for v in self.params.values():
            v.requires_grad_(True)
        self.names = [n + s for n in PARAM_ORDER for s in (".weight", ".bias")]
        self.adam_m = {k: torch.zeros_like(self.params[k]) for k in self.names}
        self.adam_v = {k: torch.zeros_like(self.params[k]) for k in self.names}
        self.adam_steps = 0
        # normalizers (normalizer.py, diff_normalizer.py)
        self.obs_count, self.obs_mean, self.obs_std = 0, torch.zeros(self.obs_dim), torch.ones(self.obs_dim)
        self.obs_mean_sq = None
        self.obs_new = [0, torch.zeros(self.obs_dim), torch.zeros(self.obs_dim)]
        self.diff_count, self.diff_mean_abs = 0, torch.ones(self.disc_dim)
        self.diff_new = [0, torch.zeros(self.disc_dim)]
        # experience buffer (experience_buffer.py:4-21; base/ppo/amp/add _build_exp_buffer)
        self.T = a["steps_per_iter"]
        T = self.T
        self.perm = self.rng.randperm(T * N)
        self.perm[:] = self.rng.randperm(T * N)
        self.perm_head = 0
        if mimic_reference_rng:
            self.sample_time(N)                                   # AMPAgent._build_exp_buffer -> get_disc_obs_shape
            self.sample_time(N)                                   # ADDAgent._build_pos_diff -> get_disc_obs_space
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
        self.buf = {"obs": z(T, N, self.obs_dim), "next_obs": z(T, N, self.obs_dim), "action": z(T, N, self.act_dim),
                    "reward": z(T, N), "done": z(T, N, dt=torch.int32), "a_logp": z(T, N), "tar_val": z(T, N),
                    "adv": z(T, N), "rand_action_mask": z(T, N), "disc_obs": z(T, N, self.disc_dim),
                    "disc_obs_demo": z(T, N, self.disc_dim), "motion_ids": z(T, N, dt=torch.long),
                    "motion_times": z(T, N)}
        self.buf_head, self.total_samples = 0, 0
        self.ret_buf, self.len_buf = torch.zeros(N), torch.zeros(N, dtype=torch.long)
        self.ep_sum, self.len_sum, self.episodes = 0.0, 0.0, 0
        self.sample_count = 0
        self.trace = {"reset_ids": [], "reset_times": [], "reset_envs": []}
        self.record_masks, self.last_masks = False, None          # test hook: ReLU masks of the last optimizer step

    # ---- nets ------------------------------------------------------------------------------------------------
    def _trunk(self, prefix, n):
        return [(self.params["%s.%d.weight" % (prefix, 2 * i)], self.params["%s.%d.bias" % (prefix, 2 * i)], True)
                for i in range(n)]

    def actor_mean(self, x, masks=None, rec=None):
        h = mlp_forward(x, self._trunk("_actor_layers", 3), masks, rec)
        return torch.nn.functional.linear(h, self.params["_action_dist._mean_net.weight"],
                                          self.params["_action_dist._mean_net.bias"])

    def critic(self, x, masks=None, rec=None):
        h = mlp_forward(x, self._trunk("_critic_layers", 3), masks, rec)
        return torch.nn.functional.linear(h, self.params["_critic_out.weight"], self.params["_critic_out.bias"])

    def disc(self, x, masks=None, rec=None):
        h = mlp_forward(x, self._trunk("_disc_layers", 2), masks, rec)
        return torch.nn.functional.linear(h, self.params["_disc_logits.weight"], self.params["_disc_logits.bias"])

    def logp(self, x, mean):                                                # distribution_gaussian_diag.py:90-94
        std = torch.exp(torch.broadcast_to(self.logstd, mean.shape))
        lp = -0.5 * torch.sum(torch.square((x - mean) / std), dim=-1)
        lp += -0.5 * self.act_dim * np.log(2.0 * np.pi) - torch.sum(torch.broadcast_to(self.logstd, mean.shape), dim=-1)
        return lp

    # ---- sampling --------------------------------------------------------------------------------------------
    def sample_time(self, n):                                               # add_motion.py:53-61, sampler.py:57-92
        ids = self.rng.motions(self.lib.weights, n)
        if not self.rand_reset:
            return ids, torch.zeros(n, dtype=torch.float)
        ce = self.errors[ids]
        temp = torch.max(ce) + 1e-6 if self.temperature is None else self.temperature
        probs = torch.nn.functional.softmax(ce / temp, dim=-1)
        seg = self.rng.segments(probs)
        sz = self.seg_sizes[ids]
        time = seg * sz + self.rng.uniform(n) * sz
        time = (time // self.dt) * self.dt
        return ids, torch.clamp(time, min=self.min_start)

    # ---- observation plumbing ----------------------------------------------------------------------------------
    def sim(self):
        r = self.env.robot
        return [r.base_pos, r.base_quat, r.base_lin_vel, r.base_ang_vel, r.dof_pos, r.dof_vel]

    def motion_times(self):
        return self.time_buf + self.offsets

    def update_ref(self):                                                   # add_observation.py:163-175
        rows = self.lib.step(self.motion_ids, self.motion_times())
        for dst, src in zip(self.ref, rows):
            dst[:] = src

    def hist_all(self, k):                                                  # circular_buffer.py:46-57
        b = self.hist[k]
        return b if self.head == 0 else torch.cat([b[:, self.head:], b[:, :self.head]], dim=1)

    def demo_rows(self, ids, t0):                                           # add_observation.py:356-400
        n = ids.shape[0]
        steps = torch.flip(-self.dt * torch.arange(0, self.H), dims=[0])
        times = (t0.unsqueeze(-1) + steps).view(-1)
        rows = self.lib.step(torch.tile(ids.unsqueeze(-1), [1, self.H]).view(-1), times)
        return [r.reshape(n, self.H, r.shape[-1]) for r in rows]

    def compute_obs(self):                                                  # add_observation.py:209-274
        mt = self.motion_times()
        pos, rot, vel, ang, dof, dofvel = self.sim()
        phase = None
        if self.t["enable_phase_obs"]:
            L = self.lib.lengths[self.motion_ids]
            phase = mt / L
            wrap = self.lib.loop_modes[self.motion_ids] == WRAP
            phase = torch.clip(torch.where(wrap, phase - torch.floor(phase), phase), 0.0, 1.0)
        tp = tr = td = None
        if self.t["enable_tar_obs"]:
            K = self.tar_steps.shape[0]
            times = (mt.unsqueeze(-1) + self.dt * self.tar_steps).flatten()
            ids = torch.broadcast_to(self.motion_ids.unsqueeze(-1), (self.N, K)).flatten()
            r = self.lib.step(ids, times)
            tp, tr, td = r[0].reshape(self.N, K, 3), r[1].reshape(self.N, K, 4), r[4].reshape(self.N, K, self.D)
        return policy_obs(pos, rot, vel, ang, dof, dofvel, phase, tp, tr, td, self.t)

    def compute_all_obs(self):                                              # add_observation.py:301-306
        self.obs = self.compute_obs()
        self.disc_obs = disc_obs_rows(*[self.hist_all(k) for k in range(6)], self.t)
        pr, rr, vr, ar, dr, dvr = self.demo_rows(self.motion_ids, self.motion_times())
        self.disc_obs_demo = disc_obs_rows(pr, rr, vr, ar, dr, dvr, self.t)

    def contact_bool(self):                                                 # robot.py:221-231
        c = self.env.robot.entity.get_contacts(with_entity=self.env.plane, exclude_self_contact=True)
        a = torch.isin(c["link_a"], self.noncontact) & c["valid_mask"]
        b = torch.isin(c["link_b"], self.noncontact) & c["valid_mask"]
        return a.any(dim=1) | b.any(dim=1)

    # ---- env step / reset ---------------------------------------------------------------------------------------
    def step_env(self, action):                                             # add_agent.py:204-219, env.py:150-155
        self.env.robot.apply_action(action)
        self.env.scene.step()
        self.time_buf += self.dt
        self.update_ref()
        s = self.sim()
        for k in range(6):                                                  # add_observation.py:192-207
            self.hist[k][:, self.head] = s[k]
        self.head = (self.head + 1) % self.H
        self.compute_all_obs()
        r = tracking_reward(s, self.ref, self.dof_err_w, {**self.t})
        mt = self.motion_times()
        L = self.lib.lengths[self.motion_ids]
        len_term = self.lib.loop_modes[self.motion_ids] != WRAP
        self.done_buf[:] = done_flags(self.time_buf, self.ep_len, s[0], s[4], self.ref[0], self.ref[4],
                                      self.contact_bool(), mt, L, len_term, self.t)
        return self.obs, r, self.done_buf

    def reset_envs(self, env_ids=None):                                     # add_agent.py:221-233
        if env_ids is None:
            env_ids = torch.arange(self.N)
        if len(env_ids) > 0:
            self.time_buf[env_ids] = 0
            self.done_buf[env_ids] = NULL
            ids, times = self.sample_time(len(env_ids))                     # add_observation.py:308-332
            self.trace["reset_envs"].append(env_ids.clone()); self.trace["reset_ids"].append(ids.clone())
            self.trace["reset_times"].append(times.clone())
            self.motion_ids[env_ids] = ids
            self.offsets[env_ids] = times
            self.update_ref()
            ent = self.env.robot.entity
            ent.set_qpos(torch.cat([self.ref[0][env_ids], self.ref[1][env_ids], self.ref[4][env_ids]], dim=-1), envs_idx=env_ids)
            ent.set_dofs_velocity(torch.cat([self.ref[2][env_ids], self.ref[3][env_ids], self.ref[5][env_ids]], dim=-1),
                                  envs_idx=env_ids)
            rows = self.demo_rows(self.motion_ids[env_ids], self.time_buf[env_ids] + self.offsets[env_ids])
            for k in range(6):                                              # circular_buffer.py:22-29
                b, d, h = self.hist[k], rows[k], self.head
                b[env_ids, :h] = d[:, self.H - h:]
                b[env_ids, h:] = d[:, :self.H - h]
            self.compute_all_obs()
        else:
            self.trace["reset_envs"].append(env_ids.clone()); self.trace["reset_ids"].append(torch.zeros(0, dtype=torch.long))
            self.trace["reset_times"].append(torch.zeros(0))
        return self.obs

    # ---- rollout --------------------------------------------------------------------------------------------------
    def norm_obs(self, x):
        return (x - self.obs_mean) / self.obs_std

    def need_norm_update(self):
        return self.sample_count < self.acfg.get("normalizer_samples", np.inf)

    @torch.no_grad()
    def rollout(self, steps=None):                                          # base_agent.py:379-391
        b = self.buf
        for _ in range(steps or self.T):
            h = self.buf_head
            obs = self.curr_obs
            mean = self.actor_mean(self.norm_obs(obs))                      # ppo_agent.py:72-104
            std = torch.exp(torch.broadcast_to(self.logstd, mean.shape))
            a_rand = mean + std * self.rng.action_noise(self.N, self.act_dim)
            mask = self.rng.exp_mask(self.N, self.acfg.get("exp_prob_beg", 1.0))
            norm_a = torch.where(mask.unsqueeze(-1) == 1.0, a_rand, mean)
            logp = self.logp(norm_a, mean)
            action = norm_a * self.a_std + self.a_mean
            b["obs"][h], b["action"][h], b["a_logp"][h], b["rand_action_mask"][h] = obs, action, logp, mask
            if self.need_norm_update():                                     # normalizer.py:25-35
                self.obs_new[0] += self.N
                self.obs_new[1] += torch.sum(obs, axis=0)
                self.obs_new[2] += torch.sum(torch.square(obs), axis=0)
            next_obs, r, done = self.step_env(action)
            self.ret_buf += r                                               # base_agent.py:596-621
            self.len_buf += 1
            ids = (done != NULL).nonzero(as_tuple=False).flatten()
            if len(ids) > 0:
                self.ep_sum += float(torch.sum(self.ret_buf[ids])); self.len_sum += float(torch.sum(self.len_buf[ids]))
                self.episodes += len(ids)
                self.ret_buf[ids] = 0.0; self.len_buf[ids] = 0
            b["next_obs"][h], b["reward"][h], b["done"][h] = next_obs, r, done
            b["disc_obs_demo"][h], b["disc_obs"][h] = self.disc_obs_demo, self.disc_obs
            b["motion_ids"][h], b["motion_times"][h] = self.motion_ids, self.offsets + self.time_buf
            if self.need_norm_update():                                     # diff_normalizer.py:24-31
                self.diff_new[0] += self.N
                self.diff_new[1] += torch.sum(torch.abs(self.disc_obs_demo - self.disc_obs), axis=0)
            self.curr_obs = self.reset_envs(ids).clone()
            self.buf_head = (h + 1) % self.T
            self.total_samples += self.N

    # ---- train data ---------------------------------------------------------------------------------------------
    def norm_diff(self, x):                                                 # diff_normalizer.py:56-60
        return x / torch.clamp_min(self.diff_mean_abs, 1e-4)

    @torch.no_grad()
    def build_train_data(self):                                             # add_agent.py:110-139, ppo_agent.py:111-159
        a, b = self.acfg, self.buf
        flat = lambda k: b[k].view([self.T * self.N] + list(b[k].shape[2:]))
        dobs, demo = flat("disc_obs"), flat("disc_obs_demo")
        logits = self.disc(self.norm_diff(demo - dobs)).squeeze(-1)         # amp_agent.py:194-206
        prob = 1 / (1 + torch.exp(-logits))
        disc_r = -torch.log(torch.maximum(1 - prob, torch.tensor(0.0001)))
        disc_r *= a["disc_reward_scale"]
        self.update_errors(flat("motion_ids"), flat("motion_times"), torch.sum(torch.square(dobs - demo), dim=-1))
        flat("reward")[:] = a["task_reward_weight"] * flat("reward") + a["disc_reward_weight"] * disc_r
        next_vals = self.critic(self.norm_obs(b["next_obs"])).squeeze(-1)
        next_vals[b["done"] == SUCC] = 0.0 / (1.0 - a["discount"])
        next_vals[b["done"] == FAIL] = 0.0 / (1.0 - a["discount"])
        new_vals = td_lambda_return(b["reward"], next_vals, b["done"], a["discount"], a["td_lambda"])
        vals = self.critic(self.norm_obs(b["obs"])).squeeze(-1)
        adv = new_vals - vals
        sel = adv.flatten()[(b["rand_action_mask"] == 1.0).flatten()]
        adv_std, adv_mean = torch.std_mean(sel, dim=0)
        norm_adv = torch.clamp((adv - adv_mean) / torch.clamp_min(adv_std, 1e-5), -a["norm_adv_clip"], a["norm_adv_clip"])
        b["tar_val"][:] = new_vals
        b["adv"][:] = norm_adv
        dstd, dmean = torch.std_mean(disc_r)
        return {"adv_mean": adv_mean, "adv_std": adv_std, "disc_reward_mean": dmean, "disc_reward_std": dstd,
                "next_vals": next_vals, "vals": vals, "disc_r": disc_r, "logits": logits}

    def update_errors(self, ids, times, err):                               # sampler.py:20-55
        sz = torch.clamp(self.seg_sizes[ids], min=1e-6)
        seg = torch.clamp((times / sz).long(), 0, self.num_segments - 1)
        flat = ids * self.num_segments + seg
        n = self.errors.numel()
        mean = torch.zeros(n).scatter_reduce(0, flat, err, reduce="mean", include_self=False).view_as(self.errors)
        hit = torch.zeros(n).scatter_add(0, flat, torch.ones_like(err)).view_as(self.errors) > 0
        self.errors = torch.where(hit, 0.9 * self.errors + 0.1 * mean, self.errors)

    # ---- update -------------------------------------------------------------------------------------------------
    def sample_idx(self, n):                                                # experience_buffer.py:90-113
        L = self.perm.shape[0]
        if self.perm_head + n <= L:
            idx = self.perm[self.perm_head:self.perm_head + n]
            self.perm_head += n
        else:
            idx0 = self.perm[self.perm_head:]          # view: re-drawn in place below, like the reference
            rem = n - (L - self.perm_head)
            self.perm[:] = self.rng.randperm(L)
            idx = torch.cat([idx0, self.perm[:rem]], dim=0)
            self.perm_head = rem
        return torch.remainder(idx, min(self.total_samples, L))

    def loss(self, idx, masks=None, rec=None):
        """AMPAgent._compute_loss with ADD's discriminator loss (amp_agent.py:98-114; ppo_agent.py:194-261;
        base_agent.py:522-546; add_agent.py:141-202).

        Test hooks (see mlp_forward): `rec` (a dict) receives the ReLU masks of the five forward passes under the keys
        "actor", "critic", "disc" (minibatch rows) and "disc_pos" (the zero-difference row); `masks` (same keys) forces
        them -- used to measure how much of a gradient difference is ReLU boundary flips."""
        a, b = self.acfg, self.buf
        mk = lambda k: None if masks is None else masks[k]
        rc = lambda k: None if rec is None else rec.setdefault(k, [])
        flat = lambda k: b[k].view([self.T * self.N] + list(b[k].shape[2:]))[idx]
        norm_obs = self.norm_obs(flat("obs"))
        norm_a = (flat("action") - self.a_mean) / self.a_std
        pred = self.critic(norm_obs, mk("critic"), rc("critic")).squeeze(-1)
        critic_loss = torch.mean(torch.square(flat("tar_val") - pred))
        m = flat("rand_action_mask") == 1.0
        mean = self.actor_mean(norm_obs[m], mk("actor"), rc("actor"))
        ratio = torch.exp(self.logp(norm_a[m], mean) - flat("a_logp")[m])
        adv = flat("adv")[m]
        clip = a["ppo_clip_ratio"]
        actor_loss = -torch.mean(torch.minimum(adv * ratio, adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)))
        info = {"clip_frac": torch.mean((torch.abs(ratio - 1.0) > clip).float()).detach(),
                "imp_ratio": torch.mean(ratio).detach()}
        if a["action_bound_weight"] != 0:
            viol = torch.sum(torch.square(torch.clamp_max(mean + 1, 0.0)), dim=-1) + \
                torch.sum(torch.square(torch.clamp_min(mean - 1, 0)), dim=-1)
            bound = torch.mean(viol)
            actor_loss = actor_loss + a["action_bound_weight"] * bound
            info["action_bound_loss"] = bound.detach()
        loss = actor_loss + a["critic_loss_weight"] * critic_loss
        # discriminator
        pos_logit = self.disc(torch.zeros(1, self.disc_dim), mk("disc_pos"), rc("disc_pos")).squeeze(-1)
        x = self.norm_diff(flat("disc_obs_demo") - flat("disc_obs")).requires_grad_(True)
        neg_logit = self.disc(x, mk("disc"), rc("disc")).squeeze(-1)
        bce = torch.nn.BCEWithLogitsLoss()
        disc_loss = 0.5 * (bce(pos_logit, torch.ones_like(pos_logit) * .9) + bce(neg_logit, torch.ones_like(neg_logit) * .1))
        wl = torch.flatten(self.params["_disc_logits.weight"])
        logit_loss = torch.sum(torch.square(wl))
        disc_loss = disc_loss + a["disc_logit_reg"] * logit_loss
        g = torch.autograd.grad(neg_logit, x, grad_outputs=torch.ones_like(neg_logit), create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
        pen = torch.mean(torch.square(torch.sqrt(torch.sum(torch.square(g), dim=-1) + 1e-8) - 1))
        disc_loss = disc_loss + a["disc_grad_penalty"] * pen
        if a["disc_weight_decay"] != 0:
            w = torch.cat([torch.flatten(self.params["_disc_layers.0.weight"]),
                           torch.flatten(self.params["_disc_layers.2.weight"]), wl], dim=-1)
            disc_loss = disc_loss + a["disc_weight_decay"] * torch.sum(torch.square(w))
        loss = loss + a["disc_loss_weight"] * disc_loss
        info.update({"loss": loss, "critic_loss": critic_loss.detach(), "actor_loss": actor_loss.detach(),
                     "disc_loss": disc_loss.detach(), "disc_grad_penalty": pen.detach(),
                     "disc_logit_loss": logit_loss.detach(), "disc_pos_acc": torch.mean((pos_logit > 0).float()).detach(),
                     "disc_neg_acc": torch.mean((neg_logit < 0).float()).detach(),
                     "disc_pos_logit": torch.mean(pos_logit).detach(), "disc_neg_logit": torch.mean(neg_logit).detach()})
        return info

    def optimizer_step(self, loss, grad_hook=None):                         # mp_optimizer.py:14-23
        for p in self.params.values():
            p.grad = None
        loss.backward()
        clip = float(self.acfg["optimizer"].get("grad_clip", 0.0))             # mp_optimizer.py:10,19-20,46-47 (0 by default: Q4)
        if clip > 0.0:
            torch.nn.utils.clip_grad_norm_([self.params[k] for k in self.names], clip)
        if grad_hook is not None:
            grad_hook({k: self.params[k].grad for k in self.names})
        self.adam_steps += 1
        lr = float(self.acfg["optimizer"]["learning_rate"])
        with torch.no_grad():
            for k in self.names:
                adamw_step(self.params[k], self.params[k].grad, self.adam_m[k], self.adam_v[k], self.adam_steps, lr)

    def update_model(self, on_step=None, max_steps=None, grad_hook=None):   # ppo_agent.py:171-192
        a = self.acfg
        M = a["batch_size"] * self.N
        nb = int(np.ceil(float(min(self.total_samples, self.T * self.N)) / M))
        acc, steps = {}, 0
        for _ in range(a["update_epochs"]):
            for _ in range(nb):
                if max_steps is not None and steps >= max_steps:
                    break
                idx = self.sample_idx(M)
                rec = {} if self.record_masks else None
                info = self.loss(idx, rec=rec)
                self.last_masks = rec
                self.optimizer_step(info["loss"], grad_hook)
                if on_step is not None:
                    on_step(steps, idx, info, self)
                for k, v in info.items():
                    acc[k] = acc.get(k, 0.0) + v.detach()
                steps += 1
        return {k: v / steps for k, v in acc.items()}

    def update_normalizers(self):                                           # normalizer.py:37-80, diff_normalizer.py:33-45
        if self.obs_mean_sq is None:
            self.obs_mean_sq = torch.square(self.obs_std) + torch.square(self.obs_mean)
        n = self.obs_new[0]
        if n > 0:
            nm, nms = self.obs_new[1] / n, self.obs_new[2] / n
            tot = torch.tensor([self.obs_count + n])
            w_old = torch.tensor([self.obs_count]).float() / tot.float()
            w_new = float(n) / tot.float()
            self.obs_mean = w_old * self.obs_mean + w_new * nm
            self.obs_mean_sq = w_old * self.obs_mean_sq + w_new * nms
            self.obs_count += n
            self.obs_std = torch.sqrt(torch.clamp_min(self.obs_mean_sq - torch.square(self.obs_mean), 1e-4 * 1e-4))
            self.obs_new = [0, torch.zeros(self.obs_dim), torch.zeros(self.obs_dim)]
        n = self.diff_new[0]
        tot = torch.tensor([self.diff_count + n])
        w_old = torch.tensor([self.diff_count]).float() / tot.float()
        w_new = float(n) / tot.float()
        self.diff_mean_abs = w_old * self.diff_mean_abs + w_new * (self.diff_new[1] / n)
        self.diff_count += n
        self.diff_new = [0, torch.zeros(self.disc_dim)]

    def train_iter(self):                                                   # base_agent.py:353-374
        self.rollout()
        data = self.build_train_data()
        info = self.update_model()
        if self.need_norm_update():
            self.update_normalizers()
        return {**info, **{k: data[k] for k in ("adv_mean", "adv_std", "disc_reward_mean", "disc_reward_std")}}

    def start(self):
        """train_model prologue (base_agent.py:79-87): reset all envs, then _init_train -> exp_buffer.clear(),
        which re-draws the minibatch permutation (experience_buffer.py:33-39,83-88)."""
        self.curr_obs = self.reset_envs().clone()
        self.perm[:] = self.rng.randperm(self.perm.shape[0])
        self.perm_head, self.buf_head, self.total_samples = 0, 0, 0
