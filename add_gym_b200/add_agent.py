"""ADDAgent: rollout + ADD/PPO update on the B200 path, drop-in for the reference agent stack
``ADDAgent(AMPAgent(PPOAgent(BaseAgent)))`` (add/add_agent.py:20-266, amp_agent.py:7-206,
ppo_agent.py:9-279, base_agent.py:28-647).

Kept from the reference: constructor ``ADDAgent(env_config, distributed=False)``; the hook names
(``_decide_action``, ``_step_env``, ``_reset_envs``, ``_reset_done_envs``, ``_rollout_train``,
``_build_train_data``, ``_update_model``, ``_update_normalizers``, ``_train_iter``, ``train_model``,
``test_model``, ``save`` / ``load``); the 13 experience-buffer keys; the state-dict key names and the
checkpoint dict layout (base_agent.py:148-155) so checkpoints move both ways.

What is different underneath: one env step = 5 kernels from the actor forward + 1 fused step kernel +
the masked reset (no ``nonzero``/``len`` host sync when the engine offers ``set_state_masked``); one
optimizer step = one native call that issues the whole gather/forward/backward/AdamW launch sequence;
diagnostics stay on the device until the end of the iteration.  Gradients ARE all-reduced across ranks
(the reference's DDP wrapper never fires): one flat 17.4 MB NCCL all-reduce per step.
"""
import ctypes as C
import enum
import gc
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .add_model import ADDModel
from .add_motion import ADDMotion
from .add_observation import ADDDone, ADDObservation, ADDReward, DoneFlags
from .env import ImitationEnvironment
from .experience_buffer import ExperienceBuffer
from .normalizer import DiffNormalizer, Normalizer


class AgentMode(enum.Enum):
    TRAIN = 0
    TEST = 1


class DeviceRandom:
    """Default randomness: torch generators on the device.  Parity tests swap in a replaying source."""

    def __init__(self, device):
        self.device = device

    def action_noise(self, n, dim):
        return torch.randn(n, dim, device=self.device)

    def exp_mask(self, n, prob):
        if prob >= 1.0:
            return None
        return torch.bernoulli(torch.full([n], prob, device=self.device, dtype=torch.float))

    def reset_uniforms(self, n):
        return torch.rand(n, 3, device=self.device)

    def randperm(self, n):
        return torch.randperm(n, device=self.device, dtype=torch.long)


class MPOptimizerState:
    """state_dict-compatible holder of the flat AdamW moments (mp_optimizer.py:47-53)."""

    def __init__(self, config, model):
        self.lr = float(config["learning_rate"])
        self.weight_decay = float(config.get("weight_decay", 0.0))
        assert config["type"] == "Adam", "the B200 path implements the reference's Adam(W) optimizer"
        self._grad_clip = float(config.get("grad_clip", 0.0))   # read from config["optimizer"] like the reference (Q4)
        self.betas, self.eps = (0.9, 0.999), 1e-8
        self.exp_avg = torch.zeros_like(model.flat)
        self.exp_avg_sq = torch.zeros_like(model.flat)
        self.steps = 0
        self._model = model
        self.on_hyperparams_changed = None
        self.shard = None          # (begin, end): with the peer-memory exchange only this slice of the moments is current

    def get_steps(self):
        return self.steps

    def _gather_moments(self):
        """Peer-memory exchange: every rank keeps the moments of its own shard current -- assemble the full vectors."""
        if self.shard is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        n = self.exp_avg.numel()
        per = 4 * (((n + 3) // 4 + world - 1) // world)
        for vec in (self.exp_avg, self.exp_avg_sq):
            pad = torch.zeros(per * world, dtype=vec.dtype, device=vec.device)
            mine = torch.zeros(per, dtype=vec.dtype, device=vec.device)
            b, e = self.shard
            mine[:e - b] = vec[b:e]
            dist.all_gather_into_tensor(pad, mine)
            vec.copy_(pad[:n])

    def state_dict(self):
        self._gather_moments()
        names, tensors = self._model.trainable()
        state = {}
        for i, (n, t) in enumerate(zip(names, tensors)):
            o = self._model.offsets["o_" + n]
            state[i] = {"step": torch.tensor(float(self.steps)),
                        "exp_avg": self.exp_avg[o:o + t.numel()].view(t.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + t.numel()].view(t.shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "params": list(range(len(tensors)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        names, tensors = self._model.trainable()
        for i, (n, t) in enumerate(zip(names, tensors)):
            if i not in sd["state"]:
                continue
            o = self._model.offsets["o_" + n]
            s = sd["state"][i]
            self.exp_avg[o:o + t.numel()].view(t.shape).copy_(s["exp_avg"])
            self.exp_avg_sq[o:o + t.numel()].view(t.shape).copy_(s["exp_avg_sq"])
            self.steps = int(float(s["step"]))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g["weight_decay"])
        if self.on_hyperparams_changed is not None:      # the update context bakes these scalars in: rebuild it
            self.on_hyperparams_changed()


INFO_KEYS = ["loss", "critic_loss", "actor_loss", "clip_frac", "imp_ratio", "action_bound_loss", "disc_loss",
             "disc_grad_penalty", "disc_logit_loss", "disc_pos_acc", "disc_neg_acc", "disc_pos_logit", "disc_neg_logit"]


class ADDAgent(torch.nn.Module):
    NAME = "ADD"

    def __init__(self, env_config, distributed=False, device=None):
        super().__init__()
        if type(self)._compute_loss is not ADDAgent._compute_loss:
            raise _lib.AddkError("%s overrides _compute_loss, which is fused into addk_update_minibatch on the B200 path "
                                 "and would be ignored" % type(self).__name__)
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.AddkError("add_gym_b200.ADDAgent needs a CUDA device: there is no CPU path")
            device = "cuda:%d" % torch.cuda.current_device()
        _lib.lib()
        self._device = device
        self._distributed = distributed
        self._world = dist.get_world_size() if (distributed and dist.is_initialized()) else 1
        config = env_config["agent"]
        self._config = config
        self._iter = 0
        self._sample_count = 0
        self.rng = DeviceRandom(device)
        self._env = ImitationEnvironment(env_config, device)
        self._add_motion = ADDMotion(env_config["task"], self._env, device,
                                     fix_start_idx=env_config["task"].get("fix_start_idx", False))
        self._add_obs = ADDObservation(env_config["task"], self._env, self._add_motion, device)
        self._add_reward = ADDReward(env_config["task"], self._env, self._add_obs, device)
        self._add_done = ADDDone(env_config["task"], self._env, self._add_obs, self._add_motion, self._env.plane, device)
        self._core = self._add_obs.core
        self._load_params(config)
        self._build_normalizers()
        self._build_model(config)
        if self._world > 1:   # what DDP's constructor does: every rank starts from rank 0's weights
            dist.broadcast(self._model.flat, src=0)
        self._optimizer = MPOptimizerState(config["optimizer"], self._model)
        # The exchange step: one kernel over NVLink peer memory (gradient reduce-scatter + AdamW on the owned shard +
        # all-gather of the parameters, csrc/p2p.cu) instead of an NCCL all-reduce + a separate AdamW launch.  Needs every
        # rank on the same box with its own device; ADDK_P2P=0 or optimizer.grad_clip > 0 keep the NCCL path.
        self._p2p = None
        if (self._world > 1 and os.environ.get("ADDK_P2P", "1") != "0" and self._optimizer._grad_clip == 0.0
                and dist.get_backend() == "nccl" and self._world <= 8):
            self._p2p = _lib.P2PExchange(self._model.num_params, device, dist)
            self._model.rebind_storage(self._p2p.param, self._p2p.grad)
            self._optimizer.shard = self._p2p.shard()
        self._build_exp_buffer(config)
        self._build_update_ctx()
        self._mode = AgentMode.TRAIN
        self._curr_obs = None
        self._curr_info = None
        self._is_restored = False
        self._pos_diff = torch.zeros(self._add_obs.core.disc_dim, device=device, dtype=torch.float32)
        ent = self._env.robot.entity
        self._masked_engine = hasattr(ent, "set_state_masked")
        self.engine_time_events = None   # bench.py installs a list to time scene.step() separately
        self.stage_events = None         # bench.py installs a list: _train_iter appends 5 events around its 4 stages
        self._log_pending = None         # deferred logging: (iteration header, pinned host row, event)
        # CUDA graphs for the launch-bound rollout step (two segments around the physics step), see _rollout_train
        self._use_graphs = bool(config.get("cuda_graphs", True))
        self._graphs_pre, self._graphs_post = {}, {}
        self._graph_pool = None
        self._rollouts_done = 0
        self._host_fetch = getattr(ent, "fetch_from_host", None)

    model = property(lambda s: s._model)

    # ---- construction ----------------------------------------------------------------------------------------
    def _load_params(self, config):
        g = config.get
        self._discount = config["discount"]
        self._iters_per_output = config["iters_per_output"]
        self._normalizer_samples = g("normalizer_samples", np.inf)
        self._test_episodes = config["test_episodes"]
        self._steps_per_iter = config["steps_per_iter"]
        self._update_epochs = config["update_epochs"]
        self._batch_size = config["batch_size"]
        self._td_lambda = config["td_lambda"]
        self._ppo_clip_ratio = config["ppo_clip_ratio"]
        self._norm_adv_clip = config["norm_adv_clip"]
        self._action_bound_weight = config["action_bound_weight"]
        self._action_entropy_weight = config["action_entropy_weight"]
        self._action_reg_weight = config["action_reg_weight"]
        assert self._action_entropy_weight == 0 and self._action_reg_weight == 0, \
            "entropy / mean regularisers are zero in add_g1.yaml and not implemented in the fused loss"
        self._critic_loss_weight = config["critic_loss_weight"]
        self._exp_anneal_samples = g("exp_anneal_samples", np.inf)
        self._exp_prob_beg = g("exp_prob_beg", 1.0)
        self._exp_prob_end = g("exp_prob_end", 1.0)
        self._disc_loss_weight = config["disc_loss_weight"]
        self._disc_logit_reg = config["disc_logit_reg"]
        self._disc_grad_penalty = config["disc_grad_penalty"]
        self._disc_weight_decay = config["disc_weight_decay"]
        self._disc_reward_scale = config["disc_reward_scale"]
        self._task_reward_weight = config["task_reward_weight"]
        self._disc_reward_weight = config["disc_reward_weight"]
        self._update_streams = int(g("update_streams", 3))   # 3: actor / critic / discriminator chains overlap

    def _build_normalizers(self):
        dev = self._device
        self._obs_norm = Normalizer(self._add_obs.get_obs_shape(), device=dev, dtype=torch.float)
        a_space = self._env.robot.get_action_space()
        a_mean = (0.5 * (a_space[:, 1] + a_space[:, 0])).to(dev)
        a_std = (0.5 * (a_space[:, 1] - a_space[:, 0])).to(dev)
        self._a_norm = Normalizer(a_mean.shape[:1], device=dev, init_mean=a_mean, init_std=a_std, dtype=a_space.dtype)
        self._disc_obs_norm = DiffNormalizer(self._add_obs.get_disc_obs_shape(), device=dev, dtype=torch.float)

    def _build_model(self, config):
        self._model = ADDModel(config["model"], self._env, self._add_obs.get_obs_shape(),
                               self._env.robot.get_action_space(), self._add_obs.get_disc_obs_shape(),
                               device=self._device)

    def _build_exp_buffer(self, config):
        T, N, dev = self._steps_per_iter, self.get_num_envs(), self._device
        self._exp_buffer = ExperienceBuffer(T, N, dev, randperm_fn=lambda n: self.rng.randperm(n))
        od, dd, ad = self._core.obs_dim, self._core.disc_dim, self._model.act_dim
        z = lambda *s, dt=torch.float: torch.zeros(list(s), device=dev, dtype=dt)
        for name, buf in (("obs", z(T, N, od)), ("next_obs", z(T, N, od)), ("action", z(T, N, ad)),
                          ("reward", z(T, N)), ("done", z(T, N, dt=torch.int)), ("a_logp", z(T, N)),
                          ("tar_val", z(T, N)), ("adv", z(T, N)), ("rand_action_mask", z(T, N)),
                          ("disc_obs", z(T, N, dd)), ("disc_obs_demo", z(T, N, dd)),
                          ("motion_ids", z(T, N, dt=torch.long)), ("motion_times", z(T, N))):
            self._exp_buffer.add_buffer(name, buf)

    def _build_update_ctx(self):
        m, dev, N = self._model, self._device, self.get_num_envs()
        T = self._steps_per_iter
        M = int(self._batch_size * N)
        assert M <= T * N, "minibatch larger than the rollout"
        R = M + 1
        od, dd, ad = m.obs_dim, m.disc_dim, m.act_dim
        # leading dimensions: 8 elements = 16 bytes in the 16-bit twins is what TMA needs; the wide inputs use 16 elements
        # = 32 bytes, because a pitch that is an odd multiple of 16 bytes puts every other row across a sector boundary
        # (measured: the K = 264 first layer at 60 us against 44 us with a 272-element pitch)
        wo = int(os.environ.get("ADDK_OBS_LD_ALIGN", "16"))       # 8 = the minimum TMA accepts (A/B experiments)
        wd = int(os.environ.get("ADDK_DISC_LD_ALIGN", "16"))
        al, dl, ol = (ad + 7) & ~7, (dd + wd - 1) // wd * wd, (od + wo - 1) // wo * wo
        H, E = m.hidden
        # split-K slabs of the weight gradients: 9 x (16 | 32 | 16) output tiles = 144 | 288 | 144 fill the 148 SMs of the
        # persistent f16x3 kernel; 8 suits the CTA-pair tf32x3 kernel
        S = 9 if m.precision == _lib.PRECISIONS["f16x3"] else 8
        z = lambda *s, dt=torch.float32: torch.zeros(list(s), device=dev, dtype=dt)
        fb = self._exp_buffer.get_data_flat
        self._max_steps = int(self._update_epochs * int(np.ceil(float(T * N) / M)))
        w1 = max(H[0], H[1], E[0])
        # every fp32 tensor a dense layer reads or writes is carved out of ONE arena (32-byte aligned pieces), so that in
        # the "bf16" mode the library finds the bf16 twin of any operand at arena16 + (ptr - arena)
        shapes = dict(xn=(R, ol), wa0_pad=(H[0], ol), wc0_pad=(H[0], ol), an=(R, al), dn=(R, dl), h1=(R, w1), h2=(R, w1), h3=(R, max(H[2], E[1])), g1=(R, w1),
                      g2=(R, w1), g3=(R, max(H[2], E[1])), u1=(R, E[0]), u2=(R, E[1]), gx=(R, dl), dg=(R, dl), mean=(R, al),
                      dmean=(R, al), wd0_pad=(E[0], dl))
        # the critic and the discriminator chains of an optimizer step get workspaces of their own, so the library can
        # run the three chains on three streams (csrc/mlp.cu: addk_update_minibatch)
        n_streams = int(self._update_streams)
        assert n_streams in (1, 3)
        if n_streams == 3:
            shapes.update(c_h1=(R, H[0]), c_h2=(R, H[1]), c_h3=(R, H[2]), c_g1=(R, H[0]), c_g2=(R, H[1]), c_g3=(R, H[2]),
                          d_e1=(R, E[0]), d_e2=(R, E[1]), d_dh2=(R, E[1]), d_dv1=(R, E[0]), d_du2=(R, E[1]))
        offs, total = {}, 0
        for k, shp in shapes.items():
            offs[k] = total
            total += (int(np.prod(shp)) + 127) & ~127      # 128 elements: a piece's ReLU bit plane starts on a 16-byte boundary
        self._arena = z(total)
        bf16 = m.precision == _lib.PRECISIONS["bf16"]
        h3 = m.precision == _lib.PRECISIONS["f16x3"]      # two fp16 planes (hi, lo) per twin + one max|x| word per tensor
        self._arena16 = torch.zeros(2 * total if h3 else (total if bf16 else 8), device=dev, dtype=torch.bfloat16)
        self._params16 = torch.zeros(2 * m.num_params if h3 else (m.num_params if bf16 else 8), device=dev, dtype=torch.bfloat16)
        self._amax_slots = torch.zeros(2 * (1 + 4 * 128), device=dev, dtype=torch.int32) if h3 else None
        # ReLU masks as bit planes (one bit per arena element; written by the forward layers of an optimizer step)
        self._arena_bits = torch.zeros(total // 32 + 4, device=dev, dtype=torch.int32) if (h3 or bf16) else None
        carve = {k: self._arena[offs[k]:offs[k] + int(np.prod(shp))].view(shp) for k, shp in shapes.items()}
        self._ws = dict(
            carve, old_logp=z(R), adv=z(R), tar=z(R), mask=z(R), pred=z(R), dpred=z(R), ones=torch.ones(R, device=dev),
            stats=z(32, dt=torch.float64), info=z(self._max_steps, 16), cnt=z(1, dt=torch.int32),
            slabs=z(2 * S, m.num_params), colsum_work=z(128 * 1024 + 64), arena=self._arena, arena16=self._arena16,
            params16=self._params16, amax_slots=self._amax_slots, arena_bits=self._arena_bits)
        for k in ("colpart_a", "colpart_c", "colpart_d"):
            self._ws[k] = z(148 * 8, 1024) if (h3 or bf16) else None
        if n_streams == 3:
            self._ws.update(d_pred=z(R), d_dpred=z(R), colsum_work_c=z(128 * 1024 + 64), colsum_work_d=z(128 * 1024 + 64))
        else:
            for k in ("c_h1", "c_h2", "c_h3", "c_g1", "c_g2", "c_g3", "d_e1", "d_e2", "d_dh2", "d_dv1", "d_du2", "d_pred",
                      "d_dpred", "colsum_work_c", "colsum_work_d"):
                self._ws[k] = None
        ptrs = dict(self._ws)
        ptrs.update(params=m.flat, grads=m.flat_grad, exp_avg=self._optimizer.exp_avg,
                    exp_avg_sq=self._optimizer.exp_avg_sq, obs_mean=self._obs_norm._mean, obs_std=self._obs_norm._std,
                    a_mean=self._a_norm._mean, a_std=self._a_norm._std, disc_mean_abs=self._disc_obs_norm._mean_abs,
                    logstd=m._action_dist._logstd_net, buf_obs=fb("obs"), buf_action=fb("action"),
                    buf_a_logp=fb("a_logp"), buf_adv=fb("adv"), buf_tar_val=fb("tar_val"),
                    buf_mask=fb("rand_action_mask"), buf_disc_obs=fb("disc_obs"), buf_disc_demo=fb("disc_obs_demo"))
        ints = dict(obs_dim=od, obs_ld=ol, act_dim=ad, disc_dim=dd, act_ld=al, disc_ld=dl, mb_rows=M, num_params=m.num_params,
                    split_k=S, arena_elems=total, precision=m.precision, n_streams=n_streams, hid_a1=H[0], hid_a2=H[1], hid_a3=H[2], hid_d1=E[0], hid_d2=E[1],
                    params16_current=0)
        ints.update(m.offsets)
        opt = self._optimizer
        f64 = dict(ppo_clip_ratio=self._ppo_clip_ratio, action_bound_weight=self._action_bound_weight,
                   critic_loss_weight=self._critic_loss_weight, disc_loss_weight=self._disc_loss_weight,
                   disc_logit_reg=self._disc_logit_reg, disc_grad_penalty=self._disc_grad_penalty,
                   disc_weight_decay=self._disc_weight_decay, lr=opt.lr, beta1=opt.betas[0], beta2=opt.betas[1],
                   adam_eps=opt.eps, weight_decay=opt.weight_decay, grad_scale=1.0 / self._world,
                   grad_clip=opt._grad_clip)
        self._mb_rows = M
        self._ctx = _lib.UpdateCtx(ptrs, ints, f64)
        # the same context for the env steps of ONE rollout: the weights do not change between them, so the 16-bit twin
        # of the parameters and the padded first-layer weights are converted once (addk_params_refresh at the start of
        # _rollout_train) instead of on each of the 32 actor calls
        self._ctx_rollout = self._ctx.rebuild(params16_current=1)
        self._actor_ctx = self._ctx
        opt.on_hyperparams_changed = self._sync_optimizer_scalars
        # rollout scratch
        self._action = z(N, ad)
        self._a_logp = z(N)
        self._vals = z(T * N)
        self._next_vals = z(T * N)
        self._logits = z(T * N)
        self._work3 = z(3, dt=torch.float64)
        self._adv_stats = z(2)
        self._disc_r_stats = z(2)

    def _sync_optimizer_scalars(self):
        """torch.optim.AdamW.load_state_dict overrides the param_group values (mp_optimizer.py:52-53): the scalars baked
        into the update context follow the optimizer state after a checkpoint load."""
        opt = self._optimizer
        self._ctx = self._ctx.rebuild(lr=opt.lr, beta1=opt.betas[0], beta2=opt.betas[1], adam_eps=opt.eps,
                                      weight_decay=opt.weight_decay, grad_clip=opt._grad_clip)
        in_rollout = self._actor_ctx is self._ctx_rollout
        self._ctx_rollout = self._ctx.rebuild(params16_current=1)
        self._actor_ctx = self._ctx_rollout if in_rollout else self._ctx

    # ---- small reference API -------------------------------------------------------------------------------------
    def get_num_envs(self):
        return self._env.num_envs

    def get_action_size(self):
        return self._model.act_dim

    def set_mode(self, mode):
        assert mode in (AgentMode.TRAIN, AgentMode.TEST)
        self._mode = mode

    def calc_num_params(self):
        return sum(t.numel() for t in self._model.trainable()[1])

    def _need_normalizer_update(self):
        return self._sample_count < self._normalizer_samples

    def _get_exp_prob(self):
        if np.isfinite(self._exp_anneal_samples):
            l = float(np.clip(float(self._sample_count) / self._exp_anneal_samples, 0.0, 1.0))
            return (1.0 - l) * self._exp_prob_beg + l * self._exp_prob_end
        return self._exp_prob_beg

    # ---- rollout ---------------------------------------------------------------------------------------------------
    def _exp_row(self, t):
        b = self._exp_buffer.get_data
        row = _lib.AddkExpRow()
        for k, name in (("next_obs", "next_obs"), ("reward", "reward"), ("done", "done"), ("disc_obs", "disc_obs"),
                        ("disc_obs_demo", "disc_obs_demo"), ("motion_ids", "motion_ids"), ("motion_times", "motion_times")):
            setattr(row, k, b(name)[t].data_ptr())
        return row

    def _decide_action(self, obs, info, record_t=None):
        """Actor forward + Gaussian sample (TRAIN) or mode (TEST) -> (action, {"a_logp", "rand_action_mask"})."""
        N = obs.shape[0]
        train = self._mode == AgentMode.TRAIN
        noise = self.rng.action_noise(N, self._model.act_dim) if train else torch.zeros(N, self._model.act_dim, device=self._device)
        mask = self.rng.exp_mask(N, self._get_exp_prob()) if train else torch.zeros(N, device=self._device)
        b = self._exp_buffer.get_data
        rec = [None] * 4
        if record_t is not None:
            rec = [b("obs")[record_t], b("action")[record_t], b("a_logp")[record_t], b("rand_action_mask")[record_t]]
        rc = _lib.lib().addk_actor_step(
            _lib.stream(), self._actor_ctx.buf, _lib.ptr(obs), _lib.ptr(noise), _lib.ptr(mask), C.c_int(N),
            _lib.ptr(self._action), _lib.ptr(self._a_logp), *[_lib.ptr(r) for r in rec])
        _lib.check(rc, "addk_actor_step")
        m = mask if mask is not None else torch.ones(N, device=self._device)
        return self._action, {"a_logp": self._a_logp, "rand_action_mask": m}

    def _step_env(self, action, record_t=None):
        """Physics step behind the engine interface, then the fused post-step kernel."""
        env = self._env
        ev = self.engine_time_events
        env.robot.apply_action(action)
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        env.scene.step()
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        flags = _lib.F_ADVANCE | _lib.F_UPDATE_MOTION | _lib.F_REWARD_DONE
        self._core.step(flags, exp_row=self._exp_row(record_t) if record_t is not None else None)
        return self._core.obs_buf, self._core.reward, self._core.done_buf, self._add_obs.info

    def _reset_envs(self, env_ids=None):
        """Index-list API of the reference (add_agent.py:221-233); env_ids None = all."""
        core = self._core
        if env_ids is None:
            self._fill_reset_candidates(None)
            flags = torch.ones(core.N, dtype=torch.int32, device=self._device)
            core.reset(flags, reset_all=True, zero_time_done=True)
            core.done_buf.zero_()
            self._env.time_buf.zero_()
            self._write_physics(core.reset_mask)
            core.step(0, track_returns=False)
        elif len(env_ids) > 0:
            self._env.reset(env_ids)
            self._add_done.reset_idx(env_ids)
            self._add_obs.reset_idx(env_ids)
            self._add_obs.compute_obs()
        return self._add_obs.obs_buf, self._add_obs.info

    def _fill_reset_candidates(self, done):
        core = self._core
        self._add_motion.sample_time_masked(done, core.new_ids, core.new_times, uniforms=self.rng.reset_uniforms(core.N))

    def _write_physics(self, reset_mask):
        core, ent = self._core, self._env.robot.entity
        if self._masked_engine:
            ent.set_state_masked(reset_mask.bool(), core.qpos_out, core.qvel_out)
        else:   # engines without the extension: the reference's index path (one host sync)
            ids = reset_mask.nonzero(as_tuple=False).flatten()
            if len(ids) > 0:
                ent.set_qpos(core.qpos_out[ids], envs_idx=ids)
                ent.set_dofs_velocity(core.qvel_out[ids], envs_idx=ids)

    def _reset_done_envs(self, done):
        """Masked, sync-free version of BaseAgent._reset_done_envs (base_agent.py:449-453)."""
        core = self._core
        self._fill_reset_candidates(core.done_buf)
        core.reset(core.done_buf, reset_all=False, zero_time_done=True)
        self._write_physics(core.reset_mask)
        core.step(_lib.F_MASKED, env_mask=core.reset_mask, track_returns=False)
        return core.obs_buf, self._add_obs.info

    def _graphs_ok(self):
        """The rollout step is ~30 small launches and launch-bound at 4096 envs; it is replayed from CUDA graphs when
        everything it touches has a fixed address: device RNG, an engine whose getters hand out persistent tensors and
        that takes masked state writes, a constant exploration probability.  The first rollout always runs eagerly
        (lazy one-time initialisation inside the library must not happen during stream capture)."""
        ent = self._env.robot.entity
        return (self._use_graphs and self._rollouts_done >= 1 and isinstance(self.rng, DeviceRandom) and self._masked_engine
                and getattr(ent, "persistent_state_tensors", False) and self._mode == AgentMode.TRAIN
                and not np.isfinite(self._exp_anneal_samples))

    def _capture_all(self):
        """Capture every rollout-step graph up front: segment 1 per buffer row t, segment 2 per (t, history head).
        Capturing does not execute anything, so the Python-side bookkeeping it runs (history head) is restored.
        thread_local error mode: other threads (NCCL watchdog, pinned-memory allocator) keep querying events during
        capture.  The cyclic GC is paused: collecting an older agent's graphs mid-capture frees device memory, which
        invalidates the capture (seen when a second agent was built after a first one had captured its graphs)."""
        core, T, nH = self._core, self._steps_per_iter, self._core.hist.shape[1]
        self._graph_pool = torch.cuda.graph_pool_handle()
        head0 = core.hist_head
        flags = _lib.F_ADVANCE | _lib.F_UPDATE_MOTION | _lib.F_REWARD_DONE
        was_enabled = gc.isenabled()
        gc.collect()
        gc.disable()
        try:
            for t in range(T):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                    self._decide_action(self._curr_obs, self._curr_info, record_t=t)
                self._graphs_pre[t] = g
                for head in range(nH):
                    core.hist_head = head
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                        core.step(flags, exp_row=self._exp_row(t))
                        self._reset_done_envs(core.done_buf)
                    self._graphs_post[(t, head)] = g
        finally:
            core.hist_head = head0
            if was_enabled:
                gc.enable()

    def release_graphs(self):
        """Drop the captured rollout graphs (and their private memory pool) deterministically."""
        self._graphs_pre.clear()
        self._graphs_post.clear()
        self._graph_pool = None

    def _rollout_step_graphed(self, t):
        core, env = self._core, self._env
        nH = core.hist.shape[1]
        self._graphs_pre[t].replay()       # segment 1: action noise + actor forward + sample / log-prob + record row t
        ev = self.engine_time_events
        env.robot.apply_action(self._action)
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        env.scene.step()                                   # physics: never captured
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        if self._host_fetch is not None:                   # host-resident simulator: its H2D staging copy stays eager
            self._host_fetch()                             # (pinned-memory bookkeeping is not capturable)
        # segment 2: fused post-step kernel + reset candidates + masked reset + masked recompute
        self._graphs_post[(t, core.hist_head)].replay()
        core.hist_head = (core.hist_head + 1) % nH
        self._curr_obs, self._curr_info = core.obs_buf, self._add_obs.info

    # ---- reference hooks whose work is fused into the kernels ---------------------------------------------------
    def _record_data_pre_step(self, obs, info, action, action_info):
        """Reference hook (base_agent.py:437-442, ppo_agent.py:106-109).  On this path the actor kernel has already
        written obs / action / a_logp / rand_action_mask into row t of the experience buffer, and the normalizer sums
        are taken over the whole buffer at the end of the iteration: nothing is left to do here.  A subclass that
        overrides it is called once per env step, before `_exp_buffer.inc()`, with row t of those buffers (the values
        the reference passes; `obs_buf` itself has moved on by then) and may `_exp_buffer.record(...)` keys it added."""

    def _record_data_post_step(self, next_obs, r, done, next_info):
        """Reference hook (base_agent.py:444-447, amp_agent.py:52-59, add_agent.py:93-104): next_obs, reward, done,
        disc_obs, disc_obs_demo, motion_ids and motion_times of row t are written by the fused step kernel.  Overrides
        are called once per env step with the pre-reset values of row t, as in the reference."""

    def _compute_loss(self, batch):
        """Reference hook (ppo_agent.py:194-208, amp_agent.py:98-114).  The loss, its 13 diagnostics and the backward
        pass of a minibatch are ONE native call here (`addk_update_minibatch`); there is no per-batch Python loss to
        override, and silently ignoring an override would train something else than the subclass asked for."""
        raise _lib.AddkError("ADDAgent._compute_loss is fused into addk_update_minibatch on the B200 path and cannot be "
                             "overridden or called; change the loss in add_gym_b200/csrc/mlp.cu (DESIGN.md section 1)")

    def _record_hooks_overridden(self):
        cls = type(self)
        return (cls._record_data_pre_step is not ADDAgent._record_data_pre_step
                or cls._record_data_post_step is not ADDAgent._record_data_post_step)

    def _call_record_hooks(self, t):
        b = self._exp_buffer.get_data
        self._record_data_pre_step(b("obs")[t], self._add_obs.info, b("action")[t],
                                   {"a_logp": b("a_logp")[t], "rand_action_mask": b("rand_action_mask")[t]})
        self._record_data_post_step(b("next_obs")[t], b("reward")[t], b("done")[t],
                                    {"disc_obs": b("disc_obs")[t], "disc_obs_demo": b("disc_obs_demo")[t]})

    def _rollout_train(self, num_steps):
        # one conversion of the parameters' 16-bit twin / padded first-layer weights for all env steps of this rollout
        _lib.check(_lib.lib().addk_params_refresh(_lib.stream(), self._ctx.buf), "addk_params_refresh")
        self._actor_ctx = self._ctx_rollout
        try:
            graphed = self._graphs_ok()
            if graphed and not self._graphs_pre:
                self._capture_all()
            hooks = self._record_hooks_overridden()
            for _ in range(num_steps):
                t = self._exp_buffer.get_buffer_head()
                if graphed:
                    self._rollout_step_graphed(t)
                else:
                    action, _ = self._decide_action(self._curr_obs, self._curr_info, record_t=t)
                    _, _, done, _ = self._step_env(action, record_t=t)
                    self._curr_obs, self._curr_info = self._reset_done_envs(done)
                if hooks:
                    self._call_record_hooks(t)
                self._exp_buffer.inc()
        finally:
            self._actor_ctx = self._ctx
        self._rollouts_done += 1

    # ---- training data ------------------------------------------------------------------------------------------
    def _build_train_data(self):
        L, st, fb = _lib.lib(), _lib.stream, self._exp_buffer.get_data_flat
        T, N = self._steps_per_iter, self.get_num_envs()
        n = T * N
        dobs, demo = fb("disc_obs"), fb("disc_obs_demo")
        _lib.check(L.addk_disc_eval(st(), self._ctx.buf, _lib.ptr(dobs), _lib.ptr(demo), C.c_longlong(n),
                                    _lib.ptr(self._logits)), "addk_disc_eval")
        self._add_motion.sampler.update_errors(fb("motion_ids"), fb("motion_times"), dobs, demo)
        _lib.check(L.addk_disc_reward(st(), _lib.ptr(self._logits), _lib.ptr(fb("reward")), C.c_int(n),
                                      C.c_float(self._disc_reward_scale), C.c_float(self._task_reward_weight),
                                      C.c_float(self._disc_reward_weight), _lib.ptr(self._work3),
                                      _lib.ptr(self._disc_r_stats)), "addk_disc_reward")
        _lib.check(L.addk_critic_eval(st(), self._ctx.buf, _lib.ptr(fb("next_obs")), C.c_longlong(n),
                                      _lib.ptr(self._next_vals)), "addk_critic_eval")
        _lib.check(L.addk_critic_eval(st(), self._ctx.buf, _lib.ptr(fb("obs")), C.c_longlong(n), _lib.ptr(self._vals)),
                   "addk_critic_eval")
        succ = self._env.get_reward_succ() / (1.0 - self._discount)
        fail = self._env.get_reward_fail() / (1.0 - self._discount)
        _lib.check(L.addk_td_lambda(st(), _lib.ptr(fb("reward")), _lib.ptr(self._next_vals), _lib.ptr(self._vals),
                                    _lib.ptr(fb("done")), C.c_int(T), C.c_int(N), C.c_float(self._discount),
                                    C.c_float(self._td_lambda), C.c_float(succ), C.c_float(fail), _lib.ptr(fb("tar_val")),
                                    _lib.ptr(fb("adv"))), "addk_td_lambda")
        _lib.check(L.addk_adv_normalize(st(), _lib.ptr(fb("adv")), _lib.ptr(fb("rand_action_mask")), C.c_int(n),
                                        C.c_float(self._norm_adv_clip), _lib.ptr(self._work3), _lib.ptr(self._adv_stats)),
                   "addk_adv_normalize")
        return {"adv_mean": self._adv_stats[0], "adv_std": self._adv_stats[1],
                "disc_reward_mean": self._disc_r_stats[0], "disc_reward_std": self._disc_r_stats[1]}

    def _update_model(self):
        L = _lib.lib()
        N = self.get_num_envs()
        num_samples = self._exp_buffer.get_sample_count()
        M = self._mb_rows
        num_batches = int(np.ceil(float(num_samples) / M))
        step = 0
        opt = self._optimizer
        for _ in range(self._update_epochs):
            for _ in range(num_batches):
                idx = self._exp_buffer.sample_indices(M)
                local = self._world == 1
                rc = L.addk_update_minibatch(_lib.stream(), self._ctx.buf, _lib.ptr(idx), C.c_int(step),
                                             C.c_int(opt.steps + 1 if local else 0))
                _lib.check(rc, "addk_update_minibatch")
                if self._p2p is not None:   # the one exchange step of the path, fused with the optimizer (csrc/p2p.cu)
                    self._p2p.step(opt.exp_avg, opt.exp_avg_sq, opt.steps + 1, opt.lr, opt.betas, opt.eps, opt.weight_decay)
                elif not local:   # ... or as an NCCL all-reduce of the flat gradient + AdamW
                    dist.all_reduce(self._model.flat_grad, op=dist.ReduceOp.SUM)
                    if opt._grad_clip > 0.0:   # global-norm clip of the rank-averaged gradient (mp_optimizer.py:19-20)
                        rc = L.addk_clip_grad_norm(_lib.stream(), _lib.ptr(self._model.flat_grad),
                                                   C.c_longlong(self._model.num_params), C.c_double(opt._grad_clip),
                                                   C.c_double(1.0 / self._world), _lib.ptr(self._ws["stats"][30:]),
                                                   _lib.ptr(self._ws["info"][step, 14:]))
                        _lib.check(rc, "addk_clip_grad_norm")
                    rc = L.addk_adamw(_lib.stream(), _lib.ptr(self._model.flat), _lib.ptr(self._model.flat_grad),
                                      _lib.ptr(opt.exp_avg), _lib.ptr(opt.exp_avg_sq), C.c_longlong(self._model.num_params),
                                      C.c_int(opt.steps + 1), C.c_double(opt.lr), C.c_double(opt.betas[0]),
                                      C.c_double(opt.betas[1]), C.c_double(opt.eps), C.c_double(opt.weight_decay),
                                      C.c_double(1.0 / self._world))
                    _lib.check(rc, "addk_adamw")
                opt.steps += 1
                step += 1
        info = self._ws["info"][:step].mean(dim=0)
        return {k: info[i] for i, k in enumerate(INFO_KEYS)}

    def _update_normalizers(self):
        fb = self._exp_buffer.get_data_flat
        self._obs_norm.record(fb("obs"))
        self._obs_norm.update()
        self._disc_obs_norm.record_pair(fb("disc_obs_demo"), fb("disc_obs"))
        self._disc_obs_norm.update()

    def _train_iter(self):
        """One training iteration (base_agent.py:353-374).  Nothing in here synchronises with the host: every value of the
        returned dict is a 0-d DEVICE tensor (the reference returns Python floats for the tracker entries and pays
        `.item()` syncs for them, base_agent.py:603-621).  `stage_events`, when a list, receives five CUDA events around
        the four stages (bench.py's per-stage timing)."""
        marks = self.stage_events
        ev = (lambda: None) if marks is None else (lambda: marks.append(torch.cuda.Event(enable_timing=True)) or marks[-1].record())
        self.set_mode(AgentMode.TRAIN)
        ev()
        self._rollout_train(self._steps_per_iter)
        ev()
        data_info = self._build_train_data()
        ev()
        train_info = self._update_model()
        ev()
        if self._need_normalizer_update():
            self._update_normalizers()
        ev()
        info = {**train_info, **data_info}
        info.update(self._tracker_info())
        return info

    def _tracker_info(self):
        """ReturnTracker.get_mean_return / get_mean_ep_len / get_episodes (base_agent.py:576-590) as device scalars."""
        c = self._core
        cnt = c.tracker_count.to(torch.float64)
        mean = c.tracker_sums / torch.clamp(cnt, min=1.0)
        return {"mean_return": mean[0], "mean_ep_len": mean[1], "num_eps": cnt[0]}

    def _reset_tracker(self):
        c = self._core
        for t in (c.tracker_sums, c.tracker_count, c.return_buf, c.ep_len_buf, c.eps_per_env):
            t.zero_()

    # ---- outer loops ---------------------------------------------------------------------------------------------
    def train_model(self, out_model_file=None, int_output_dir="", log_file=None, max_iters=None):
        """BaseAgent.train_model (base_agent.py:79-113).  Like the reference, the iteration that reaches the stop
        condition (`max_samples`; here also `max_iters`) is forced to be an output iteration: final test rollout, log
        flush and checkpoint, so the saved model is never stale."""
        max_samples = self._config.get("max_samples", int(1e6))
        start = time.time()
        self._curr_obs, self._curr_info = self._reset_envs()
        if not self._is_restored:
            self._iter, self._sample_count = 0, 0
        self._exp_buffer.clear()
        self._reset_tracker()
        test_info = None
        while self._sample_count < max_samples and (max_iters is None or self._iter < max_iters):
            output_iter = self._iter % self._iters_per_output == 0
            if output_iter:
                test_info = self.test_model(self._test_episodes)
            info = self._train_iter()
            self._sample_count = self._exp_buffer.get_total_samples()
            last = self._sample_count >= max_samples or (max_iters is not None and self._iter + 1 >= max_iters)
            if last:
                output_iter = True
                test_info = self.test_model(self._test_episodes)
            self._log(info, test_info, start)
            if output_iter:
                self._flush_log()
                if out_model_file and (not self._distributed or dist.get_rank() == 0):
                    self.save(out_model_file)
                    if int_output_dir:
                        self.save(os.path.join(int_output_dir, "model_{:010d}.pt".format(self._iter)))
                self._reset_tracker()
                self._curr_obs, self._curr_info = self._reset_envs()
            self._iter += 1
        self._flush_log()

    def test_model(self, num_episodes):
        """Deterministic-action evaluation over all envs until every env finished
        ceil(num_episodes / num_envs) episodes (base_agent.py:116-126,393-425)."""
        self.set_mode(AgentMode.TEST)
        self._reset_tracker()
        self._curr_obs, self._curr_info = self._reset_envs()
        if num_episodes == 0:
            info = {"mean_return": 0.0, "mean_ep_len": 0.0, "num_eps": 0}
        else:
            min_eps = int(np.ceil(num_episodes / self.get_num_envs()))
            while True:
                action, _ = self._decide_action(self._curr_obs, self._curr_info)
                flags = _lib.F_ADVANCE | _lib.F_UPDATE_MOTION | _lib.F_REWARD_DONE
                self._env.robot.apply_action(action)
                self._env.scene.step()
                self._core.step(flags)
                self._curr_obs, self._curr_info = self._reset_done_envs(self._core.done_buf)
                if bool(torch.all(self._core.eps_per_env > min_eps - 1)):
                    break
            info = {k: float(v) for k, v in self._tracker_info().items()}
        self._reset_tracker()
        self.set_mode(AgentMode.TRAIN)
        return info

    LOG_KEYS = ("mean_return", "mean_ep_len", "num_eps", "loss", "critic_loss", "actor_loss", "disc_loss", "disc_reward_mean",
                "adv_mean", "adv_std")

    def _log(self, info, test_info, start):
        """Logger row of one iteration (base_agent.py:465-520, util/logger.py:160-184) WITHOUT serialising the iteration:
        the row is stacked on the device, averaged over the ranks like the reference's Logger does (one small all-reduce),
        copied to pinned host memory asynchronously and printed when the NEXT iteration is logged (or at a flush)."""
        row = torch.stack([torch.as_tensor(info[k], device=self._device).to(torch.float64).reshape(()) for k in self.LOG_KEYS])
        if self._distributed and self._world > 1:
            dist.all_reduce(row, op=dist.ReduceOp.SUM)
            row = row / self._world
        host = torch.empty(row.shape, dtype=row.dtype, pin_memory=True)
        host.copy_(row, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._flush_log()        # the previous iteration's row: its copy finished an iteration ago
        t = test_info or {}
        head = "iter %d samples %d wall %.3fh test_return %.4f" % (
            self._iter, self._sample_count, (time.time() - start) / 3600.0, float(t.get("mean_return", 0.0)))
        self._log_pending = (head, host, ev)

    def _flush_log(self):
        if self._log_pending is None:
            return
        head, host, ev = self._log_pending
        self._log_pending = None
        ev.synchronize()
        if self._distributed and dist.get_rank() != 0:
            return
        v = dict(zip(self.LOG_KEYS, host.tolist()))
        print("%s train_return %.4f loss %.5f disc_reward %.4f" % (head, v["mean_return"], v["loss"], v["disc_reward_mean"]),
              flush=True)

    # ---- checkpoints (base_agent.py:148-208) -------------------------------------------------------------------------
    def save(self, out_file):
        torch.save({"model": self.state_dict(), "optimizer": self._optimizer.state_dict(), "iter": self._iter,
                    "sample_count": self._sample_count}, out_file)

    def load(self, in_file):
        ckpt = torch.load(in_file, map_location=self._device)
        self._is_restored = True
        if "model" in ckpt and "optimizer" in ckpt:
            sd = ckpt["model"]
            self._optimizer.load_state_dict(ckpt["optimizer"])
            self._iter = ckpt.get("iter", 0)
            self._sample_count = ckpt.get("sample_count", 0)
        else:
            sd, self._is_restored = ckpt, False
        sd = {k.replace("_model.module.", "_model."): v for k, v in sd.items()}   # DDP-prefixed reference checkpoints
        self.load_state_dict(sd)
        self._obs_norm._mean_sq = None
