"""GPU parity: the CUDA path (through the C-ABI, via the drop-in classes) against the CPU oracle and against
the golden vectors produced by the executed reference.

Bars (BASELINE.json north_star): bit-exact motion-frame indices, done / reset masks, motion ids and minibatch
permutations; norm-wise relative error <= 1e-5 (fp32 MLPs) on observations, rewards, advantages, the loss and
the gradients.  Tolerances are spelled out at each assert.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from add_gym_b200 import config as b200_config

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP32_TOL = 1e-5
THREE_CLIPS = os.path.join(b200_config.ASSET_DIR, "three_clips.yaml")
SEVEN_CLIPS = os.path.join(b200_config.ASSET_DIR, "seven_clips.yaml")
ALL_CLIPS = os.path.join(b200_config.ASSET_DIR, "motions_all.addkc")
FLOAT_KEYS = ["obs", "next_obs", "action", "reward", "a_logp", "disc_obs", "disc_obs_demo", "motion_times"]
EXACT_KEYS = ["done", "rand_action_mask", "motion_ids"]


def _need_legacy(precision):
    from add_gym_b200 import _lib
    if precision in ("tf32x3", "tf32") and not _lib.has_legacy_kernels():
        pytest.skip("precision %s: superseded kernels, only in libaddk_legacy.so (make LEGACY=1, ADDK_LIB=...)" % precision)


def _lib_mod():
    from add_gym_b200 import _lib
    return _lib


def _gpu_relu_masks(agent):
    """ReLU masks (activation > 0) of the optimizer step that just ran, read from the three chains' workspaces
    (update_streams = 3: every chain keeps its own activations).  Keys as OracleAgent.loss(masks=...)."""
    ws, M, m = agent._ws, agent._mb_rows, agent._model
    H, E = m.hidden
    assert agent._ctx.ints["n_streams"] == 3

    # bf16 mode: h1 / h2 / e1 of every chain exist only as their bf16 copies (same element offsets in the 16-bit arena)
    # (when every consumer can read them: at least 64 rows per split-K slab of the weight gradients, csrc/mlp.cu is_16only)
    # f16x3 mode: the same tensors exist only as their fp16 planes (hi plane at the same element offsets; the 16-bit arena is
    # typed bfloat16 but holds fp16 bit patterns there: "> 0" is a sign / zero test on the int16 view in both cases)
    P = _lib_mod().PRECISIONS
    mode16 = m.precision in (P["bf16"], P["f16x3"]) and M >= 64 * agent._ctx.ints["split_k"]
    only16 = {"h1", "h2", "c_h1", "c_h2", "d_e1"} if mode16 else set()

    def g(key, width, r0, r1):
        t = ws[key]
        if key in only16:
            off = (t.data_ptr() - agent._arena.data_ptr()) // 4
            t = agent._arena16[off:off + t.numel()].view(torch.int16)
        return (t.flatten()[:(M + 1) * width].view(M + 1, width)[r0:r1] > 0).cpu()
    return {"actor": [g("h1", H[0], 0, M), g("h2", H[1], 0, M), g("h3", H[2], 0, M)],
            "critic": [g("c_h1", H[0], 0, M), g("c_h2", H[1], 0, M), g("c_h3", H[2], 0, M)],
            "disc": [g("d_e1", E[0], 0, M), g("d_e2", E[1], 0, M)],
            "disc_pos": [g("d_e1", E[0], M, M + 1), g("d_e2", E[1], M, M + 1)]}


def _device_joint_rot(lib):
    return [j.cpu() for j in lib._frame_joint_rot]


def _pair(num_envs, motion=None, fall_prob=0.01, precision="fp32", task_overrides=None, grad_clip=None, engine=None):
    """(oracle agent with recorded randomness, CUDA agent replaying it) over the same synthetic physics stream.

    The oracle's motion library is built from the DEVICE-computed 30 fps joint rotations (stage A of the table
    build, checked on its own in test_motion_table_*): the reference's resampling has discontinuous branches
    (slerp's |sin| < 1e-3 midpoint rule, the 1e-5 axis-angle cut-off) that amplify a 1-ulp cos/sin difference
    between libm flavours into 1e-3 jumps, so the branchy stage is compared on identical inputs."""
    from add_gym_b200.add_agent import ADDAgent
    from oracle import harness
    import parity_helpers as helpers
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    gcfg = b200_config.default_config(num_envs=num_envs, motion_file=motion, mlp_precision=precision)
    gcfg["task"].update(task_overrides or {})
    if grad_clip is not None:
        gcfg["agent"]["optimizer"]["grad_clip"] = grad_clip
    gcfg["engine"].update(seed=1234, noise_device="cpu", fall_prob=fall_prob)
    if engine is not None:
        gcfg["engine"]["_target_"] = engine
    torch.manual_seed(0)
    agent = ADDAgent(gcfg, device="cuda:0")
    cfg = b200_config.default_config(num_envs=num_envs, motion_file=motion, mlp_precision=precision)
    if engine is not None:
        cfg["engine"]["_target_"] = engine
    cfg["task"].update(task_overrides or {})
    if grad_clip is not None:
        cfg["agent"]["optimizer"]["grad_clip"] = grad_clip
    olib = harness.make_oracle_lib(cfg, jrot_override=_device_joint_rot(agent._add_motion.motion_lib))
    rec = helpers.RecordRandom()
    oracle = harness.make_oracle_agent(num_envs, seed=0, engine_seed=1234, cfg=cfg, rng=rec, fall_prob=fall_prob,
                                       lib=olib)
    helpers.load_oracle_weights(agent, oracle)
    replay = helpers.ReplayRandom(rec, oracle.trace, "cuda:0", first_reset=0, first_noise=0, first_perm=2)
    helpers.install_replay(agent, replay)
    return oracle, agent, rec


def _start(oracle, agent):
    oracle.start()
    agent._curr_obs, agent._curr_info = agent._reset_envs()
    agent._exp_buffer.clear()
    agent._reset_tracker()


def _check_buffers(agent, ref, keys_float=FLOAT_KEYS, keys_exact=EXACT_KEYS, tol=FP32_TOL):
    from parity_helpers import rel_err
    for k in keys_exact:
        got = agent._exp_buffer.get_data(k).cpu()
        assert torch.equal(got.to(ref[k].dtype), ref[k]), "%s must be bit-exact" % k
    for k in keys_float:
        e = rel_err(agent._exp_buffer.get_data(k), ref[k])
        assert e <= tol, "%s: rel err %.3e > %.0e" % (k, e, tol)


# ---------------------------------------------------------------------------------------------------------
# motion table + gather
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("motion", [None, THREE_CLIPS, SEVEN_CLIPS, ALL_CLIPS])
def test_motion_table_matches_oracle_and_golden(motion):
    from add_gym_b200.env import ImitationEnvironment
    from add_gym_b200.add_motion import ADDMotion
    from oracle import harness
    from parity_helpers import rel_err
    cfg = b200_config.default_config(num_envs=4, motion_file=motion)
    env = ImitationEnvironment(cfg, "cuda:0")
    lib = ADDMotion(cfg["task"], env, "cuda:0").motion_lib
    D, h = lib._num_dofs, lib._row_stride // 2
    tab = lib.step_table.cpu()
    got = torch.cat([tab[:, :7 + D], tab[:, h:h + 6 + D]], dim=1)
    # ---- stage A (hinge angle -> joint quaternion, 30 fps): smooth, so elementwise: a few ulp of 1.0
    ocpu = harness.make_oracle_lib(cfg)
    for m, (jd, jo) in enumerate(zip(_device_joint_rot(lib), ocpu.frame_joint_rot)):
        assert jd.shape == jo.shape
        diff = (jd - jo).abs()
        bad = ~(diff <= 3e-7)                       # (NaN counts as a mismatch)
        assert not bool(bad.any()), "clip %d: %d of %d stage-A entries off (max %.3e, NaNs %d, first at %s)" % (
            m, int(bad.sum()), bad.numel(), float(diff[~torch.isnan(diff)].max()), int(torch.isnan(jd).sum()),
            tuple(int(x) for x in bad.nonzero()[0]))
    # ---- stage B (100 Hz resampling: lerp / slerp / twist angle / frame velocities) on identical stage-A input
    olib = harness.make_oracle_lib(cfg, jrot_override=_device_joint_rot(lib))
    assert got.shape == olib.table.shape
    assert torch.equal(lib._frame_idx.cpu(), olib.frame_idx), "source-frame indices of every 100 Hz step: bit-exact"
    assert torch.equal(lib._motion_start_idx.cpu(), olib.start_idx)            # quirk Q2 (30 fps cumsum)
    assert torch.equal(lib._motion_lengths.cpu(), olib.lengths)
    assert rel_err(got, olib.table) <= FP32_TOL
    assert float((got - olib.table).abs().max()) <= 2e-5          # rad, m, rad/s, m/s
    # ---- against the executed reference's table (golden, torch-CPU cos/sin): the reference's own branch
    # discontinuities make a small set of entries libm-dependent; everything else must agree to 1e-5.
    if motion == ALL_CLIPS:      # BASELINE configs[2]: all 42 clips, 906,203 rows (261 MB in HBM); no golden table of that size
        assert len(olib.lengths) == 42 and got.shape[0] == 906203
        return
    case = {None: "walk_n12", THREE_CLIPS: "three_clips_n10", SEVEN_CLIPS: "seven_clips_n14"}[motion]
    g = np.load(os.path.join(GOLD, case + ".npz"))
    assert list(g["table_shape"]) == list(got.shape)
    gs, ds = torch.from_numpy(g["table_sample"]), got[torch.from_numpy(g["table_rows"])]
    fragile = (gs - ds).abs() > 2e-5
    assert float(fragile.float().mean()) <= 1e-3, "branch-flip entries must stay a < 0.1 % minority"
    assert float((gs - ds).abs().max()) <= 5e-3
    assert rel_err(torch.where(fragile, gs, ds), gs) <= FP32_TOL
    np.testing.assert_allclose(got.double().sum(0).numpy(), g["table_colsum"], rtol=1e-4, atol=5e-2)


def test_motion_gather_indices_bit_exact():
    """get_precomputed_motion_step: trunc(t*100) with fp32 rounding, global clip, Q2 start offsets; edge times."""
    from add_gym_b200.env import ImitationEnvironment
    from add_gym_b200.add_motion import ADDMotion
    from oracle import harness
    cfg = b200_config.default_config(num_envs=4, motion_file=THREE_CLIPS)
    env = ImitationEnvironment(cfg, "cuda:0")
    motion = ADDMotion(cfg["task"], env, "cuda:0")
    lib = motion.motion_lib
    olib = harness.make_oracle_lib(cfg)
    g = torch.Generator().manual_seed(5)
    n = 20000
    ids = torch.randint(0, 3, (n,), generator=g)
    times = torch.rand(n, generator=g) * 130.0
    # iterated fp32 accumulation, like env.time_buf (quirk Q3)
    acc = torch.zeros(2000)
    cur = torch.zeros(1)
    for i in range(2000):
        cur = cur + 0.01
        acc[i] = cur[0]
    times[:2000] = acc
    ids[:2000] = 0
    times[2000:2010] = torch.tensor([0.0, -1.0, 1e9, 0.00999, 0.01, 0.019999, 124.16, 124.17, 200.0, 1e-30])
    ids[2000:2010] = 0
    out = lib.get_precomputed_motion_step(ids.cuda(), times.cuda(), return_index=True)
    oidx = olib.rows(ids, times)
    assert torch.equal(out[6].cpu(), oidx)
    D, h = lib._num_dofs, lib._row_stride // 2
    tab = lib.step_table.cpu()
    safe = oidx.clamp(0, tab.shape[0] - 1)
    for k, (a, b) in enumerate(zip(out[:6], [tab[safe, 0:3], tab[safe, 3:7], tab[safe, h:h + 3], tab[safe, h + 3:h + 6],
                                             tab[safe, 7:7 + D], tab[safe, h + 6:h + 6 + D]])):
        assert torch.equal(a.cpu(), b), "gather output %d is a pure copy of the table row" % k
    # empty request
    e = lib.get_precomputed_motion_step(torch.zeros(0, dtype=torch.long, device="cuda"), torch.zeros(0, device="cuda"))
    assert e[0].shape == (0, 3)


def test_calc_motion_frame_at_arbitrary_times():
    """MotionLib.calc_motion_frame (motion_lib.py:61-88): runtime frame interpolation -- root lerp, root / joint slerp,
    twist angle, frame-i0 velocities, loop offset -- at 4000 arbitrary (clip, time) queries on the seven-clip library
    with two WRAP clips, against the golden outputs of the executed reference (which the oracle reproduces bit for bit,
    tests/test_oracle_golden.py).  Frame selection is exact; the blended values agree to 1e-5 outside the reference's own
    branch discontinuities (slerp's |sin| < 1e-3 rule, the 1e-5 axis-angle cut-off; see the table test)."""
    from add_gym_b200.env import ImitationEnvironment
    from add_gym_b200.add_motion import ADDMotion
    from parity_helpers import rel_err
    g = np.load(os.path.join(GOLD, "motion_frame_queries.npz"))
    cfg = b200_config.default_config(num_envs=2, motion_file=SEVEN_CLIPS)
    env = ImitationEnvironment(cfg, "cuda:0")
    lib = ADDMotion(cfg["task"], env, "cuda:0").motion_lib
    lib._motion_loop_modes.copy_(torch.from_numpy(g["loop_modes"]).to(lib._motion_loop_modes.dtype))
    out = lib.calc_motion_frame(torch.from_numpy(g["ids"]).cuda(), torch.from_numpy(g["times"]).cuda())
    names = ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_pos", "dof_vel")
    for k, v in zip(names, out):
        ref = torch.from_numpy(g[k])
        got = v.cpu()
        assert got.shape == ref.shape, k
        if k in ("root_vel", "root_ang_vel", "dof_vel"):
            # un-blended copies of frame i0's finite-difference velocities: a wrong frame index would show as O(1) jumps
            assert float((got - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), k
        fragile = (got - ref).abs() > 2e-5
        assert float(fragile.float().mean()) <= 2e-3, (k, float(fragile.float().mean()))
        assert rel_err(torch.where(fragile, ref, got), ref) <= FP32_TOL, k
    # empty request and the baked-table refusal
    e = lib.calc_motion_frame(torch.zeros(0, dtype=torch.long, device="cuda"), torch.zeros(0, device="cuda"))
    assert e[4].shape == (0, 29, 4)


def test_baked_step_table_is_a_drop_in_for_the_built_one(tmp_path):
    """MotionLib.save_table -> MotionLib("lib.addkt"): same bytes in HBM, same lookups (incl. the Q2 start
    offsets of a multi-clip library), same lengths / weights / loop modes, and an agent constructed on it runs."""
    import copy
    from add_gym_b200.env import ImitationEnvironment
    from add_gym_b200.add_motion import ADDMotion
    cfg = b200_config.default_config(num_envs=4, motion_file=THREE_CLIPS)
    env = ImitationEnvironment(cfg, "cuda:0")
    lib = ADDMotion(cfg["task"], env, "cuda:0").motion_lib
    baked_path = str(tmp_path / "three.addkt")
    lib.save_table(baked_path)
    cfg2 = copy.deepcopy(cfg)
    cfg2["task"]["motion_file"] = baked_path
    env2 = ImitationEnvironment(cfg2, "cuda:0")
    lib2 = ADDMotion(cfg2["task"], env2, "cuda:0").motion_lib
    assert torch.equal(lib.step_table, lib2.step_table)
    from add_gym_b200 import _lib as _l
    with pytest.raises(_l.AddkError):                      # a baked table carries no 30 fps source frames
        lib2.calc_motion_frame(torch.zeros(1, dtype=torch.long, device="cuda"), torch.zeros(1, device="cuda"))
    for a, b in ((lib._motion_lengths, lib2._motion_lengths), (lib._motion_weights, lib2._motion_weights),
                 (lib._motion_loop_modes, lib2._motion_loop_modes), (lib._motion_start_idx, lib2._motion_start_idx),
                 (lib._true_start_idx, lib2._true_start_idx), (lib._motion_num_frames, lib2._motion_num_frames)):
        assert torch.equal(a, b)
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(0, 3, (5000,), generator=g).cuda()
    times = (torch.rand(5000, generator=g) * 130.0).cuda()
    for x, y in zip(lib.get_precomputed_motion_step(ids, times, return_index=True),
                    lib2.get_precomputed_motion_step(ids, times, return_index=True)):
        assert torch.equal(x, y)
    from add_gym_b200.add_agent import ADDAgent
    acfg = b200_config.default_config(num_envs=8, motion_file=baked_path)
    acfg["agent"]["steps_per_iter"] = 4
    acfg["engine"].update(seed=3, noise_device="device", fall_prob=0.01)
    agent = ADDAgent(acfg, device="cuda:0")
    agent._curr_obs, agent._curr_info = agent._reset_envs()
    agent._exp_buffer.clear()
    agent._rollout_train(4)
    assert bool(torch.isfinite(agent._exp_buffer.get_data("obs")).all())


# ---------------------------------------------------------------------------------------------------------
# one full iteration: rollout -> train data -> 40 optimizer steps -> normalizers
# ---------------------------------------------------------------------------------------------------------
def _iteration_parity(num_envs, motion, gold_case=None, task_overrides=None, steps_synced=None, precision="fp32",
                      tol=FP32_TOL, reduced_grad_tol=2e-2, grad_clip=None):
    from parity_helpers import rel_err
    from add_gym_b200 import _lib
    _need_legacy(precision)
    oracle, agent, rec = _pair(num_envs, motion, task_overrides=task_overrides, precision=precision, grad_clip=grad_clip)
    _start(oracle, agent)
    assert rel_err(agent._curr_obs, oracle.curr_obs) <= FP32_TOL
    assert torch.equal(agent._add_obs._motion_ids.cpu(), oracle.motion_ids)
    assert torch.equal(agent._add_obs._motion_time_offsets.cpu(), oracle.offsets)
    # ---- rollout
    oracle.rollout()
    agent._rollout_train(agent._steps_per_iter)
    _check_buffers(agent, oracle.buf, keys_float=[k for k in FLOAT_KEYS if k not in ("action", "a_logp")])
    _check_buffers(agent, oracle.buf, keys_float=["action", "a_logp"], keys_exact=[], tol=tol)
    assert torch.equal(agent._env.time_buf.cpu(), oracle.time_buf), "time_buf is an iterated fp32 sum: bit-exact"
    assert int(agent._core.tracker_count.item()) == oracle.episodes
    if oracle.episodes:
        assert abs(agent._core.tracker_sums[0].item() - oracle.ep_sum) <= 1e-4 * max(1.0, abs(oracle.ep_sum))
        assert agent._core.tracker_sums[1].item() == oracle.len_sum
    if gold_case is not None:
        g = np.load(os.path.join(GOLD, gold_case + ".npz"))
        # the golden buffers come from the executed reference with torch-CPU cos/sin in the table's stage A: the
        # libm-dependent branch flips (see _pair) reach a few observation entries -> exact keys still bit-exact,
        # float keys to 1e-4 norm-wise here and to 1e-5 against the oracle on identical stage-A input above.
        gold = {k: torch.from_numpy(g["buf/" + k]) for k in FLOAT_KEYS + EXACT_KEYS}
        # ("reward" is overwritten by the discriminator reward in _build_train_data: compared further down)
        _check_buffers(agent, gold, keys_float=[k for k in FLOAT_KEYS if k != "reward"], tol=1e-4)
    # ---- train data
    od = oracle.build_train_data()
    info = agent._build_train_data()
    assert rel_err(agent._logits, od["logits"]) <= tol
    keep = (oracle.buf["done"] != 1) & (oracle.buf["done"] != 2)       # SUCC / FAIL rows are overwritten by 0 later
    nv = agent._next_vals.view(oracle.T, -1).cpu()
    assert rel_err(nv[keep], od["next_vals"][keep]) <= tol
    assert rel_err(agent._vals.view(oracle.T, -1), od["vals"]) <= tol
    _check_buffers(agent, oracle.buf, keys_float=["reward", "tar_val", "adv"], keys_exact=[], tol=tol)
    if gold_case is not None:
        gold = {k: torch.from_numpy(g["buf/" + k]) for k in ("reward", "tar_val", "adv")}
        _check_buffers(agent, gold, keys_float=["reward", "tar_val", "adv"], keys_exact=[], tol=1e-4)
    for k in ("adv_mean", "adv_std", "disc_reward_mean", "disc_reward_std"):
        assert abs(float(info[k]) - float(od[k])) <= tol * max(1.0, abs(float(od[k]))), k
    assert rel_err(agent._add_motion.sampler.errors, oracle.errors) <= FP32_TOL
    # ---- optimizer steps, parameter-synchronised: same weights in, compare loss / gradient / weights out.
    # Two effects bound what any second fp32 implementation can reproduce, and the asserts MEASURE them:
    #  (1) ReLU boundary flips: a hidden unit whose pre-activation is within fp32 rounding of zero takes the other
    #      branch (the reference run on CPU vs on GPU differs the same way).  Every step the test reads the ReLU masks the
    #      CUDA path used (activation > 0 in its workspaces) and the oracle's own, counts the units that differ, and
    #      recomputes the oracle's gradient with the CUDA path's masks FORCED (OracleAgent.loss(masks=...)): with equal
    #      masks EVERY step must meet the bar on every gradient tensor; a step without flips must meet it as it is; and
    #      the unforced gradient may be off by at most FLIP_GRAD per flipped unit.
    #  (2) AdamW divides by sqrt(v): in the first steps the update is lr*sign(g), so an entry whose gradient is rounding
    #      noise can move by 2*lr either way.  Weights are compared norm-wise (5e-5) and entrywise (|d| <= 4*lr).
    L = _lib.lib()
    names = oracle.names
    gparams = dict(agent._model.named_parameters())
    snap = {}
    lr = float(agent._optimizer.lr)
    M = agent._mb_rows
    # one flipped unit removes / adds one row-unit's contribution to its layer's gradients: ~1/sqrt(#active row-units) of the
    # tensor's norm (256 active units per row is a lower bound for these nets), with a factor for heavy rows
    FLIP_GRAD = 8.0 / np.sqrt(M * 256.0)
    fp32_class = tol <= FP32_TOL
    oracle.record_masks = True

    def grad_hook(grads):
        snap["pre"] = {k: oracle.params[k].detach().clone() for k in names}

    INFO = ["loss", "critic_loss", "actor_loss", "clip_frac", "imp_ratio", "action_bound_loss", "disc_loss",
            "disc_grad_penalty", "disc_logit_loss", "disc_pos_acc", "disc_neg_acc", "disc_pos_logit", "disc_neg_logit"]
    report = []

    def grad_excess(got, ref):
        diff = float(torch.linalg.norm(got.detach().double().cpu().flatten() - ref.double().flatten()))
        # 1e-7 * sqrt(numel): fp32 epsilon on O(1) summands -- the floor for few-element tensors that are sums of
        # cancelling terms (the logit bias gradient is ONE number)
        return max(0.0, diff - 1e-7 * np.sqrt(ref.numel())) / max(float(torch.linalg.norm(ref.double())), 1e-30)

    def on_step(step, idx, oinfo, o):
        for k in names:
            gparams[k].data.copy_(snap["pre"][k])
        opt = agent._optimizer
        gidx = idx.cuda().contiguous()
        _lib.check(L.addk_update_minibatch(_lib.stream(), agent._ctx.buf, _lib.ptr(gidx), C.c_int(step),
                                           C.c_int(opt.steps + 1)), "addk_update_minibatch")
        opt.steps += 1
        row = agent._ws["info"][step].cpu()
        worst_info = max(abs(float(row[i]) - float(oinfo[k])) / max(1.0, abs(float(oinfo[k]))) for i, k in enumerate(INFO))
        worst_grad, per_tensor = 0.0, {}
        for k in names:
            ref = o.params[k].grad
            per_tensor[k] = grad_excess(gparams[k].grad, ref)
            worst_grad = max(worst_grad, per_tensor[k])
            d = (gparams[k].detach().cpu() - o.params[k].detach()).double().flatten()
            # entries whose gradient is rounding noise (|g| below 1e-3 of the tensor's RMS gradient: dead or nearly dead
            # units) get lr * sign(noise) from AdamW's first steps: they are only held to the entrywise bound below
            solid = (ref.abs() > 1e-3 * ref.double().pow(2).mean().sqrt()).flatten()
            pd = float(torch.linalg.norm(d[solid]))
            # 5e-5 of the weight norm + an RMS entry difference of 5 % of lr (zero-initialised biases have no norm yet)
            pbound = (5e-5 * float(torch.linalg.norm(o.params[k].detach().double())) + 0.05 * lr * np.sqrt(ref.numel())) * max(1.0, tol / FP32_TOL / 20)
            assert pd <= pbound, "step %d param %s: |d| %.3e > %.3e" % (step, k, pd, pbound)
            dmax = float(d.abs().max())
            assert dmax <= 4.0 * lr, "step %d param %s: max |d| %.3e" % (step, k, dmax)
        if not fp32_class:
            print("step %d gradient errors per tensor: %s" % (step, {k.replace("_layers", "").replace(".weight", ".w").replace(".bias", ".b"): "%.1e" % v for k, v in per_tensor.items()}))
        # ReLU masks the CUDA path used vs the oracle's own: count the units that differ, then recompute the oracle's
        # loss and gradient at the pre-step weights with the CUDA path's masks forced
        gm = _gpu_relu_masks(agent)
        om = o.last_masks
        flips = sum(int((a != b).sum()) for k in gm for a, b in zip(gm[k], om[k]))
        units = sum(a.numel() for k in gm for a in gm[k])
        if flips:
            post = {k: o.params[k].detach().clone() for k in names}
            with torch.no_grad():
                for k in names:
                    o.params[k].copy_(snap["pre"][k])
            finfo = o.loss(idx, masks=gm)
            fg = list(torch.autograd.grad(finfo["loss"], [o.params[k] for k in names]))
            if grad_clip:      # the oracle's clip_grad_norm_ on the recomputed gradient
                total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in fg]))
                fg = [g * torch.clamp(grad_clip / (total + 1e-6), max=1.0) for g in fg]
            with torch.no_grad():
                for k in names:
                    o.params[k].copy_(post[k])
            forced_grad = max(grad_excess(gparams[k].grad, g) for k, g in zip(names, fg))
            forced_info = max(abs(float(row[i]) - float(finfo[k])) / max(1.0, abs(float(finfo[k]))) for i, k in enumerate(INFO))
        else:
            forced_grad, forced_info = worst_grad, worst_info
        if fp32_class:
            assert forced_grad <= tol, "step %d: gradient off by %.3e with the ReLU masks forced equal (%d flips)" % (step, forced_grad, flips)
            assert forced_info <= tol, "step %d: loss terms off by %.3e with the ReLU masks forced equal" % (step, forced_info)
            assert worst_grad <= tol + FLIP_GRAD * flips, \
                "step %d: gradient off by %.3e with %d flipped units (allowance %.1e each)" % (step, worst_grad, flips, FLIP_GRAD)
        else:
            # reduced precision perturbs every pre-activation by ~2^-9 relative, so a FRACTION of the units flips, and a
            # gradient that is an incoherent sum over the rows (PPO's surrogate at ratio ~ 1: advantage x noise) moves by
            # ~sqrt(fraction) norm-wise however many rows there are.  With the masks forced equal every tensor must meet the
            # north star's reduced-precision bar; as it is, the error must be explained by the flipped fraction.
            assert forced_grad <= reduced_grad_tol, "step %d: gradient off by %.3e with the ReLU masks forced equal" % (step, forced_grad)
            assert worst_grad <= reduced_grad_tol + 4.0 * np.sqrt(flips / units), (step, worst_grad, flips, units)
        report.append((step, worst_info, worst_grad, flips, forced_grad))
        assert worst_info <= (1e-3 if fp32_class else max(2e-2, 2.5 * tol)), "step %d: loss terms off by %.3e" % (step, worst_info)

    oinfo = oracle.update_model(on_step=on_step, grad_hook=grad_hook, max_steps=steps_synced)
    clean = [r for r in report if r[1] <= tol and r[2] <= tol]
    print("optimizer steps: %d, within tol on every loss term and gradient as they are: %d; ReLU flips per step %s; worst "
          "gradient error: as is %.2e, masks forced %.2e" % (len(report), len(clean), [r[3] for r in report],
                                                             max(r[2] for r in report), max(r[4] for r in report)))
    if fp32_class:
        assert float(np.median([r[1] for r in report])) <= tol, report
    else:
        # reduced-precision modes (bf16 / single-pass TF32): the loss terms within tol on the median step and 2e-2 (the
        # north star's reduced-precision bar) on every step; the gradients are asserted per step above
        assert float(np.median([r[1] for r in report])) <= tol and all(r[1] <= max(2e-2, 2.5 * tol) for r in report), report
    # ---- normalizers
    if steps_synced is None:
        oracle.update_normalizers()
        agent._update_normalizers()
        assert rel_err(agent._obs_norm._mean, oracle.obs_mean) <= FP32_TOL
        assert rel_err(agent._obs_norm._std, oracle.obs_std) <= FP32_TOL
        assert rel_err(agent._disc_obs_norm._mean_abs, oracle.diff_mean_abs) <= FP32_TOL
        assert int(agent._obs_norm._count.item()) == oracle.obs_count
    return oracle, agent, oinfo


def test_iteration_parity_walk_n12_and_golden():
    """The golden case of the executed reference (tests/golden/walk_n12.npz), full 40 optimizer steps."""
    _iteration_parity(12, None, gold_case="walk_n12")


def test_iteration_parity_three_clips_golden():
    """Multi-clip library with the Q2 start-index quirk, WRAP and CLAMP clips, SUCC terminations."""
    _iteration_parity(10, THREE_CLIPS, gold_case="three_clips_n10", steps_synced=8)


def test_iteration_parity_full_library_42_clips():
    """BASELINE configs[2]: the whole assets/motions library (42 clips, 261 MB step table -- not L2-resident), clip ids up
    to 41 through the Q2 start-index quirk, the global row clamp at the end of the table and CLAMP clips that end inside
    the rollout: one full iteration against the oracle on the same library."""
    _iteration_parity(24, ALL_CLIPS, steps_synced=8)


def test_iteration_parity_config0_n64():
    """BASELINE configs[0] shape: 64 envs, one rollout + one ADD/PPO update."""
    _iteration_parity(64, None, steps_synced=16)


def test_iteration_parity_tensor_core_tf32x3_n64():
    """The tensor-core fp32-parity mode (tcgen05 kind::tf32, 3-pass hi/lo split) must meet the same 1e-5 bar."""
    _iteration_parity(64, None, steps_synced=12, precision="tf32x3")


def test_iteration_parity_tensor_core_f16x3_n64():
    """The fp16-pipe fp32-parity mode (tcgen05 kind::f16 on hi/lo fp16 planes scaled by the tensor's max|x|, twins
    cached across the dense layers that share an operand) must meet the same 1e-5 bar."""
    _iteration_parity(64, None, steps_synced=12, precision="f16x3")


def test_iteration_parity_tensor_core_tf32_n64():
    """Single-pass TF32 -- what the reference itself runs on a GPU (`allow_tf32 = True`, main.py:17-18).  10-bit
    operand mantissas: judged against the north star's reduced-precision bar (2e-2) on the loss terms; median step within
    1e-2.  Not a parity mode of this library (legacy build only): at 256-row minibatches the worst gradient tensor with
    the masks forced equal measures 0.20 at step 2 (a nearly cancelling sum, 10-bit operands), hence the 0.25 bound."""
    _iteration_parity(64, None, steps_synced=8, precision="tf32", tol=1e-2, reduced_grad_tol=0.25)


def test_iteration_parity_tensor_core_bf16_n64():
    """BASELINE config 4 arithmetic: bf16 operands (twins written by the producing kernels), fp32 accumulate, fp32 master
    weights / AdamW.  North-star bar for bf16 MLPs: 2e-2 -- met by every loss term on every step.  Gradients at this
    size (256-row minibatches, half the samples at the PPO clip boundary by the second step): 0.1 with the ReLU masks
    forced equal -- the clip decision (ratio inside / outside 1 +- 0.2) is a second discontinuity that 256 rows do not
    average; the 2e-2 gradient bar is asserted at the full 16384-row minibatch (test_full_size_iteration_parity_bf16)."""
    _iteration_parity(64, None, steps_synced=8, precision="bf16", tol=2e-2, reduced_grad_tol=0.1)


def test_full_size_iteration_parity_f16x3_4096_envs():
    """BASELINE configs[1] at full size through the kernels the benchmark times: 4096 envs, 32-step rollout through the
    lean fused step kernel, build_train_data over 131072 rows, and two optimizer steps at the 16384-row minibatch through
    the PERSISTENT f16x3 dense-layer kernel (gemm_tc_h3p_kernel) -- every buffer, loss term and gradient tensor against the
    oracle, ReLU flips counted and forced (see _iteration_parity)."""
    from add_gym_b200 import _lib
    _iteration_parity(4096, None, steps_synced=2, precision="f16x3")
    assert _lib.lib().addk_debug_last_gemm_kernel() in (30, 31, 32)


def test_full_size_iteration_parity_bf16_4096_envs():
    """BASELINE configs[3] arithmetic at the full minibatch (16384 rows): bf16 operands, fp32 accumulate / master weights.
    North star: <= 2e-2 on the per-iteration loss AND gradients.  Measured here: with the ReLU masks forced equal every
    gradient tensor is inside 2e-2; as they are, the critic / discriminator gradients (coherent sums over the rows) are at
    3-5e-3 and the actor's (an incoherent sum: advantage x exploration noise) at sqrt(flipped fraction) -- flips do NOT
    average out with more rows for such a sum, they are counted and bounded instead (see _iteration_parity)."""
    _iteration_parity(4096, None, steps_synced=2, precision="bf16", tol=2e-2)


def test_full_size_iteration_parity_bf16_8192_envs():
    """BASELINE configs[3] at its own size: 8192 envs, bf16 dense layers, one optimizer step at the 32768-row minibatch
    (the rollout and build_train_data over 262144 rows included), same bars as the 4096-env case."""
    _iteration_parity(8192, None, steps_synced=1, precision="bf16", tol=2e-2)


def test_iteration_parity_local_obs_golden():
    """Golden case 3 of the executed reference: local-frame observations with velocity and phase features
    (global_obs=False, enable_vel_obs=True, enable_phase_obs=True) on the three-clip library."""
    _iteration_parity(9, THREE_CLIPS, gold_case="local_vel_phase_n9", steps_synced=4,
                      task_overrides={"global_obs": False, "enable_vel_obs": True, "enable_phase_obs": True})


def test_iteration_parity_with_gradient_clipping():
    """optimizer.grad_clip > 0 (mp_optimizer.py:19-20,46-47; off in the reference's shipped config, quirk Q4): the global
    norm clip runs natively before AdamW.  The oracle's gradients are compared AFTER its clip_grad_norm_, the CUDA path's
    after addk_clip_grad_norm; the clip must have been active (coefficient < 1)."""
    oracle, agent, _ = _iteration_parity(16, None, steps_synced=6, grad_clip=0.05)
    coef = agent._ws["info"][:6, 14].cpu()
    assert bool((coef > 0).all()) and bool((coef < 1).all()), coef


def test_iteration_parity_seven_clips_golden():
    """Golden case 4 of the executed reference: a seven-clip library (clip ids 0..6) -- the Q2 start-index quirk lands
    every clip >= 1 in rows of EARLIER clips, pinned here beyond clip 2; short CLAMP clips end inside the rollout."""
    _iteration_parity(14, SEVEN_CLIPS, gold_case="seven_clips_n14", steps_synced=4)


def test_free_running_iteration_matches_oracle():
    """No parameter re-synchronisation: the CUDA agent runs the reference's whole `_train_iter` on its own and the
    per-iteration diagnostics must agree.  Tolerance 1e-3 relative: without re-synchronisation the ReLU boundary
    flips and AdamW's lr*sign(g) amplification (see _iteration_parity) accumulate over the 40 steps."""
    oracle, agent, rec = _pair(16, None)
    _start(oracle, agent)
    oi = oracle.train_iter()
    gi = agent._train_iter()
    _check_buffers(agent, oracle.buf)
    for k in ("loss", "critic_loss", "actor_loss", "disc_loss", "disc_grad_penalty", "imp_ratio", "adv_mean", "adv_std",
              "disc_reward_mean", "disc_reward_std"):
        assert abs(float(gi[k]) - float(oi[k])) <= 1e-3 * max(1.0, abs(float(oi[k]))), (k, float(gi[k]), float(oi[k]))
    # minibatch permutations were replayed from the oracle's randperm draws: same consumption count
    assert agent.rng.i_perm == len(rec.perms)
    # second iteration keeps tracking (normalizers updated, sampler errors updated, permutation wrap)
    oi = oracle.train_iter()
    gi = agent._train_iter()
    _check_buffers(agent, oracle.buf, keys_float=["obs", "next_obs", "disc_obs", "disc_obs_demo", "motion_times"])
    assert abs(float(gi["loss"]) - float(oi["loss"])) <= 5e-3 * max(1.0, abs(float(oi["loss"])))


# ---------------------------------------------------------------------------------------------------------
# plugin-by-plugin API (how the unmodified reference agent would drive the drop-in classes)
# ---------------------------------------------------------------------------------------------------------
def test_plugin_api_step_by_step():
    from parity_helpers import rel_err
    oracle, agent, rec = _pair(8, None)
    _start(oracle, agent)
    obs_p, rew_p, done_p = agent._add_obs, agent._add_reward, agent._add_done
    for _ in range(5):
        a = torch.zeros(8, 29)
        o_obs, o_r, o_done = oracle.step_env(a)
        agent._env.step(a.cuda())                       # apply_action + scene.step + time_buf += dt (env.py:150-155)
        obs_p.update_motion()
        obs = obs_p.compute_obs()
        r = rew_p.compute_reward()
        d = done_p.compute_done()
        assert rel_err(obs, o_obs) <= FP32_TOL
        assert rel_err(r, o_r) <= FP32_TOL
        assert torch.equal(d.cpu(), o_done)
        assert rel_err(obs_p.info["disc_obs"], oracle.disc_obs) <= FP32_TOL
        assert rel_err(obs_p.info["disc_obs_demo"], oracle.disc_obs_demo) <= FP32_TOL
        assert rel_err(obs_p.ref_dof_pos, oracle.ref[4]) <= FP32_TOL
    # model API: eval_actor / eval_critic / eval_disc against the oracle's forward
    x = torch.randn(33, agent._model.obs_dim)
    dx = torch.randn(17, agent._model.disc_dim)
    with torch.no_grad():
        assert rel_err(agent._model.eval_actor(x.cuda()).mode, oracle.actor_mean(x)) <= FP32_TOL
        assert rel_err(agent._model.eval_critic(x.cuda()), oracle.critic(x)) <= FP32_TOL
        assert rel_err(agent._model.eval_disc(dx.cuda()), oracle.disc(dx)) <= FP32_TOL


def test_contact_list_whose_width_changes_every_step():
    """Real backends hand out a contact list whose width changes from step to step (MuJoCo-Warp: [N, 0] while nothing
    touches -- also at construction time --, Genesis: padded to the per-step maximum; robot.py:221-231,
    mjwarp_engine.py:896-986).  DynamicContactEngine cycles the width through 0, 3, 40, 1, 33, 0, 7 with the contact in a
    different column every step: the FAIL terminations must follow the oracle bit for bit, including the steps with more
    than 32 slots and with none."""
    oracle, agent, rec = _pair(64, None, fall_prob=0.08, engine="add_gym_b200.engine.DynamicContactEngine")
    assert not agent._graphs_ok()
    _start(oracle, agent)
    oracle.rollout()
    agent._rollout_train(agent._steps_per_iter)
    _check_buffers(agent, oracle.buf, keys_float=["obs", "next_obs", "reward", "disc_obs"], keys_exact=["done", "motion_ids"])
    done = oracle.buf["done"]
    assert int((done == 1).sum()) >= 20, "the case must exercise contact terminations"
    per_step = (done == 1).sum(dim=1)
    # widths cycle with scene.t (1-based at the first step): steps whose width is 0 can only fail through the pose test
    assert int(per_step[torch.tensor([i for i in range(32) if (i + 1) % 7 in (2, 4)])].sum()) > 0, "wide (40 / 33 slot) steps fired"


def test_no_cpu_path():
    """The product refuses host tensors instead of falling back."""
    from add_gym_b200 import _lib
    with pytest.raises(_lib.AddkError):
        _lib.ptr(torch.zeros(4))


def test_checkpoint_roundtrip_and_reference_key_names(tmp_path):
    from add_gym_b200.add_agent import ADDAgent
    cfg = b200_config.default_config(num_envs=4)
    a = ADDAgent(cfg, device="cuda:0")
    keys = set(a.state_dict().keys())
    for k in ("_obs_norm._count", "_obs_norm._mean", "_obs_norm._std", "_a_norm._mean", "_disc_obs_norm._mean_abs",
              "_model._actor_layers.0.weight", "_model._actor_layers.4.bias", "_model._action_dist._mean_net.weight",
              "_model._action_dist._logstd_net", "_model._critic_out.weight", "_model._disc_layers.2.weight",
              "_model._disc_logits.bias"):
        assert k in keys, k
    p = str(tmp_path / "model.pt")
    a.save(p)
    b = ADDAgent(cfg, device="cuda:0")
    b.load(p)
    assert torch.equal(a._model.flat, b._model.flat)


def test_loads_a_checkpoint_written_by_the_reference(tmp_path):
    """f4: a checkpoint written by the executed reference's own `save()` (tests/golden/make_ref_checkpoint.py: real
    ADDAgent, one iteration, per-tensor patterns in the weights and the AdamW state) loads into the drop-in agent --
    every trainable tensor lands at its offset of the flat vector, the AdamW moments and step count follow, normalizer
    statistics and counters are restored -- and training continues from optimizer step 41."""
    import sys
    import zipfile
    from add_gym_b200.add_agent import ADDAgent
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    from make_ref_checkpoint import pattern
    with zipfile.ZipFile(os.path.join(golden, "ref_checkpoint.zip")) as z:
        z.extract("model.pt", str(tmp_path))
    path = str(tmp_path / "model.pt")
    ck = torch.load(path, map_location="cpu")
    cfg = b200_config.default_config(num_envs=4)
    cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.01)
    a = ADDAgent(cfg, device="cuda:0")
    a.load(path)
    assert a._iter == 7 and a._sample_count == 7 * 128 and a._is_restored
    names, tensors = a._model.trainable()
    opt = a._optimizer
    assert opt.steps == 40 and opt.lr == 1e-4 and tuple(opt.betas) == (0.9, 0.999)
    for i, (n, t) in enumerate(zip(names, tensors)):
        o = a._model.offsets["o_" + n]
        assert torch.equal(a._model.flat[o:o + t.numel()].cpu(), pattern(t.numel(), 0, i)), n
        assert torch.equal(t.detach().flatten().cpu(), pattern(t.numel(), 0, i)), n          # the module view is the flat vector
        assert torch.equal(opt.exp_avg[o:o + t.numel()].cpu(), pattern(t.numel(), 1, i)), n
        assert torch.equal(opt.exp_avg_sq[o:o + t.numel()].cpu(), pattern(t.numel(), 2, i)), n
    sd = a.state_dict()
    for k in ("_obs_norm._count", "_obs_norm._mean", "_obs_norm._std", "_a_norm._mean", "_a_norm._std",
              "_disc_obs_norm._count", "_disc_obs_norm._mean_abs", "_model._action_dist._logstd_net"):
        assert torch.equal(sd[k].cpu(), ck["model"][k]), k
    # the same file through the reference's DDP key prefix (base_agent.py:169-186)
    ck2 = dict(ck)
    ck2["model"] = {k.replace("_model.", "_model.module.", 1): v for k, v in ck["model"].items()}
    path2 = str(tmp_path / "model_ddp.pt")
    torch.save(ck2, path2)
    b = ADDAgent(cfg, device="cuda:0")
    b.load(path2)
    assert torch.equal(a._model.flat, b._model.flat)
    # training continues: one iteration = 40 optimizer steps on top of the restored count, everything finite
    a._curr_obs, a._curr_info = a._reset_envs()
    a._exp_buffer.clear()
    a._reset_tracker()
    info = a._train_iter()
    assert a._optimizer.steps == 80
    assert all(math.isfinite(float(v)) for v in info.values())
    assert bool(torch.isfinite(a._model.flat).all()) and not torch.equal(a._model.flat, b._model.flat)


def test_cuda_graph_rollout_is_bit_identical_to_eager():
    """The rollout step replayed from CUDA graphs (two segments around the physics step) must write exactly what the
    eager launch sequence writes: same kernels, same order, fixed addresses.  Deterministic draws make the two runs
    comparable (torch's generator advances differently under capture)."""
    from add_gym_b200.add_agent import ADDAgent, DeviceRandom

    class FixedRandom(DeviceRandom):
        def __init__(self, device):
            super().__init__(device)
            self.k = 0

        def action_noise(self, n, dim):
            return torch.full((n, dim), 0.25, device=self.device)

        def reset_uniforms(self, n):
            return torch.full((n, 3), 0.37, device=self.device)

    def run(graphs):
        cfg = b200_config.default_config(num_envs=96)
        cfg["agent"]["cuda_graphs"] = graphs
        cfg["engine"].update(seed=77, noise_device="device", fall_prob=0.02)
        torch.manual_seed(0)
        a = ADDAgent(cfg, device="cuda:0")
        a.rng = FixedRandom("cuda:0")
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        a._reset_tracker()
        snaps = []
        for _ in range(3):                       # first rollout eager in both; then graphs are captured and replayed
            a._rollout_train(a._steps_per_iter)
            snaps.append({k: a._exp_buffer.get_data(k).clone() for k in FLOAT_KEYS + EXACT_KEYS})
        torch.cuda.synchronize()
        return a, snaps

    eager, s0 = run(False)
    graphed, s1 = run(True)
    assert len(graphed._graphs_pre) == graphed._steps_per_iter and len(graphed._graphs_post) == 3 * graphed._steps_per_iter
    assert len(eager._graphs_pre) == 0
    for it, (x, y) in enumerate(zip(s0, s1)):
        for k in x:
            assert torch.equal(x[k], y[k]), (it, k)
    assert torch.equal(eager._env.time_buf, graphed._env.time_buf)
    assert torch.equal(eager._core.hist, graphed._core.hist) and eager._core.hist_head == graphed._core.hist_head
    assert int(eager._core.tracker_count.item()) == int(graphed._core.tracker_count.item())


@pytest.mark.parametrize("precision", ["tf32x3", "bf16", "f16x3"])
def test_three_stream_update_is_bit_identical_to_one_stream(precision):
    """The actor / critic / discriminator chains of an optimizer step on three streams (separate workspaces, tail waves
    overlapping) must produce exactly the parameters the back-to-back single-stream order produces: every kernel is
    deterministic and the chains write disjoint slab segments and statistics slots."""
    from add_gym_b200.add_agent import ADDAgent
    _need_legacy(precision)

    def run(streams):
        cfg = b200_config.default_config(num_envs=96, mlp_precision=precision)
        cfg["agent"]["update_streams"] = streams
        cfg["engine"].update(seed=5, noise_device="device", fall_prob=0.02)
        torch.manual_seed(1)
        a = ADDAgent(cfg, device="cuda:0")
        a._curr_obs, a._curr_info = a._reset_envs()
        infos = []
        for _ in range(2):
            a._exp_buffer.clear()
            a._rollout_train(a._steps_per_iter)
            a._build_train_data()
            infos.append(dict(a._update_model()))
        torch.cuda.synchronize()
        return a, infos

    one, i1 = run(1)
    three, i3 = run(3)
    assert one._ctx.ints["n_streams"] == 1 and three._ctx.ints["n_streams"] == 3
    assert torch.equal(one._model.flat, three._model.flat)
    assert torch.equal(one._optimizer.exp_avg_sq, three._optimizer.exp_avg_sq)
    for x, y in zip(i1, i3):
        for k in x:
            assert float(x[k]) == float(y[k]), k
