"""GPU: the smaller entry points of the C-ABI one by one, edge cases, and size-independent properties at BASELINE's full
sizes (4096 / 32768 envs), where running the oracle for a whole iteration would take minutes.

Bit-exact where the result is an index / flag / copy; 1e-5 norm-wise for floating point (the tolerance is at each assert).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from add_gym_b200 import config as b200_config

pytestmark = pytest.mark.gpu
TOL = 1e-5
THREE_CLIPS = os.path.join(b200_config.ASSET_DIR, "three_clips.yaml")
SEVEN_CLIPS = os.path.join(b200_config.ASSET_DIR, "seven_clips.yaml")


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu().flatten(), torch.as_tensor(b).double().cpu().flatten()
    d = float(torch.linalg.norm(b))
    return float(torch.linalg.norm(a - b)) / (d if d > 0 else 1.0)


def _L():
    from add_gym_b200 import _lib
    return _lib, _lib.lib()


# ---------------------------------------------------------------------------------------------------------------
# returns / advantages / rewards / statistics (csrc/gae.cu)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,N", [(32, 4096), (32, 1), (1, 7), (5, 33)])
def test_td_lambda_matches_oracle_and_is_linear(T, N):
    from oracle import add_oracle
    _lib, L = _L()
    g = torch.Generator().manual_seed(T * 1000 + N)
    r, nv, v = (torch.randn(T, N, generator=g) for _ in range(3))
    done = torch.randint(0, 4, (T, N), generator=g).int() * (torch.rand(T, N, generator=g) < 0.15).int()

    def run(r_, nv_, v_):
        tar, adv = torch.empty(T, N, device="cuda"), torch.empty(T, N, device="cuda")
        rd, nd, vd, dd = r_.cuda(), nv_.cuda(), v_.cuda(), done.cuda()      # keep the device copies alive across the call
        _lib.check(L.addk_td_lambda(_lib.stream(), _lib.ptr(rd), _lib.ptr(nd), _lib.ptr(vd), _lib.ptr(dd), C.c_int(T),
                                    C.c_int(N), C.c_float(0.99), C.c_float(0.95), C.c_float(0.0), C.c_float(0.0),
                                    _lib.ptr(tar), _lib.ptr(adv)), "addk_td_lambda")
        torch.cuda.synchronize()
        return tar.cpu(), adv.cpu()
    tar, adv = run(r, nv, v)
    nvm = nv.clone()
    nvm[(done == 1) | (done == 2)] = 0.0                                  # FAIL / SUCC bootstrap value (ppo_agent.py:126-131)
    ref = add_oracle.td_lambda_return(r, nvm, done, 0.99, 0.95)
    assert torch.equal(tar, ref), "same fp32 op order as the reference loop: bit-exact"
    assert torch.equal(adv, ref - v)
    # linearity in (reward, bootstrap values) -- holds at any size
    r2, nv2 = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
    t12, _ = run(r + 2 * r2, nv + 2 * nv2, v)
    t2, _ = run(r2, nv2, v)
    assert _rel(t12, tar + 2 * t2) <= TOL


@pytest.mark.parametrize("n", [131072, 1, 37])
def test_adv_normalize_and_disc_reward(n):
    _lib, L = _L()
    g = torch.Generator().manual_seed(n)
    adv = torch.randn(n, generator=g) * 3 + 1
    mask = (torch.rand(n, generator=g) < 0.9).float() if n > 1 else torch.ones(1)
    a, md = adv.cuda(), mask.cuda()
    work, stats = torch.zeros(3, dtype=torch.float64, device="cuda"), torch.zeros(2, device="cuda")
    _lib.check(L.addk_adv_normalize(_lib.stream(), _lib.ptr(a), _lib.ptr(md), C.c_int(n), C.c_float(4.0), _lib.ptr(work),
                                    _lib.ptr(stats)), "addk_adv_normalize")
    sel = adv[mask == 1.0]
    if sel.numel() > 1:
        sd, mu = torch.std_mean(sel)                                           # unbiased, as ppo_agent.py:147-153
        ref = torch.clamp((adv - mu) / torch.clamp_min(sd, 1e-5), -4.0, 4.0)
        assert _rel(a, ref) <= TOL and abs(float(stats[0]) - float(mu)) <= TOL * max(1, abs(float(mu)))
        assert abs(float(stats[1]) - float(sd)) <= TOL * float(sd)
    logits, task_r = torch.randn(n, generator=g) * 4, torch.rand(n, generator=g)
    rew, lg = task_r.cuda(), logits.cuda()
    _lib.check(L.addk_disc_reward(_lib.stream(), _lib.ptr(lg), _lib.ptr(rew), C.c_int(n), C.c_float(2.0), C.c_float(0.25),
                                  C.c_float(1.0), _lib.ptr(work), _lib.ptr(stats)), "addk_disc_reward")
    prob = 1 / (1 + torch.exp(-logits))
    dr = -torch.log(torch.maximum(1 - prob, torch.tensor(0.0001))) * 2.0       # amp_agent.py:201-205
    assert _rel(rew, 0.25 * task_r + dr) <= TOL
    assert abs(float(stats[0]) - float(dr.mean())) <= TOL * max(1.0, abs(float(dr.mean())))


def test_adamw_matches_torch_at_full_parameter_count():
    _lib, L = _L()
    n = 4349983 + 25
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g) * 0.05
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([q], 1e-4, weight_decay=0.0)
    p, m, v = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * 10.0 ** (-step)
        q.grad = grad.clone()
        opt.step()
        gd = grad.cuda()
        _lib.check(L.addk_adamw(_lib.stream(), _lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v), C.c_longlong(n),
                                C.c_int(step), C.c_double(1e-4), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-8),
                                C.c_double(0.0), C.c_double(1.0)), "addk_adamw")
        assert _rel(p, q.data) <= 1e-7 and float((p.cpu() - q.data).abs().max()) <= 2e-8, "single-tensor AdamW op order (step %d)" % step


@pytest.mark.parametrize("n,off", [(4349983, 0), (1003, 0), (1003, 1), (3, 0)])
def test_adamw_vector_tail_and_unaligned_vectors(n, off):
    """The 128-bit kernel handles n % 4 != 0 through one tail thread; vectors that are not 16-byte aligned take the scalar
    kernel.  Both must give the bits of torch's single-tensor AdamW."""
    _lib, L = _L()
    g = torch.Generator().manual_seed(n + off)
    p0, grad = torch.randn(n, generator=g) * 0.05, torch.randn(n, generator=g) * 0.01
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([q], 1e-4, weight_decay=0.0)
    q.grad = grad.clone()
    opt.step()
    bufs = [torch.zeros(n + off, device="cuda") for _ in range(4)]
    p, gd, m, v = (b[off:] for b in bufs)
    p.copy_(p0); gd.copy_(grad)
    _lib.check(L.addk_adamw(_lib.stream(), _lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v), C.c_longlong(n), C.c_int(1),
                            C.c_double(1e-4), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-8), C.c_double(0.0),
                            C.c_double(1.0)), "addk_adamw")
    torch.cuda.synchronize()
    assert float((p.cpu() - q.data).abs().max()) <= 2e-8
    assert torch.equal(m.cpu(), (0.1 * grad)) or _rel(m, 0.1 * grad) <= 1e-7
    assert all(float(b[:off].abs().sum()) == 0.0 for b in bufs), "nothing written before the vectors"


def test_column_stats_and_normalizer_updates():
    from add_gym_b200.normalizer import DiffNormalizer, Normalizer
    g = torch.Generator().manual_seed(3)
    x = torch.randn(32 * 4096, 264, generator=g) * 2 + 0.5
    nz = Normalizer([264], device="cuda")
    xa, xb = x[:70000].cuda(), x[70000:].cuda()
    nz.record(xa); nz.record(xb)
    nz.update()
    torch.cuda.synchronize()
    assert int(nz._count.item()) == x.shape[0]
    assert _rel(nz._mean, x.double().mean(0)) <= TOL
    assert _rel(nz._std, x.double().std(0, unbiased=False)) <= TOL
    x2 = torch.randn(1000, 264, generator=g)
    x2d = x2.cuda()
    nz.record(x2d); nz.update()                                              # weighted merge with the running moments
    torch.cuda.synchronize()
    allx = torch.cat([x, x2]).double()
    assert _rel(nz._mean, allx.mean(0)) <= TOL and _rel(nz._std, allx.std(0, unbiased=False)) <= 5 * TOL
    a, b = torch.randn(5000, 114, generator=g), torch.randn(5000, 114, generator=g)
    dn = DiffNormalizer([114], device="cuda")
    ad, bd = a.cuda(), b.cuda()
    dn.record_pair(ad, bd); dn.update()
    torch.cuda.synchronize()
    assert _rel(dn._mean_abs, (a - b).abs().double().mean(0)) <= TOL


# ---------------------------------------------------------------------------------------------------------------
# reset sampling (csrc/step.cu): the host-sync-free masked variant draws by inverse CDF from three uniforms per env
# ---------------------------------------------------------------------------------------------------------------
def test_masked_reset_sampling_distribution_and_time_grid():
    from add_gym_b200.add_motion import ADDMotion
    from add_gym_b200.env import ImitationEnvironment
    cfg = b200_config.default_config(num_envs=4, motion_file=THREE_CLIPS)
    env = ImitationEnvironment(cfg, "cuda:0")
    motion = ADDMotion(cfg["task"], env, "cuda:0")
    n = 300000
    samp = motion.sampler
    samp.errors.copy_(torch.rand_like(samp.errors) * 3)
    ids, times = torch.full((n,), -1, dtype=torch.long, device="cuda"), torch.full((n,), -1.0, device="cuda")
    done = (torch.arange(n, device="cuda") % 3 != 0).int()                   # every third env keeps its clip
    u = torch.rand(n, 3, device="cuda")
    motion.sample_time_masked(done, ids, times, uniforms=u)
    keep = done == 0
    assert bool((ids[keep] == -1).all()) and bool((times[keep] == -1.0).all()), "envs that are not done are untouched"
    sel = ~keep
    w = motion.motion_lib.get_motion_weights().cpu()
    freq = torch.bincount(ids[sel].cpu(), minlength=3).double() / int(sel.sum())
    assert float((freq - w.double()).abs().max()) < 5e-3, "clip frequencies follow the normalised weights"
    # start times: on the ctrl_dt grid (reference: (t // dt) * dt), >= (num_disc_obs_steps - 1) * dt, inside the clip
    t = times[sel].cpu()
    k = torch.round(t / 0.01)
    assert float((t - k * 0.01).abs().max()) < 1e-5
    assert float(t.min()) >= 0.02 - 1e-7
    lengths = motion.motion_lib.get_motion_lengths().cpu()
    assert bool((t <= lengths[ids[sel].cpu()] + 1e-4).all())
    # segment frequencies of clip 0 follow softmax(err / (max err over the drawn clips + 1e-6))  (sampler.py:57-73)
    c0 = (ids == 0) & sel
    seg = torch.clamp((times[c0] / samp.segment_sizes[0]).long(), 0, samp.num_segments - 1).cpu()
    temp = float(samp.errors.max()) + 1e-6
    probs = torch.softmax(samp.errors[0].cpu().double() / temp, dim=-1)
    sfreq = torch.bincount(seg, minlength=samp.num_segments).double() / int(c0.sum())
    assert float((sfreq - probs).abs().max()) < 6e-3
    # deterministic in the uniforms
    ids2, times2 = torch.zeros_like(ids), torch.zeros_like(times)
    motion.sample_time_masked(done, ids2, times2, uniforms=u)
    assert torch.equal(ids2[sel], ids[sel]) and torch.equal(times2[sel], times[sel])


def test_start_time_arithmetic_is_bit_exact_on_the_oracle_draws():
    """the start-time arithmetic of AdaptiveSegmentSampler.sample_start_frame / ADDMotion.sample_time
    (sampler.py:75-92, add_motion.py:53-61) -- seg * size + U * size, (t // dt) * dt (c10 floor division on fp32),
    clamp(min = 0.02) -- fed with the ORACLE's recorded (clip, segment, U) draws must give the oracle's start times bit
    for bit.  The oracle draws come from the reference's own torch calls (multinomial / rand on the CPU generator)."""
    from add_gym_b200.add_motion import ADDMotion
    from add_gym_b200.env import ImitationEnvironment
    from oracle import harness
    import parity_helpers as helpers
    for motion in (None, THREE_CLIPS, SEVEN_CLIPS):
        cfg = b200_config.default_config(num_envs=4, motion_file=motion)
        env = ImitationEnvironment(cfg, "cuda:0")
        gmotion = ADDMotion(cfg["task"], env, "cuda:0")
        rec = helpers.RecordRandom()
        oracle = harness.make_oracle_agent(4, seed=3, engine_seed=1234, cfg=b200_config.default_config(num_envs=4, motion_file=motion), rng=rec)
        assert torch.equal(gmotion.sampler.segment_sizes.cpu(), oracle.seg_sizes), "segment sizes: fp32 len / 20"
        oracle.errors = torch.rand_like(oracle.errors) * 3.0           # non-uniform segment probabilities
        torch.manual_seed(11)
        n0 = len(rec.clip_draws)
        ids, times = oracle.sample_time(200000)
        clip, seg, u = rec.clip_draws[n0], rec.segment_draws[n0], rec.uniform_draws[n0]
        assert torch.equal(clip, ids)
        got = gmotion.sampler.start_times_from_draws(clip.cuda(), seg.cuda(), u.cuda())
        assert torch.equal(got.cpu(), times), "start times must be bit-exact on identical draws"
        # edge draws: U = 0, U just below 1, first / last segment, the clamp at (num_disc_obs_steps - 1) * dt
        C_ = oracle.seg_sizes.shape[0]
        e_clip = torch.arange(C_).repeat_interleave(6)
        e_seg = torch.tensor([0, 0, 19, 19, 7, 0]).repeat(C_)
        e_u = torch.tensor([0.0, 1.0 - 2.0 ** -24, 0.0, 1.0 - 2.0 ** -24, 0.5, 1e-4]).repeat(C_)
        sz = oracle.seg_sizes[e_clip]
        t = e_seg * sz + e_u * sz
        ref = torch.clamp((t // oracle.dt) * oracle.dt, min=oracle.min_start)
        got = gmotion.sampler.start_times_from_draws(e_clip.cuda(), e_seg.cuda(), e_u.cuda())
        assert torch.equal(got.cpu(), ref)


def test_sampler_update_errors_matches_oracle():
    from add_gym_b200.add_motion import AdaptiveSegmentSampler
    g = torch.Generator().manual_seed(5)
    lengths = torch.tensor([124.17, 23.3, 14.97])
    s = AdaptiveSegmentSampler(lengths.cuda(), 0.01, num_segments=20, min_start_time=0.02)
    n = 50000
    ids = torch.randint(0, 3, (n,), generator=g)
    times = torch.rand(n, generator=g) * lengths[ids]
    a, b = torch.randn(n, 114, generator=g), torch.randn(n, 114, generator=g)
    idd, td, ad, bd = ids.cuda(), times.cuda(), a.cuda(), b.cuda()
    s.update_errors(idd, td, ad, bd)
    torch.cuda.synchronize()
    err = torch.sum(torch.square(a - b), dim=-1)
    sz = torch.clamp(lengths / 20, min=1e-6)[ids]
    seg = torch.clamp((times / sz).long(), 0, 19)
    flat = ids * 20 + seg
    mean = torch.zeros(60).scatter_reduce(0, flat, err, reduce="mean", include_self=False)
    hit = torch.zeros(60).scatter_add(0, flat, torch.ones(n)) > 0
    ref = torch.where(hit, 0.9 * torch.ones(60) + 0.1 * mean, torch.ones(60)).view(3, 20)
    assert _rel(s.errors, ref) <= TOL


# ---------------------------------------------------------------------------------------------------------------
# full-size checks (BASELINE configs[1] / configs[4] shapes)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num_envs", [4096, 32768])
def test_full_size_rollout_steps_match_oracle(num_envs):
    """Four env steps at 4096 envs (BASELINE configs[1]) and at 32768 envs (configs[4], the largest configuration) through
    the lean fused kernel + masked reset, against the oracle on the same synthetic physics stream: flags / ids bit-exact,
    rows 1e-5."""
    import parity_helpers as helpers
    import test_gpu_parity as T
    oracle, agent, rec = T._pair(num_envs, None, fall_prob=0.01)
    T._start(oracle, agent)
    oracle.rollout(4)
    agent._rollout_train(4)
    for k in T.EXACT_KEYS:
        assert torch.equal(agent._exp_buffer.get_data(k)[:4].cpu().to(oracle.buf[k].dtype), oracle.buf[k][:4]), k
    for k in T.FLOAT_KEYS:
        e = helpers.rel_err(agent._exp_buffer.get_data(k)[:4], oracle.buf[k][:4])
        assert e <= TOL, (k, e)
    assert torch.equal(agent._env.time_buf.cpu(), oracle.time_buf)
    assert int((oracle.buf["done"][:4] != 0).sum()) > 0, "the case must exercise resets"


@pytest.mark.parametrize("tc_mode", ["f16x3", "tf32x3"])
def test_full_size_update_tensor_core_vs_cuda_core_gradients(tc_mode):
    """A SECOND check next to the oracle comparison at full size (test_gpu_parity.py::test_full_size_iteration_parity_*):
    one optimizer step at the full minibatch (16384 rows) computed twice by independent arithmetic inside the library,
    the tcgen05 fp32-parity tiles (f16x3: fp16 hi/lo planes; tf32x3 in the legacy build) and the exact-fp32 CUDA-core
    kernel.  The loss terms must agree to 2e-5; the median gradient tensor to 5e-5 and the worst one to 5e-3: the two
    runs take different ReLU branches for a few of the 50M pre-activations (1/sqrt(8M active units) = 3.5e-4 per flipped
    unit and tensor -- the per-flip accounting is done against the oracle, where the masks can be forced equal).  The
    same call repeated must reproduce itself bit for bit."""
    from add_gym_b200 import _lib
    from add_gym_b200.add_agent import ADDAgent
    import test_gpu_parity as T
    T._need_legacy(tc_mode)

    def grads(prec):
        cfg = b200_config.default_config(num_envs=4096, mlp_precision=prec)
        cfg["engine"].update(seed=99, noise_device="device", fall_prob=0.002)
        torch.manual_seed(0)
        a = ADDAgent(cfg, device="cuda:0")
        torch.manual_seed(1)
        torch.cuda.manual_seed(1)
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        a._rollout_train(a._steps_per_iter)
        a._build_train_data()
        idx = torch.arange(0, a._mb_rows, device="cuda", dtype=torch.long) * 7 % (a._steps_per_iter * 4096)
        out = []
        for rep in range(2):
            _lib.check(_lib.lib().addk_update_minibatch(_lib.stream(), a._ctx.buf, _lib.ptr(idx), C.c_int(rep), C.c_int(0)), "mb")
            out.append((a._model.flat_grad.clone(), a._ws["info"][rep].clone()))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]), "deterministic"
        names, tensors = a._model.trainable()
        return {n: out[0][0][a._model.offsets["o_" + n]:a._model.offsets["o_" + n] + t.numel()].clone() for n, t in zip(names, tensors)}, out[0][1]

    g1, i1 = grads(tc_mode)
    g0, i0 = grads("fp32")
    for k in range(13):
        assert abs(float(i1[k]) - float(i0[k])) <= 2e-5 * max(1.0, abs(float(i0[k]))), (k, float(i1[k]), float(i0[k]))
    errs = {n: _rel(g1[n], g0[n]) for n in g0}
    print("full-size gradient agreement %s vs fp32:" % tc_mode, {k: "%.1e" % v for k, v in errs.items()})
    assert float(np.median(list(errs.values()))) <= 5e-5
    assert max(errs.values()) <= 5e-3, "beyond a few ReLU boundary flips (1/sqrt(8M active units) = 3.5e-4 each)"


def test_fused_optimizer_tail_is_bit_identical_to_reduce_then_adamw():
    """One GPU runs the tail of an optimizer step as ONE launch (reduce_slabs_adamw_kernel: split-K slab reduction +
    AdamW + diagnostics row); the multi-rank path runs reduce_slabs_kernel (slab sums + diagnostics row) inside
    addk_update_minibatch(do_optim=0) and addk_adamw afterwards.  Same arithmetic element by element: parameters, both
    moment vectors, the summed gradient and the diagnostics row must come out bit-identical over several steps."""
    from add_gym_b200 import _lib
    from add_gym_b200.add_agent import ADDAgent

    def make():
        cfg = b200_config.default_config(num_envs=96)
        cfg["engine"].update(seed=11, noise_device="device", fall_prob=0.02)
        torch.manual_seed(3)
        torch.cuda.manual_seed(3)
        a = ADDAgent(cfg, device="cuda:0")
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        a._rollout_train(a._steps_per_iter)
        a._build_train_data()
        return a

    fused, split = make(), make()
    assert torch.equal(fused._model.flat, split._model.flat)
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(4):
        idx = torch.randint(0, 96 * fused._steps_per_iter, (fused._mb_rows,), device="cuda", generator=g)
        _lib.check(L.addk_update_minibatch(_lib.stream(), fused._ctx.buf, _lib.ptr(idx), C.c_int(step), C.c_int(step + 1)), "fused")
        _lib.check(L.addk_update_minibatch(_lib.stream(), split._ctx.buf, _lib.ptr(idx), C.c_int(step), C.c_int(0)), "split")
        opt = split._optimizer
        _lib.check(L.addk_adamw(_lib.stream(), _lib.ptr(split._model.flat), _lib.ptr(split._model.flat_grad), _lib.ptr(opt.exp_avg),
                                _lib.ptr(opt.exp_avg_sq), C.c_longlong(split._model.num_params), C.c_int(step + 1),
                                C.c_double(opt.lr), C.c_double(opt.betas[0]), C.c_double(opt.betas[1]), C.c_double(opt.eps),
                                C.c_double(opt.weight_decay), C.c_double(1.0)), "addk_adamw")
        assert torch.equal(fused._model.flat_grad, split._model.flat_grad), step
        assert torch.equal(fused._model.flat, split._model.flat), step
        assert torch.equal(fused._optimizer.exp_avg, opt.exp_avg) and torch.equal(fused._optimizer.exp_avg_sq, opt.exp_avg_sq), step
        assert torch.equal(fused._ws["info"][step], split._ws["info"][step]), step
    assert float(fused._model.flat.abs().max()) > 0 and bool(torch.isfinite(fused._model.flat).all())


def test_empty_and_ragged_requests():
    from add_gym_b200.add_motion import ADDMotion
    from add_gym_b200.env import ImitationEnvironment
    _lib, L = _L()
    cfg = b200_config.default_config(num_envs=3)
    env = ImitationEnvironment(cfg, "cuda:0")
    motion = ADDMotion(cfg["task"], env, "cuda:0")
    out = motion.get_motion_step(torch.zeros(0, dtype=torch.long, device="cuda"), torch.zeros(0, device="cuda"))
    assert [o.shape[0] for o in out] == [0] * 6
    one = motion.get_motion_step(torch.zeros(1, dtype=torch.long, device="cuda"), torch.tensor([1.234], device="cuda"))
    assert one[4].shape == (1, 29)
    # invalid arguments are refused with an error code, not a crash
    assert L.addk_td_lambda(_lib.stream(), None, None, None, None, C.c_int(0), C.c_int(0), C.c_float(0.99), C.c_float(0.95),
                            C.c_float(0), C.c_float(0), None, None) != 0
    assert L.addk_adamw(_lib.stream(), None, None, None, None, C.c_longlong(0), C.c_int(1), C.c_double(1e-4), C.c_double(0.9),
                        C.c_double(0.999), C.c_double(1e-8), C.c_double(0.0), C.c_double(1.0)) != 0


# ---------------------------------------------------------------------------------------------------------------
# the exchange step as one kernel over peer memory (csrc/p2p.cu), emulated on ONE device: three "ranks" = three
# concurrent launches on three streams that address each other's buffers directly (no cudaIpc needed in one process)
# ---------------------------------------------------------------------------------------------------------------
def test_p2p_exchange_kernel_three_ranks_on_one_device():
    _lib, L = _L()
    world, n = 3, 4349984                       # the model's padded parameter count
    g = torch.Generator(device="cuda").manual_seed(3)
    grads = [torch.randn(n, device="cuda", generator=g) * 1e-3 for _ in range(world)]
    p0 = torch.randn(n, device="cuda", generator=g) * 0.05
    m0 = torch.randn(n, device="cuda", generator=g) * 1e-4
    v0 = torch.rand(n, device="cuda", generator=g) * 1e-6
    params = [p0.clone() for _ in range(world)]
    ms, vs = [m0.clone() for _ in range(world)], [v0.clone() for _ in range(world)]
    flags = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(world)]
    tickets = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    arr = lambda ts: (C.c_void_p * world)(*[t.data_ptr() for t in ts])
    gp, pp, fp = arr(grads), arr(params), arr(flags)
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for step in (1, 2):                          # two consecutive epochs: the flags are reused
        for r in reversed(range(world)):         # launch order must not matter
            rc = L.addk_p2p_adamw(C.c_void_p(streams[r].cuda_stream), C.c_int(r), C.c_int(world), gp, pp, fp, _lib.ptr(ms[r]),
                                  _lib.ptr(vs[r]), C.c_longlong(n), C.c_int(step), C.c_double(1e-4), C.c_double(0.9),
                                  C.c_double(0.999), C.c_double(1e-8), C.c_double(0.0), _lib.ptr(tickets[r]), C.c_int(32))
            _lib.check(rc, "addk_p2p_adamw")
        torch.cuda.synchronize()
        # reference: the all-reduced gradient (rank-order sum) through the flat AdamW kernel with grad_scale = 1 / world
        gsum = grads[0].clone()
        for r in range(1, world):
            gsum += grads[r]
        _lib.check(L.addk_adamw(_lib.stream(), _lib.ptr(p0), _lib.ptr(gsum), _lib.ptr(m0), _lib.ptr(v0), C.c_longlong(n), C.c_int(step),
                                C.c_double(1e-4), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-8), C.c_double(0.0),
                                C.c_double(1.0 / world)), "addk_adamw")
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(params[r], p0), "every rank holds the same bits as all-reduce + AdamW (step %d, rank %d)" % (step, r)
        per = 4 * (((n + 3) // 4 + world - 1) // world)
        for r in range(world):                   # the moments are sharded: rank r keeps [r * per, (r + 1) * per) current
            b, e = r * per, min(n, (r + 1) * per)
            assert torch.equal(ms[r][b:e], m0[b:e]) and torch.equal(vs[r][b:e], v0[b:e])
        for t in grads:                          # next epoch: new gradients
            t.mul_(-0.5).add_(1e-4)
