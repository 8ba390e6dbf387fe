"""Data-parallel semantics: each rank rolls out its own env slice; the ONE exchange step is the gradient
all-reduce (mean) before AdamW.  The reference's DDP wrapper never fires (quirk Q5), so the parity oracle is R oracle
agents with identical initial weights, per-rank engine seeds and gradients explicitly averaged.

* CPU, world_size 2 over gloo: the oracle-side protocol itself (broadcast, all-reduce mean, identical weights after).
* GPU: the product's non-local update path (addk_update_minibatch without the optimizer, summed flat gradients,
  addk_adamw with grad_scale = 1/world) for two emulated ranks on one device against that oracle.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from add_gym_b200 import config as b200_config


def _small_cfg(n):
    cfg = b200_config.default_config(num_envs=n)
    cfg["agent"]["steps_per_iter"] = 4
    cfg["agent"]["update_epochs"] = 3
    return cfg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import harness
    torch.set_num_threads(2)
    n = 4
    agent = harness.make_oracle_agent(n, seed=rank, engine_seed=1234 + rank, cfg=_small_cfg(n), fall_prob=0.01)
    for k in agent.names:                                   # DDP's constructor broadcast (base_agent.py:47-57)
        dist.broadcast(agent.params[k].data, src=0)
    agent.start()
    agent.rollout()
    agent.build_train_data()

    def grad_hook(grads):                                   # the intended gradient all-reduce (mean)
        for g in grads.values():
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            g /= world
    agent.update_model(grad_hook=grad_hook)
    flat = torch.cat([agent.params[k].detach().flatten() for k in agent.names])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    obs = [torch.zeros_like(agent.buf["obs"]) for _ in range(world)]
    dist.all_gather(obs, agent.buf["obs"])
    if rank == 0:
        torch.save({"same_weights": all(torch.equal(gathered[0], g) for g in gathered),
                    "different_rollouts": not torch.equal(obs[0], obs[1]), "steps": agent.adam_steps},
                   os.path.join(out_dir, "result.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_ranks_gradient_averaging_keeps_weights_identical(tmp_path):
    mp.spawn(_gloo_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(str(tmp_path), "result.pt"))
    assert r["same_weights"] and r["different_rollouts"] and r["steps"] == 3


def test_bench_reference_arm_only_rank0_prints():
    """Under torchrun the reference arm runs on rank 0 alone; other ranks exit 0 without output."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_two_emulated_ranks_match_the_gradient_averaging_oracle():
    import parity_helpers as helpers
    from add_gym_b200 import _lib
    from add_gym_b200.add_agent import ADDAgent
    from oracle import harness
    world, n = 2, 8
    agents, oracles = [], []
    for r in range(world):
        gcfg = _small_cfg(n)
        gcfg["engine"].update(seed=1234 + r, noise_device="cpu", fall_prob=0.01)
        torch.manual_seed(0)
        a = ADDAgent(gcfg, device="cuda:0")
        olib = harness.make_oracle_lib(_small_cfg(n), jrot_override=[j.cpu() for j in a._add_motion.motion_lib._frame_joint_rot])
        rec = helpers.RecordRandom()
        o = harness.make_oracle_agent(n, seed=0, engine_seed=1234 + r, cfg=_small_cfg(n), rng=rec, fall_prob=0.01, lib=olib)
        if r > 0:
            for k in o.names:
                o.params[k].data.copy_(oracles[0].params[k].data)
        helpers.load_oracle_weights(a, o)
        helpers.install_replay(a, helpers.ReplayRandom(rec, o.trace, "cuda:0", first_reset=0, first_perm=2))
        o.start()
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        o.rollout(); a._rollout_train(a._steps_per_iter)
        o.build_train_data(); a._build_train_data()
        agents.append(a); oracles.append(o)
    M = agents[0]._mb_rows
    L = _lib.lib()
    for step in range(2):
        # oracle: per-rank backward, averaged gradients, identical AdamW on every rank
        idxs, grads = [], []
        for o in oracles:
            idx = o.sample_idx(M)
            info = o.loss(idx)
            for p in o.params.values():
                p.grad = None
            info["loss"].backward()
            idxs.append(idx)
            grads.append({k: o.params[k].grad.clone() for k in o.names})
        mean_g = {k: sum(g[k] for g in grads) / world for k in oracles[0].names}
        for o in oracles:
            o.adam_steps += 1
            with torch.no_grad():
                for k in o.names:
                    from oracle.add_oracle import adamw_step
                    adamw_step(o.params[k], mean_g[k], o.adam_m[k], o.adam_v[k], o.adam_steps, 1e-4)
        # product: gradient only, all-reduce(SUM) emulated by adding the flat vectors, AdamW with grad_scale 1/world
        for a, idx in zip(agents, idxs):
            _lib.check(L.addk_update_minibatch(_lib.stream(), a._ctx.buf, _lib.ptr(idx.cuda().contiguous()), C.c_int(step),
                                               C.c_int(0)), "addk_update_minibatch")
        total = sum(a._model.flat_grad for a in agents)
        for a in agents:
            a._model.flat_grad.copy_(total)
            opt = a._optimizer
            _lib.check(L.addk_adamw(_lib.stream(), _lib.ptr(a._model.flat), _lib.ptr(a._model.flat_grad), _lib.ptr(opt.exp_avg),
                                    _lib.ptr(opt.exp_avg_sq), C.c_longlong(a._model.num_params), C.c_int(opt.steps + 1),
                                    C.c_double(opt.lr), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-8), C.c_double(0.0),
                                    C.c_double(1.0 / world)), "addk_adamw")
            opt.steps += 1
        assert torch.equal(agents[0]._model.flat, agents[1]._model.flat), "ranks must stay bit-identical"
        gp = dict(agents[0]._model.named_parameters())
        for k in oracles[0].names:
            ref = oracles[0].params[k].detach()
            d = float(torch.linalg.norm(gp[k].detach().double().cpu().flatten() - ref.double().flatten()))
            assert d <= 5e-5 * float(torch.linalg.norm(ref.double())) + 0.05 * 1e-4 * np.sqrt(ref.numel()), (step, k, d)
