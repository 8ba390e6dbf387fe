"""Test plumbing: record the CPU oracle's random draws and replay them into the CUDA path, so that both
consume IDENTICAL noise / reset samples / permutations (torch's CPU and CUDA generators are different
streams, so equal seeds alone cannot give that)."""
import numpy as np
import torch

from oracle import add_oracle


class RecordRandom(add_oracle.TorchRandom):
    def __init__(self):
        self.noise, self.masks, self.perms = [], [], []
        self.clip_draws, self.segment_draws, self.uniform_draws = [], [], []

    def motions(self, weights, n):
        v = super().motions(weights, n)
        self.clip_draws.append(v.clone())
        return v

    def segments(self, probs):
        v = super().segments(probs)
        self.segment_draws.append(v.clone())
        return v

    def uniform(self, n):
        v = super().uniform(n)
        self.uniform_draws.append(v.clone())
        return v

    def action_noise(self, n, dim):
        v = super().action_noise(n, dim)
        self.noise.append(v.clone())
        return v

    def exp_mask(self, n, prob):
        v = super().exp_mask(n, prob)
        self.masks.append(v.clone())
        return v

    def randperm(self, n):
        v = super().randperm(n)
        self.perms.append(v.clone())
        return v


class ReplayRandom:
    """Drop-in for add_gym_b200.add_agent.DeviceRandom fed from a RecordRandom + the oracle's reset trace."""

    def __init__(self, rec, trace, device, first_reset, first_noise=0, first_perm=0):
        self.rec, self.trace, self.device = rec, trace, device
        self.i_noise, self.i_mask, self.i_perm, self.i_reset = first_noise, first_noise, first_perm, first_reset

    def action_noise(self, n, dim):
        v = self.rec.noise[self.i_noise]
        self.i_noise += 1
        assert v.shape == (n, dim)
        return v.to(self.device)

    def exp_mask(self, n, prob):
        v = self.rec.masks[self.i_mask]
        self.i_mask += 1
        return v.to(self.device).contiguous()

    def randperm(self, n):
        v = self.rec.perms[self.i_perm]
        self.i_perm += 1
        assert v.numel() == n
        return v.to(self.device)

    def reset_uniforms(self, n):
        raise AssertionError("replayed runs inject reset candidates directly")

    def fill_candidates(self, core):
        """Scatter the oracle's next reset draw (env ids, clip ids, start times) into the per-env arrays."""
        envs = self.trace["reset_envs"][self.i_reset]
        ids = self.trace["reset_ids"][self.i_reset]
        times = self.trace["reset_times"][self.i_reset]
        self.i_reset += 1
        if len(envs) > 0:
            e = envs.to(self.device)
            core.new_ids[e] = ids.to(self.device)
            core.new_times[e] = times.to(self.device)
        return envs


def install_replay(agent, replay):
    agent.rng = replay
    agent._fill_reset_candidates = lambda done: replay.fill_candidates(agent._core)


def load_oracle_weights(agent, oracle):
    sd = agent.state_dict()
    for k in oracle.names:
        sd["_model." + k].copy_(oracle.params[k].detach().to(sd["_model." + k].device))


def rel_err(a, b):
    """norm-wise relative error ||a-b|| / ||b|| in float64."""
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    den = float(torch.linalg.norm(b))
    return float(torch.linalg.norm(a - b)) / (den if den > 0 else 1.0)


def max_rel(a, b):
    """max |a-b| / max|b|"""
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    den = float(b.abs().max()) if b.numel() else 1.0
    return float((a - b).abs().max()) / (den if den > 0 else 1.0) if b.numel() else 0.0
