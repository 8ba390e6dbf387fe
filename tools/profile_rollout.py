"""Profiling driver (GPU box): a few rollout steps at full size (eager launches, so an ncu launch list names every kernel).
    python tools/profile_rollout.py [envs] [precision] [steps]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib, config as b200_config
from add_gym_b200.add_agent import ADDAgent
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = b200_config.default_config(num_envs=envs, mlp_precision=prec)
cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
cfg["agent"]["cuda_graphs"] = False
torch.manual_seed(0)
a = ADDAgent(cfg, device="cuda:0")
a._curr_obs, a._curr_info = a._reset_envs()
a._exp_buffer.clear()
a._rollout_train(2)
torch.cuda.synchronize()
a._exp_buffer.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); t0 = time.perf_counter()
a._rollout_train(steps)
host = (time.perf_counter() - t0) * 1e3 / steps
e1.record(); torch.cuda.synchronize()
print("rollout step: %.1f us GPU (eager, incl. the synthetic engine), host issue %.1f us" % (e0.elapsed_time(e1) * 1e3 / steps, host * 1e3))
