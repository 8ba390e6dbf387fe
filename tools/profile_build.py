"""Profiling driver (GPU box): one rollout, then build_train_data a few times at full size (ncu launch list of that stage).
    python tools/profile_build.py [envs] [precision] [reps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import config as b200_config
from add_gym_b200.add_agent import ADDAgent
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = b200_config.default_config(num_envs=envs, mlp_precision=prec)
cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
torch.manual_seed(0)
a = ADDAgent(cfg, device="cuda:0")
a._curr_obs, a._curr_info = a._reset_envs()
a._exp_buffer.clear()
a._rollout_train(a._steps_per_iter)
a._build_train_data()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    a._build_train_data()
e1.record(); torch.cuda.synchronize()
print("build_train_data: %.2f ms" % (e0.elapsed_time(e1) / reps))
