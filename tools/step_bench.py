import sys, os
sys.path.insert(0, os.getcwd())
import torch, json
import bench
from add_gym_b200 import config as b200_config
from add_gym_b200.add_agent import ADDAgent
for envs in (4096, 32768):
    cfg = b200_config.default_config(num_envs=envs, mlp_precision="f16x3")
    cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
    torch.manual_seed(0)
    a = ADDAgent(cfg, device="cuda:0")
    a._curr_obs, a._curr_info = a._reset_envs(); a._exp_buffer.clear()
    rs = [bench.step_kernel_roofline(a, 6446.9) for _ in range(3)]
    print(envs, ["%.2f us %.3f" % (r["avg_launch_ms"] * 1e3, r["frac"]) for r in rs])
    del a
