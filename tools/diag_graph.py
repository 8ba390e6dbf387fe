import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib, config as b200_config
from add_gym_b200.add_agent import ADDAgent
envs = int(sys.argv[1]); mode = sys.argv[2]
cfg = b200_config.default_config(num_envs=envs, mlp_precision="tf32x3")
cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
if len(sys.argv) > 3:
    cfg["engine"]["_target_"] = "add_gym_b200.engine." + sys.argv[3]
if len(sys.argv) > 4:
    import add_gym_b200.engine as E
    exec(sys.argv[4])
a = ADDAgent(cfg, device="cuda:0")
a._curr_obs, a._curr_info = a._reset_envs()
a._exp_buffer.clear()
try:
    if mode == "iter":
        a._train_iter(); a._train_iter()
    else:
        a._rollout_train(32); a._rollout_train(32)
    torch.cuda.synchronize()
    print("ok", envs, mode, len(a._graphs_pre), len(a._graphs_post))
except Exception as e:
    print("FAIL", envs, mode, repr(e)[:300], _lib.lib().addk_last_error())
