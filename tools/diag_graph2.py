import gc, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib, config as b200_config
from add_gym_b200.add_agent import ADDAgent
def mk(engine):
    cfg = b200_config.default_config(num_envs=4096, mlp_precision="tf32x3")
    cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
    if engine: cfg["engine"]["_target_"] = "add_gym_b200.engine." + engine
    torch.manual_seed(0)
    a = ADDAgent(cfg, device="cuda:0")
    a._curr_obs, a._curr_info = a._reset_envs()
    a._exp_buffer.clear()
    return a
mode = sys.argv[1]
a = mk(None)
for _ in range(3): a._train_iter()
torch.cuda.synchronize(); print("first agent ok", len(a._graphs_post))
if mode == "keep": keep = a
del a
if mode == "gc": gc.collect()
torch.cuda.empty_cache()
b = mk("HostBoundaryEngine")
try:
    for _ in range(3): b._train_iter()
    torch.cuda.synchronize(); print("second agent ok", mode, len(b._graphs_post))
except Exception as e:
    print("FAIL", mode, repr(e)[:200])
