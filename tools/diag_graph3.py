import gc, os, sys, types
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from add_gym_b200 import _lib, config as b200_config
from add_gym_b200.add_agent import ADDAgent
def mk(engine):
    cfg = b200_config.default_config(num_envs=4096, mlp_precision="tf32x3")
    cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
    if engine: cfg["engine"]["_target_"] = "add_gym_b200.engine." + engine
    torch.manual_seed(0)
    a = ADDAgent(cfg, device="cuda:0")
    a._curr_obs, a._curr_info = a._reset_envs()
    a._exp_buffer.clear(); a._reset_tracker()
    return a
mode = sys.argv[1]
args = types.SimpleNamespace(precision="tf32x3")
a = mk(None)
if "clock" in mode:
    cs = bench.ClockSampler(0); cs.start()
for _ in range(4 if "iters4" in mode else 3): a._train_iter()
if "events" in mode:
    a.engine_time_events = []
    if "rollonly" in mode: a._rollout_train(32)
    else: a._train_iter()
    torch.cuda.synchronize()
    print(sum(x.elapsed_time(y) for x, y in a.engine_time_events))
    a.engine_time_events = None
if "clock" in mode: print(cs.stop())
if "roof" in mode: print(bench.dominant_kernel_roofline(a, args, 6400.0, 1400.0, "x")["achieved"])
if "hbm" in mode: print(bench.step_kernel_roofline(a, 6400.0)["achieved"])
if "clear" in mode:
    a._graphs_pre.clear(); a._graphs_post.clear(); a._graph_pool = None
del a
if "gc" in mode: gc.collect()
torch.cuda.empty_cache()
b = mk("HostBoundaryEngine")
try:
    for _ in range(3): b._train_iter()
    torch.cuda.synchronize(); print("second agent ok", mode)
except Exception as e:
    print("FAIL", mode, repr(e)[:120])
