"""GPU box: N training iterations in a row at full size (default 300 at 4096 envs), checking every 25th iteration that the
diagnostics and the parameters are finite -- sticky plane scales, bit planes and the permutation walk over many
optimizer steps.      python tools/soak.py [iters] [envs] [precision]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import config as b200_config
from add_gym_b200.add_agent import ADDAgent
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prec = sys.argv[3] if len(sys.argv) > 3 else "f16x3"
cfg = b200_config.default_config(num_envs=envs, mlp_precision=prec)
cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
torch.manual_seed(0)
a = ADDAgent(cfg, device="cuda:0")
a._curr_obs, a._curr_info = a._reset_envs()
a._exp_buffer.clear()
a._reset_tracker()
t0 = time.perf_counter()
for it in range(iters):
    info = a._train_iter()
    if it % 25 == 0 or it == iters - 1:
        vals = {k: float(v) for k, v in info.items()}
        bad = [k for k, v in vals.items() if v != v or abs(v) == float("inf")]
        ok = bool(torch.isfinite(a._model.flat).all())
        print("iter %4d  %s%s" % (it, " ".join("%s=%.4g" % (k, vals[k]) for k in list(vals)[:8]), "" if ok and not bad else "  NON-FINITE %s" % bad), flush=True)
        if bad or not ok:
            sys.exit(1)
torch.cuda.synchronize()
print("soak ok: %d iterations, %.1f s, |params| max %.3f" % (iters, time.perf_counter() - t0, float(a._model.flat.abs().max())))
