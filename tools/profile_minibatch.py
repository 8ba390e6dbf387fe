"""Profiling driver (GPU box): one rollout + build_train_data + a few optimizer steps at full size, nothing else.
    python tools/profile_minibatch.py [envs] [precision] [steps]
Used for the ncu launch list / --set full captures committed under profiles/."""
import ctypes as C, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib, config as b200_config
from add_gym_b200.add_agent import ADDAgent

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = b200_config.default_config(num_envs=envs, mlp_precision=prec)
cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
cfg["agent"]["update_streams"] = int(os.environ.get("STREAMS", "3"))
torch.manual_seed(0)
a = ADDAgent(cfg, device="cuda:0")
a._curr_obs, a._curr_info = a._reset_envs()
a._exp_buffer.clear()
a._rollout_train(a._steps_per_iter)
a._build_train_data()
torch.cuda.synchronize()
_lib.launch_count(reset=True)
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
idxs = [a._exp_buffer.sample_indices(a._mb_rows) for _ in range(steps)]
if not os.environ.get("NO_WARM"):      # (ncu launch lists: NO_WARM=1 keeps the list short)
    for s in range(3):                 # helper streams, tensor maps, sticky scales of the planes-only layers
        _lib.check(_lib.lib().addk_update_minibatch(_lib.stream(), a._ctx.buf, _lib.ptr(idxs[s % steps]), C.c_int(s % a._max_steps), C.c_int(s + 1)), "mb")
    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
torch.cuda.synchronize()
e0.record()
t0 = time.perf_counter()
for s in range(steps):
    _lib.check(_lib.lib().addk_update_minibatch(_lib.stream(), a._ctx.buf, _lib.ptr(idxs[s]), C.c_int(s % a._max_steps), C.c_int(s + 1)), "mb")
host_ms = (time.perf_counter() - t0) * 1e3 / steps
e1.record()
torch.cuda.synchronize()
print("host issue ms/step: %.3f" % host_ms)
print("minibatch ms:", e0.elapsed_time(e1) / steps, "launches/step:", _lib.launch_count() / steps, "loss", float(a._ws["info"][0, 0]))
