"""Profiling driver (GPU box): the fused per-env step kernel at a given env count, a few launches, nothing else.
    python tools/profile_step.py [envs] [reps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib, config as b200_config
from add_gym_b200.add_agent import ADDAgent
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = b200_config.default_config(num_envs=envs, mlp_precision="f16x3")
cfg["engine"].update(seed=1234, noise_device="device", fall_prob=0.002)
torch.manual_seed(0)
a = ADDAgent(cfg, device="cuda:0")
a._curr_obs, a._curr_info = a._reset_envs()
a._exp_buffer.clear()
core = a._core
flags = _lib.F_ADVANCE | _lib.F_UPDATE_MOTION | _lib.F_REWARD_DONE
for t in range(3):
    a._env.scene.step()
    core.step(flags, exp_row=a._exp_row(t))
torch.cuda.synchronize()
rows = [a._exp_row(t % 32) for t in range(reps)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import ctypes as C
sim = core.sim_struct()
L = _lib.lib()
e0.record()
for t in range(reps):      # back-to-back launches: the GPU never waits for Python
    L.addk_env_step(_lib.stream(), C.byref(core.task), C.byref(core.c_lib), C.byref(sim), C.byref(core.c_env), C.byref(rows[t]),
                    _lib.ptr(core.dof_err_w), None, C.c_int(core.N), C.c_int(t % 3), C.c_int(flags))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("env_step_kernel N=%d: %.1f us per launch, %.0f GB/s algorithmic (5624 B/env-step)" % (envs, ms * 1e3, 5624.0 * envs / ms / 1e6))
