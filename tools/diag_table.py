"""Diagnostic (GPU box): column-wise differences between the CUDA-built step table and the oracle's."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import config as b200_config
from add_gym_b200.env import ImitationEnvironment
from add_gym_b200.add_motion import ADDMotion
from oracle import harness

cfg = b200_config.default_config(num_envs=4)
env = ImitationEnvironment(cfg, "cuda:0")
lib = ADDMotion(cfg["task"], env, "cuda:0").motion_lib
olib = harness.make_oracle_lib(cfg)
D, h = lib._num_dofs, lib._row_stride // 2
tab = lib.step_table.cpu()
got = torch.cat([tab[:, :7 + D], tab[:, h:h + 6 + D]], dim=1)
d = (got - olib.table).abs()
cm = d.max(0).values
print("per-column max abs diff (pos3 rot4 dof29 | vel3 ang3 dofvel29):")
print([float("%.2e" % v) for v in cm.tolist()])
worst = torch.nonzero(d > 1e-5)
print("entries > 1e-5:", worst.shape[0], "of", d.numel())
for r, c in worst[:20].tolist():
    fi = olib.frame_idx[r].tolist()
    print("row", r, "col", c, "got", float(got[r, c]), "ref", float(olib.table[r, c]), "frames", fi)
print("frame idx equal:", torch.equal(lib._frame_idx.cpu(), olib.frame_idx))
