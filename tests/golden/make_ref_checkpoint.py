"""Generate tests/golden/ref_checkpoint.zip: a checkpoint WRITTEN BY the unmodified reference (needs /root/reference).

    python tests/golden/make_ref_checkpoint.py

The real `ADDAgent` (reference add_gym/learning/add/add_agent.py) runs one `_train_iter()` on CPU (4 envs:
the optimizer then owns a state entry per trainable tensor, 40 steps old), every floating tensor of the model
and of the optimizer state is then overwritten IN PLACE with a low-entropy pattern (the file has to stay small: 52 MB of
fp32 noise would not) and the reference's own `save()` (base_agent.py:148-155) writes the file, which is stored deflated.
`pattern()` below is what tests/test_gpu_parity.py::test_loads_a_checkpoint_written_by_the_reference recomputes.
"""
import os
import sys
import zipfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import ref_harness  # noqa: E402


def pattern(numel, kind, index):
    """kind 0: parameter, 1: exp_avg, 2: exp_avg_sq of trainable tensor `index` (the reference's optimizer index) --
    exact in fp32, a few hundred distinct values, different for every tensor"""
    i = torch.arange(numel, dtype=torch.int64) + 17 * index
    if kind == 0:
        return ((i % 251) - 125).to(torch.float32) / 1024.0
    if kind == 1:
        return ((i % 127) - 63).to(torch.float32) / 65536.0
    return ((i % 61) + 1).to(torch.float32) / 1048576.0


def main():
    agent, cfg = ref_harness.make_reference_agent(4, seed=0, engine_seed=1234, fall_prob=0.01)
    agent._curr_obs, agent._curr_info = agent._reset_envs()
    agent._init_train_done = True
    agent._exp_buffer.clear()
    agent._train_return_tracker.reset()
    agent._train_iter()
    agent._iter, agent._sample_count = 7, 7 * 128
    with torch.no_grad():
        params = [p for p in agent._optimizer._optimizer.param_groups[0]["params"]]
        for idx, p in enumerate(params):
            st = agent._optimizer._optimizer.state[p]
            p.copy_(pattern(p.numel(), 0, idx).view(p.shape))
            st["exp_avg"].copy_(pattern(p.numel(), 1, idx).view(p.shape))
            st["exp_avg_sq"].copy_(pattern(p.numel(), 2, idx).view(p.shape))
    tmp = os.path.join(HERE, "_ref_model.pt")
    agent.save(tmp)
    out = os.path.join(HERE, "ref_checkpoint.zip")
    with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED, compresslevel=9) as z:
        z.write(tmp, "model.pt")
    os.remove(tmp)
    ck_steps = {float(s["step"]) for s in agent._optimizer._optimizer.state.values()}
    print("wrote", out, os.path.getsize(out), "bytes; optimizer steps", ck_steps, "lr", agent._optimizer._optimizer.param_groups[0]["lr"])


if __name__ == "__main__":
    main()
