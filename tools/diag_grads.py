"""Diagnostic (GPU box): per-tensor gradient error of the first optimizer step vs the oracle."""
import ctypes as C, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
import test_gpu_parity as T
from parity_helpers import rel_err
from add_gym_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
motion = T.THREE_CLIPS if (len(sys.argv) > 3 and sys.argv[3] == "three") else None
nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
INFO = ["loss", "critic_loss", "actor_loss", "clip_frac", "imp_ratio", "action_bound_loss", "disc_loss",
        "disc_grad_penalty", "disc_logit_loss", "disc_pos_acc", "disc_neg_acc", "disc_pos_logit", "disc_neg_logit"]
oracle, agent, rec = T._pair(n, motion, precision=prec)
T._start(oracle, agent)
oracle.rollout(); agent._rollout_train(agent._steps_per_iter)
oracle.build_train_data(); agent._build_train_data()
gp = dict(agent._model.named_parameters())
snap = {}
def grad_hook(g):
    snap["pre"] = {k: oracle.params[k].detach().clone() for k in oracle.names}
def on_step(step, idx, oinfo, o):
    for k in o.names:
        gp[k].data.copy_(snap["pre"][k])
    _lib.check(_lib.lib().addk_update_minibatch(_lib.stream(), agent._ctx.buf, _lib.ptr(idx.cuda().contiguous()), C.c_int(step), C.c_int(agent._optimizer.steps + 1)), "mb")
    agent._optimizer.steps += 1
    row = agent._ws["info"][step].cpu()
    print("step", step, " ".join("%s %.8g/%.8g" % (k, float(row[i]), float(oinfo[k])) for i, k in enumerate(INFO)
                                 if abs(float(row[i]) - float(oinfo[k])) > 1e-6 * max(1.0, abs(float(oinfo[k])))))
    for k in o.names:
        if rel_err(gp[k].grad, o.params[k].grad) > 1e-5 or rel_err(gp[k], o.params[k]) > 1e-5:
            print("  %-36s grad rel %.3e  |g| %.3e   param rel %.3e" % (k, rel_err(gp[k].grad, o.params[k].grad), float(o.params[k].grad.norm()), rel_err(gp[k], o.params[k])))
oracle.update_model(on_step=on_step, grad_hook=grad_hook, max_steps=nsteps)
