"""Run the UNMODIFIED reference (rsamf/add-gym) in this container.  TEST INFRASTRUCTURE ONLY.

Only usable where /root/reference exists (the build container).  It is how the oracle port in
oracle/add_oracle.py is pinned and how tests/golden/*.npz are generated (tests/golden/make_golden.py);
nothing in the product path, in the `-m gpu` tests, in smoke() or in bench.py imports this file.

What it does:
  * puts four stub modules in sys.modules (genesis, hydra(+utils.instantiate), omegaconf,
    matplotlib(+pyplot)) -- none of them is on the hot path;
  * builds the reference config dict from its own four YAML files;
  * hands the reference ``Environment`` our SyntheticEngine through the stubbed ``instantiate``;
  * copies the clip to a writable directory because ``load_motion`` writes a .pkl beside it
    (reference add_gym/anim/motion.py:40-42).
"""
import os
import shutil
import sys
import tempfile
import types

REF_ROOT = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "add_gym"))


_engine_factory = {"fn": None}


def _install_stubs():
    if "add_gym" in sys.modules:
        return
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    sys.path.insert(0, REF_ROOT)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub("genesis")
    hyd = stub("hydra", main=lambda *a, **k: (lambda f: f))
    hyd.utils = stub("hydra.utils", instantiate=lambda cfg, *a, **k: _engine_factory["fn"](cfg))
    stub("omegaconf", DictConfig=dict, OmegaConf=object)
    mpl = stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = stub("matplotlib.pyplot")


def load_reference_config(num_envs, motion_file=None, workdir=None):
    import yaml
    cfg = {}
    for key, rel in (("agent", "agent/add_g1.yaml"), ("engine", "engine/genesis.yaml"),
                     ("robot", "robot/g1.yaml"), ("task", "task/pose.yaml")):
        with open(os.path.join(REF_ROOT, "add_gym/configs", rel), "r") as f:
            cfg[key] = yaml.safe_load(f)
    cfg["engine"]["num_envs"] = int(num_envs)
    cfg["engine"]["enable_viewer"] = True          # the False branch imports pyglet (env.py:29-32)
    cfg["engine"]["enable_video_recording"] = False
    cfg["robot"]["urdf_path"] = os.path.join(REF_ROOT, "assets/g1_description/g1_29.xml")
    workdir = workdir or tempfile.mkdtemp(prefix="addref_")
    if motion_file is None:
        motion_file = "walk1_subject1_trimmed.motion"
    if isinstance(motion_file, (list, tuple)):          # [(file, weight, max_frames|None), ...] -> yaml library
        lines = ["motions:"]
        for name, w, max_frames in motion_file:
            dst = os.path.join(workdir, name)
            with open(os.path.join(REF_ROOT, "assets/motions", name), "r") as fi, open(dst, "w") as fo:
                for i, line in enumerate(fi):
                    if max_frames is not None and i >= max_frames:
                        break
                    fo.write(line)
            lines += ["  - file: \"{}\"".format(dst), "    weight: {}".format(w)]
        lib = os.path.join(workdir, "lib.yaml")
        with open(lib, "w") as f:
            f.write("\n".join(lines) + "\n")
        cfg["task"]["motion_file"] = lib
    else:
        dst = os.path.join(workdir, os.path.basename(motion_file))
        shutil.copy(os.path.join(REF_ROOT, "assets/motions", motion_file), dst)
        os.chmod(dst, 0o644)
        cfg["task"]["motion_file"] = dst
    return cfg


def make_reference_agent(num_envs, seed=0, engine_seed=1234, motion_file=None, fall_prob=0.002,
                         task_overrides=None, agent_overrides=None):
    """Construct the reference ``ADDAgent`` on CPU over a SyntheticEngine.  Returns (agent, cfg)."""
    import torch
    _install_stubs()
    from add_gym_b200.engine import SyntheticEngine

    def factory(engine_cfg):
        return SyntheticEngine(num_envs=engine_cfg["num_envs"], ctrl_dt=engine_cfg["ctrl_dt"],
                               seed=engine_seed, noise_device="cpu", fall_prob=fall_prob, device="cpu")

    _engine_factory["fn"] = factory
    cfg = load_reference_config(num_envs, motion_file)
    cfg["task"].update(task_overrides or {})
    cfg["agent"].update(agent_overrides or {})
    import add_gym.learning.add.add_agent as ref_add_agent
    torch.manual_seed(seed)
    agent = ref_add_agent.ADDAgent(cfg, distributed=False)
    return agent, cfg


if __name__ == "__main__":
    import time
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    agent, cfg = make_reference_agent(n)
    agent._curr_obs, agent._curr_info = agent._reset_envs()
    agent._init_train_ = None
    agent._exp_buffer.clear()
    t0 = time.perf_counter()
    info = agent._train_iter()
    print("one reference iteration at N=%d: %.2fs" % (n, time.perf_counter() - t0))
    print({k: (float(v) if torch.is_tensor(v) else v) for k, v in info.items()})
