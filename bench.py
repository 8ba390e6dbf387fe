"""bench.py -- add-gym rollout + update hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs 4096] [--precision f16x3|bf16|fp32] [--motions walk|all]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU implementation (oracle port), host cores

One "step" = one training iteration of the reference agent (`BaseAgent._train_iter`, base_agent.py:353-374):
a 32-step rollout over N envs, `_build_train_data`, 5 epochs x 8 minibatches of the ADD/PPO update and the
normalizer update -- T*N = 131,072 env-steps at N = 4096.  Physics is excluded: the synthetic engine's
`scene.step()` is timed with its own CUDA events and subtracted.  metric = env-steps/s = T*N*R / t.

The JSON line carries, besides the contract's keys: `roofline` (dominant kernel: the 16384x1024x1024 dense layer, timed
alone -> burst tensor peak), `roofline_with_prepass` (the same layer including the conversion of its activation operand),
`roofline_in_step` (the whole update stage's algorithmic flops / its time -> sustained peak), `roofline_hbm_stage` (fused
step kernel), `cpu_baseline` (oracle port on the host cores), `gpu_eager_baseline` (the same port in PyTorch eager ON THIS
GPU with TF32 allowed, as the reference ships: add_gym/main.py:16-18 -- what a user of the reference has today),
`secondary` (BASELINE configs[2], [3], [4] measured in the same job) and `params_identical_across_ranks`.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "env-steps/s (rollout+update, physics excl.)"
UNIT = "env-steps/s"
T_ITER = 32                 # steps_per_iter of the reference config (add_g1.yaml)
REF_SAMPLE_STEPS = 4        # rollout steps of one bounded CPU sample (see cpu_sample_rate)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (BASELINE configs[1]: 4096)")
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "tf32x3", "tf32", "bf16", "f16x3"],
                    help="MLP arithmetic: f16x3 = tcgen05 kind::f16 on fp16 hi/lo planes (fp32-parity mode, default); bf16 = "
                         "configs[3]; fp32 = CUDA cores; tf32x3 / tf32 = superseded kernels (make LEGACY=1, ADDK_LIB=...)")
    ap.add_argument("--motions", default="walk", choices=["walk", "all"],
                    help="walk = walk1_subject1_trimmed (configs[1]); all = the 42-clip library (configs[2], 261 MB step table)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[2] / [3] / [4] lines")
    ap.add_argument("--stages-out", default=os.path.join(REPO, "gpurun_out", "bench_stages.json"))
    return ap.parse_args()


def motion_file_of(args):
    from add_gym_b200 import config as b200_config
    return os.path.join(b200_config.ASSET_DIR, "motions_all.addkc") if args.motions == "all" else None


def workload_config(envs, precision, motions, world):
    """The `config` object: IDENTICAL for the B200 arm and the reference arm of one invocation (what is measured, not how)."""
    which = {4096: "BASELINE configs[1]", 8192: "BASELINE configs[3] shape", 32768: "BASELINE configs[4] shape"}.get(envs, "BASELINE configs[1] at another env count")
    if motions == "all":
        which = "BASELINE configs[2]"
    clip = "all 42 assets/motions clips (906k-row step table)" if motions == "all" else "walk1_subject1_trimmed"
    return {"workload": "%s: G1 %s, %d envs/GPU, %s MLPs, one iteration = 32-step rollout + build_train_data + 5x8 ADD/PPO "
                        "minibatches of %d rows; synthetic engine stands in for Genesis (not installed), scene.step() timed "
                        "separately and excluded" % (which, clip, envs, "fp32" if precision in ("f16x3", "tf32x3", "fp32") else precision, 4 * envs),
            "envs_per_gpu": envs, "steps_per_iter": T_ITER, "minibatch": 4 * envs, "optimizer_steps": 40, "motions": motions,
            "mlp_arithmetic": "fp32" if precision in ("f16x3", "tf32x3", "fp32") else precision,
            "parallelism": "dp%d" % world,
            "l2": "working set per iteration (experience buffers %.0f MB + activations) exceeds the 126 MB L2" %
                  (T_ITER * envs * (264 * 2 + 114 * 2 + 29 + 8) * 4 / 1e6)}


# ---------------------------------------------------------------------------------------------------------------
# CPU / GPU-eager baselines: the oracle port of the reference (oracle/ may only be executed here as the baseline /
# checker, never on the product path)
# ---------------------------------------------------------------------------------------------------------------
def _oracle_iteration_times(agent, n_warm, n_timed, sync=None, budget_s=None):
    times = []
    t_begin = time.perf_counter()
    for i in range(n_warm + n_timed):
        phys = [0.0]
        scene = agent.env.scene
        orig = scene.step

        def timed_step(orig=orig, phys=phys):
            if sync:
                sync()
            t0 = time.perf_counter()
            orig()
            if sync:
                sync()
            phys[0] += time.perf_counter() - t0
        scene.step = timed_step
        if sync:
            sync()
        t0 = time.perf_counter()
        agent.train_iter()
        if sync:
            sync()
        dt = time.perf_counter() - t0 - phys[0]
        scene.step = orig
        if i >= n_warm:
            times.append(dt)
            if budget_s is not None and time.perf_counter() - t_begin > budget_s:
                break
    return times


def cpu_sample_rate(envs, motion_file, steps, warmup, budget_s):
    """The reference's algorithm (oracle port, torch CPU fp32, all host threads) on a BOUNDED SAMPLE of the workload: the
    full env count and the full 4*envs-row minibatch, but REF_SAMPLE_STEPS = 4 of the 32 rollout steps per sample --
    i.e. the reference agent with steps_per_iter = 4: 4 env steps over all envs, build_train_data over those 4*envs rows,
    then 5 epochs x 1 minibatch of 4*envs rows.  Every stage's cost per env-step is that of the full iteration (each
    collected sample is still visited 5 times in full-size minibatches); one full 32-step iteration at 4096 envs takes
    ~40 s on 16 host threads, a sample ~5 s."""
    import torch
    from add_gym_b200 import config as b200_config
    from oracle import harness
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = b200_config.default_config(num_envs=envs, motion_file=motion_file)
    cfg["agent"]["steps_per_iter"] = REF_SAMPLE_STEPS
    agent = harness.make_oracle_agent(envs, seed=0, engine_seed=1234, cfg=cfg, fall_prob=0.002)
    agent.start()
    times = _oracle_iteration_times(agent, warmup, steps, budget_s=budget_s)
    mean = sum(times) / len(times)
    return REF_SAMPLE_STEPS * envs / mean, mean, cores, len(times)


def cpu_sample_text(envs, cores, n_timed, sec):
    return ("oracle port of the reference agent (torch CPU fp32, %d threads) at the full %d envs and %d-row minibatches; one "
            "sample = the reference iteration with steps_per_iter = %d instead of 32 (%d env steps, build_train_data over "
            "them, 5 epochs x 1 minibatch): the same work per env-step as the full iteration; %d timed samples of %.1f s" %
            (cores, envs, 4 * envs, REF_SAMPLE_STEPS, REF_SAMPLE_STEPS, n_timed, sec))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    value, sec, cores, n = cpu_sample_rate(args.envs, motion_file_of(args), args.steps, min(args.warmup, 2), budget_s=240.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.envs, args.precision, args.motions, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(args.envs, cores, n, sec), "timed_samples": n},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_eager_rate(envs, motion_file, dev):
    """The same-box comparator SURVEY 8(d) asks for: the reference's algorithm (oracle port) in PyTorch eager ON THE B200,
    TF32 matmuls allowed exactly as the reference sets them (add_gym/main.py:16-18), full iteration (T = 32) at the full env
    count, 1 warm-up + 2 timed iterations, physics excluded the same way."""
    import torch
    from add_gym_b200 import config as b200_config
    from oracle import harness
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        cfg = b200_config.default_config(num_envs=envs, motion_file=motion_file)
        agent = harness.make_oracle_agent(envs, seed=0, engine_seed=1234, cfg=cfg, fall_prob=0.002, device=dev)
        with torch.device(dev):
            agent.start()
            times = _oracle_iteration_times(agent, 1, 2, sync=torch.cuda.synchronize)
        mean = sum(times) / len(times)
        return {"value": T_ITER * envs / mean, "unit": UNIT, "ms_per_step": mean * 1e3, "kind": "port",
                "what": "oracle port of the reference agent run in PyTorch eager on this GPU (torch %s, TF32 allowed as in "
                        "add_gym/main.py:16-18), full iteration at %d envs, 1 warm-up + 2 timed, physics excluded" %
                        (torch.__version__, envs)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def flops_per_iteration(N, T=T_ITER, epochs=5):
    """MACs/sample actor 1,858,048, critic 1,843,712, disc 641,536."""
    rollout = 2 * 1858048 * T * N
    build = 2 * (2 * 1843712 + 641536) * T * N
    update = 2 * 14413824 * (epochs * T * N)          # 160*N minibatch samples per iteration
    return rollout, build, update


def run_b200(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from add_gym_b200 import _lib
    from add_gym_b200 import config as b200_config
    from add_gym_b200.add_agent import ADDAgent

    peaks = {}
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            peaks = json.load(f)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_burst = float(peaks.get("bf16_tflops", 1650.0))
    tc_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    def make_agent(envs, precision, motion_file, engine_cls=None):
        cfg = b200_config.default_config(num_envs=envs, mlp_precision=precision, motion_file=motion_file)
        cfg["engine"].update(seed=1234 + rank, noise_device="device", fall_prob=0.002)
        if engine_cls is not None:
            cfg["engine"]["_target_"] = "add_gym_b200.engine." + engine_cls
        torch.manual_seed(0)
        a = ADDAgent(cfg, distributed=world > 1, device=dev)
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        a._reset_tracker()
        return a

    def drop(agent):
        agent.release_graphs()
        del agent
        gc.collect()
        torch.cuda.empty_cache()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def to_host(info):
        keys = list(info.keys())
        row = torch.stack([torch.as_tensor(info[k], device=dev).to(torch.float64).reshape(()) for k in keys])
        return dict(zip(keys, row.tolist()))

    def timed(agent, K, W, e2e):
        """K timed calls of ADDAgent._train_iter (the call a user's train_model makes per iteration).
        returns (seconds total excl. physics, physics seconds, per-stage seconds, launches, last info on the host)"""
        stage = {"rollout": 0.0, "build_train_data": 0.0, "update": 0.0, "normalizers": 0.0}
        ev = lambda: torch.cuda.Event(enable_timing=True)
        info_host = None
        for _ in range(max(W, 2)):      # at least two: the second rollout captures the CUDA graphs of the rollout step
            info = agent._train_iter()
            if e2e:
                info_host = to_host(info)
        barrier()
        agent.engine_time_events = []
        agent.stage_events = []
        _lib.launch_count(reset=True)
        t_begin, t_end = ev(), ev()
        t_begin.record()
        for _ in range(K):
            info = agent._train_iter()
            if e2e:   # the user-visible result of the iteration, read back to the host (one D2H copy of the stacked row)
                info_host = to_host(info)
                bad = [k for k, v in info_host.items() if v != v or v in (float("inf"), float("-inf"))]
                if bad:
                    raise SystemExit("bench.py: non-finite diagnostics %s -- the timed path is numerically broken" % bad)
        t_end.record()
        barrier()
        if not e2e:   # outside the timed region: the last iteration's diagnostics and the parameters must be finite
            last = to_host(info)
            bad = [k for k, v in last.items() if v != v or v in (float("inf"), float("-inf"))]
            if bad or not bool(torch.isfinite(agent._model.flat).all()):
                raise SystemExit("bench.py: non-finite diagnostics %s / parameters -- the timed path is numerically broken" % bad)
        total = t_begin.elapsed_time(t_end) * 1e-3
        phys = sum(a.elapsed_time(b) for a, b in agent.engine_time_events) * 1e-3
        marks = agent.stage_events
        agent.engine_time_events, agent.stage_events = None, None
        for it in range(K):
            m = marks[5 * it:5 * it + 5]
            for i, k in enumerate(("rollout", "build_train_data", "update", "normalizers")):
                stage[k] += m[i].elapsed_time(m[i + 1]) * 1e-3
        stage["rollout"] -= phys
        launches = _lib.launch_count()
        return total - phys, phys, stage, launches, info_host

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def params_identical(agent):
        """After the timed region every rank must hold bit-identical weights (same all-reduced gradient, same AdamW)."""
        if world == 1:
            return True
        flat = agent._model.flat.view(torch.int32)
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))

    K, W = args.steps, max(args.warmup, 0)
    N = args.envs
    mfile = motion_file_of(args)
    # ---- device-resident arm -------------------------------------------------------------------------------
    agent = make_agent(N, args.precision, mfile)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    sec, phys, stage, launches, _ = timed(agent, K, W, e2e=False)
    clk = clocks.stop() if rank == 0 else None
    sec = max_over_ranks(sec)
    value = T_ITER * N * world * K / sec
    stage_max = {k: max_over_ranks(v) for k, v in stage.items()}
    same_params = params_identical(agent)

    # ---- dominant kernel: the dense-layer contraction of the update, timed launch by launch --------------------
    roof = roof_pre = hbm_roof = None
    if rank == 0:
        roof, roof_pre = dominant_kernel_roofline(agent, args, tc_burst, peak_src)
        hbm_roof = step_kernel_roofline(agent, hbm_peak, args.motions)
    drop(agent)

    # ---- end to end: simulator state arrives from pinned host memory every env step --------------------------
    e2e = None
    if not args.no_e2e:
        agent = make_agent(N, args.precision, mfile, "HostBoundaryEngine")
        ent = agent._env.robot.entity
        sec_e, phys_e, _, _, info_host = timed(agent, K, W, e2e=True)
        sec_e = max_over_ranks(sec_e)
        info_bytes = 8 * len(info_host or {})
        e2e = {"value": T_ITER * N * world * K / sec_e, "unit": UNIT,
               "h2d_bytes_per_step": int(ent.h2d_bytes_per_env_step * T_ITER), "d2h_bytes_per_step": int(ent.d2h_bytes_per_env_step * T_ITER + info_bytes),
               "what": "ADDAgent._train_iter with the simulator state copied from pinned host memory before every env "
                       "step, the action copied back to pinned host memory after every actor forward and the "
                       "iteration diagnostics read back as Python floats"}
        drop(agent)

    # ---- secondary configurations of BASELINE.json, same job, same timing rules (short: 2 warm-up + 3 timed) ------
    secondary = []
    if not args.no_secondary and N == 4096 and args.motions == "walk" and args.precision == "f16x3":
        for envs, prec, motions, name in ((4096, "f16x3", "all", "configs[2]"), (8192, "bf16", "walk", "configs[3]"),
                                          (32768, "f16x3", "walk", "configs[4]")):
            try:
                mf = os.path.join(b200_config.ASSET_DIR, "motions_all.addkc") if motions == "all" else None
                a2 = make_agent(envs, prec, mf)
                s2, p2, st2, _, _ = timed(a2, 3, 2, e2e=False)
                s2 = max_over_ranks(s2)
                st2 = {k: max_over_ranks(v) for k, v in st2.items()}
                h2 = step_kernel_roofline(a2, hbm_peak, motions) if rank == 0 else None
                fl2 = flops_per_iteration(envs)
                secondary.append({"config": workload_config(envs, prec, motions, world), "baseline_config": name,
                                  "value": T_ITER * envs * world * 3 / s2, "unit": UNIT, "ms_per_step": s2 / 3 * 1e3, "steps": 3, "warmup": 2,
                                  "stages_ms": {k: v / 3 * 1e3 for k, v in st2.items()},
                                  "ppo_update_samples_per_s": 5 * T_ITER * envs * world * 3 / max(st2["update"], 1e-9),
                                  "update_tflops": fl2[2] * 3 / max(st2["update"], 1e-9) / 1e12,
                                  "roofline_hbm_stage": h2, "params_identical_across_ranks": params_identical(a2)})
                drop(a2)
            except Exception as e:      # a secondary line must never take the headline down
                secondary.append({"baseline_config": name, "error": "%s: %s" % (type(e).__name__, e)})
                gc.collect()
                torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- baselines (rank 0, N=1 only) ---------------------------------------------------------------------------
    cpu = gpu_eager = None
    if world == 1 and not args.no_gpu_eager:
        try:
            gpu_eager = gpu_eager_rate(N, mfile, dev)
        except Exception as e:
            gpu_eager = {"error": "%s: %s" % (type(e).__name__, e)}
        gc.collect()
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        v, s, cores, n = cpu_sample_rate(N, mfile, 3, 1, budget_s=60.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_text(N, cores, n, s), "timed_samples": n}

    fl = flops_per_iteration(N)
    upd_tflops = fl[2] * K / max(stage_max["update"], 1e-9) / 1e12
    stages = {
        "envs_per_gpu": N, "n_gpus": world, "precision": args.precision, "seconds_per_iteration": sec / K,
        "physics_seconds_per_iteration_excluded": phys / K,
        "stages_seconds_per_iteration": {k: v / K for k, v in stage_max.items()},
        "rollout_env_steps_per_s": T_ITER * N * world * K / max(stage_max["rollout"], 1e-9),
        "ppo_update_samples_per_s": 5 * T_ITER * N * world * K / max(stage_max["update"], 1e-9),
        "tflops": {"rollout_actor": fl[0] * K / max(stage_max["rollout"], 1e-9) / 1e12,
                   "build_train_data": fl[1] * K / max(stage_max["build_train_data"], 1e-9) / 1e12,
                   "update": upd_tflops},
        "launches_per_iteration": launches / K,
    }
    roof_step = {"bound": "tensor", "kernel": "update stage: 40 optimizer steps (every launch of it, not only the dense layers)",
                 "achieved": upd_tflops, "peak": tc_sust, "unit": "TFLOP/s", "frac": upd_tflops / tc_sust,
                 "peak_source": peak_src + ", sustained bf16 (timed inside a long step)",
                 "algorithmic_flops_per_iteration": fl[2], "ms_per_optimizer_step": stage_max["update"] / K / 40 * 1e3}
    try:
        os.makedirs(os.path.dirname(args.stages_out), exist_ok=True)
        with open(args.stages_out, "w") as f:
            json.dump({"stages": stages, "roofline": roof, "roofline_hbm_stage": hbm_roof, "clocks": clk}, f, indent=1)
    except OSError:
        pass
    print("stages: " + json.dumps(stages), file=sys.stderr, flush=True)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": sec / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split)", "tf32": "tf32",
                  "bf16": "bf16 (fp32 accumulate, fp32 master weights)",
                  "f16x3": "f32 (fp16 hi/lo split on the tensor cores, fp32 accumulate)"}[args.precision],
        "data": "synthetic",
        "config": workload_config(N, args.precision, args.motions, world),
        "timed_call": "ADDAgent._train_iter",
        "ppo_update_samples_per_s": stages["ppo_update_samples_per_s"],
        "physics_ms_per_step_excluded": phys / K * 1e3,
        "stages_ms": {k: v / K * 1e3 for k, v in stage_max.items()},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "roofline_with_prepass": roof_pre,
        "roofline_in_step": roof_step, "roofline_hbm_stage": hbm_roof, "cpu_baseline": cpu, "gpu_eager_baseline": gpu_eager,
        "secondary": secondary, "params_identical_across_ranks": same_params,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _ncu_traffic(key):
    """dram bytes read + written per launch from the committed `ncu --set full` captures (profiles/r02_ncu_metrics.json,
    else round 1's)."""
    for name in ("r02_ncu_metrics.json", "r01_ncu_metrics.json"):
        try:
            with open(os.path.join(REPO, "profiles", name)) as f:
                m = json.load(f)[key]
            return int(m["dram_bytes_read"]) + int(m["dram_bytes_write"])
        except (OSError, KeyError, ValueError):
            continue
    return None


def step_kernel_roofline(agent, hbm_peak, motions="walk"):
    """The HBM-bound stage: the fused per-env step kernel (csrc/step.cu), 5,624 algorithmic bytes per env-step
    (SURVEY 8d).  `reps` back-to-back launches through the C-ABI (structs prepared beforehand, successive experience
    rows), captured into one CUDA graph; the replay sits between one CUDA-event pair on the launch stream -> average
    launch duration."""
    import ctypes as C

    import torch
    from add_gym_b200 import _lib
    core, N, T = agent._core, agent.get_num_envs(), agent._steps_per_iter
    flags = _lib.F_ADVANCE | _lib.F_UPDATE_MOTION | _lib.F_REWARD_DONE
    for t in range(3):
        core.step(flags, exp_row=agent._exp_row(t % T))
    torch.cuda.synchronize()
    reps = 20
    rows = [agent._exp_row(i % T) for i in range(reps)]
    sim = core.sim_struct()
    L = _lib.lib()

    def issue():
        for i in range(reps):
            L.addk_env_step(_lib.stream(), C.byref(core.task), C.byref(core.c_lib), C.byref(sim), C.byref(core.c_env),
                            C.byref(rows[i]), _lib.ptr(core.dof_err_w), None, C.c_int(N), C.c_int(i % 3), C.c_int(flags))

    # The rollout runs this kernel from a captured CUDA graph (add_agent._capture_all); so does the measurement: `reps`
    # launches captured once, the replay timed -- at 4096 envs the kernel (about 10 us) is shorter than one ctypes call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        issue()
    # warm-up replays: the kernel is issue-bound, i.e. it follows the SM clock, and right after the update stage's
    # tensor load the clock is still at its power-capped 1.3-1.4 GHz (the rollout itself runs at full clock)
    for _ in range(40):
        g.replay()
    # back to episode time 0: the launches above ran 800+ env steps without resets, and envs past the end of their clip
    # report done on every step -- three double atomics each on the return tracker's three words (17.6 vs 8.1 us at 4096 envs)
    agent._env.time_buf.zero_()
    torch.cuda.synchronize()
    runs = []
    for _ in range(5):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        runs.append(e0.elapsed_time(e1) / reps)
    ms = sorted(runs)[len(runs) // 2]
    del g
    nbytes = 5624.0 * N
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "env_step_kernel (fused obs/disc-obs/reward/done/record)", "achieved": achieved,
            "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": _ncu_traffic("env_step_kernel_32768") if (N == 32768 and motions == "walk") else None, "avg_launch_ms": ms,
            "algorithmic_bytes_per_launch": nbytes, "envs": N, "step_table_mb": agent._add_motion.motion_lib.step_table.numel() * 4 / 1e6,
            "note": "one launch at 4096 envs (23 MB) is a single wave shorter than the DRAM pipeline fill: the fraction is "
                    "meaningful at 32768 envs (secondary configs[4])"}


def dominant_kernel_roofline(agent, args, tc_peak, peak_src):
    """Dominant kernel of the step = the dense-layer contraction (csrc/gemm_tc.cu); its most frequent big shape is the
    1024x1024 hidden layer over one minibatch (M = 4N rows).  Timed launch by launch with CUDA events on the launch
    stream, rotating over operand buffers larger than L2.  Returns (kernel alone with its operands' twins ready -> burst
    peak, the same layer including the conversion pass of its activation operand)."""
    import ctypes as C

    import torch
    from add_gym_b200 import _lib
    m = agent._model
    M = agent._mb_rows
    Kd = Nd = 1024
    nbuf = max(2, int(300e6 // (M * Kd * 4)) + 1)
    A = [torch.randn(M, Kd, device=m.flat.device) for _ in range(nbuf)]
    Cc = [torch.empty(M, Nd, device=m.flat.device) for _ in range(2)]
    Wt = torch.randn(Nd, Kd, device=m.flat.device) * 0.03
    bias = torch.zeros(Nd, device=m.flat.device)
    L = _lib.lib()

    bf16 = args.precision == "bf16"
    A16 = [a.to(torch.bfloat16) for a in A] if bf16 else None
    W16 = Wt.to(torch.bfloat16) if bf16 else None
    C16 = [torch.empty(M, Nd, device=m.flat.device, dtype=torch.bfloat16) for _ in range(2)] if bf16 else None

    h3 = args.precision == "f16x3"
    if h3:   # fp16 hi/lo planes + max|x| word per operand; filled by the first (untimed) calls
        A16 = [torch.zeros(2 * M * Kd, device=m.flat.device, dtype=torch.float16) for _ in range(nbuf)]
        W16 = torch.zeros(2 * Nd * Kd, device=m.flat.device, dtype=torch.float16)
        slots = torch.zeros(2 * (nbuf + 1), device=m.flat.device, dtype=torch.int32)     # {sticky scale word, max|x|} pairs

    def launch(i, ready_a, ready_b=1):
        a = _lib.AddkGemmArgs(A=A[i % nbuf].data_ptr(), lda=Kd, B=Wt.data_ptr(), ldb=Kd, C=Cc[i % 2].data_ptr(), ldc=Nd,
                              M=M, N=Nd, K=Kd, bias=bias.data_ptr(), a_mean=None, a_std=None, relu_mask_src=None,
                              ld_mask=0, trans_a=0, trans_b=1, relu=1, split_k=1, accumulate=0, slab_stride=0,
                              A16=A16[i % nbuf].data_ptr() if (bf16 or h3) else None, B16=W16.data_ptr() if (bf16 or h3) else None,
                              C16=C16[i % 2].data_ptr() if bf16 else None)
        if bf16:      # as in the update: a hidden layer's output is read by dense layers / as a mask only -> 16-bit only
            a.no_f32 = 1
        if h3:
            a.a16_plane, a.b16_plane = M * Kd, Nd * Kd
            a.a_amax, a.b_amax = slots[2 * (1 + i % nbuf):].data_ptr(), slots.data_ptr()
            a.a16_ready, a.b16_ready = ready_a, ready_b
        _lib.check(L.addk_gemm(_lib.stream(), C.byref(a), C.c_int(m.precision)), "addk_gemm")

    def timed_launches(ready_a):
        reps = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i, (a, b) in enumerate(evs):
            a.record()
            launch(i, ready_a)
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / reps

    for i in range(max(3, nbuf if h3 else 0)):
        launch(i, 0, 0)
    torch.cuda.synchronize()
    ms = timed_launches(1)
    flops = 2.0 * M * Nd * Kd
    achieved = flops / (ms * 1e-3) / 1e12
    traffic = _ncu_traffic("dense_layer_16384x1024x1024_%s" % args.precision) if M == 16384 else None
    passes = {"f16x3": 3, "tf32x3": 3}.get(args.precision, 1)
    out = {"bound": "tensor", "kernel": "dense layer %dx%dx%d (%s), operand twins ready" % (M, Nd, Kd, args.precision), "achieved": achieved,
           "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak, "traffic": traffic,
           "peak_source": peak_src + ", burst bf16 (kernel timed alone)", "avg_launch_ms": ms, "algorithmic_flops_per_launch": flops}
    if passes > 1:   # fp32-parity modes issue 3 tensor-core products per algorithmic one (hi.hi + lo.hi + hi.lo)
        out["mma_passes"] = passes
        out["tensor_pipe_tflops"] = passes * achieved
        out["note"] = ("achieved / frac count the ALGORITHMIC 2MNK flops of the fp32 layer; the tensor pipe executes %d MMA "
                       "passes per layer (overhead of the fp32-parity scheme, not algorithmic work): %.0f TFLOP/s of kind::f16 "
                       "work; the ceiling of this scheme is 1/3 of the tensor peak" % (passes, passes * achieved))
    pre = None
    if h3:   # the same layer when its activation operand still has to be converted (max pass + split pass in the timed region)
        ms2 = timed_launches(0)
        a2 = flops / (ms2 * 1e-3) / 1e12
        pre = {"bound": "tensor", "kernel": "dense layer %dx%dx%d (f16x3) incl. the max|x| + hi/lo split pre-pass of the activation "
                                            "operand (3 launches)" % (M, Nd, Kd), "achieved": a2, "peak": tc_peak, "unit": "TFLOP/s",
               "frac": a2 / tc_peak, "traffic": None, "avg_launch_ms": ms2, "algorithmic_flops_per_launch": flops,
               "peak_source": peak_src + ", burst bf16"}
    return out, pre


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
