"""bench.py -- add-gym rollout + update hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs 4096] [--precision f16x3|tf32x3|fp32|tf32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU implementation (oracle port), host cores

One "step" = one training iteration of the reference agent (`BaseAgent._train_iter`, base_agent.py:353-374):
a 32-step rollout over N envs, `_build_train_data`, 5 epochs x 8 minibatches of the ADD/PPO update and the
normalizer update -- T*N = 131,072 env-steps at N = 4096.  Physics is excluded: the synthetic engine's
`scene.step()` is timed with its own CUDA events and subtracted.  metric = env-steps/s = T*N*R / t.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (BASELINE configs[1]: 4096)")
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "tf32x3", "tf32", "bf16", "f16x3"],
                    help="MLP arithmetic: f16x3 = tcgen05 kind::f16 on fp16 hi/lo planes (fp32-parity mode, default); tf32x3 = "
                         "kind::tf32 3-pass split (fp32 parity); fp32 = CUDA cores; tf32 = single pass; bf16 = config 4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-envs", type=int, default=512, help="envs of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stages-out", default=os.path.join(REPO, "gpurun_out", "bench_stages.json"))
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference, timed on the host cores (oracle/ may only be executed here as
# the baseline / checker, never on the product path)
# ---------------------------------------------------------------------------------------------------------------
def cpu_iteration_rate(envs, steps, warmup):
    import torch
    from add_gym_b200 import config as b200_config
    from oracle import harness
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = b200_config.default_config(num_envs=envs)
    agent = harness.make_oracle_agent(envs, seed=0, engine_seed=1234, cfg=cfg, fall_prob=0.002)
    agent.start()
    T = agent.T
    times = []
    for i in range(warmup + steps):
        phys = [0.0]
        scene = agent.env.scene
        orig = scene.step

        def timed_step(orig=orig, phys=phys):
            t0 = time.perf_counter()
            orig()
            phys[0] += time.perf_counter() - t0
        scene.step = timed_step
        t0 = time.perf_counter()
        agent.train_iter()
        dt = time.perf_counter() - t0 - phys[0]
        scene.step = orig
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return T * envs / mean, mean, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec, cores = cpu_iteration_rate(args.cpu_envs, args.steps, args.warmup)
    sample = "oracle port of the reference agent (torch CPU fp32, %d threads): one full iteration " \
             "(32-step rollout + build_train_data + 40 optimizer steps) at %d envs per step" % (cores, args.cpu_envs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "G1 walk1_subject1_trimmed, synthetic engine (physics excluded), bounded sample: "
                               "%d envs x 32 steps per iteration on the host CPU" % args.cpu_envs,
                   "envs_per_step": args.cpu_envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
def flops_per_iteration(N, T=32, epochs=5, batch=4):
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
    from add_gym_b200.engine import HostBoundaryEngine

    N = args.envs
    peaks = {}
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            peaks = json.load(f)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json, sustained bf16)" if peaks else "fallback"

    def make_agent(engine_cls=None):
        cfg = b200_config.default_config(num_envs=N, mlp_precision=args.precision)
        cfg["engine"].update(seed=1234 + rank, noise_device="device", fall_prob=0.002)
        if engine_cls is not None:
            cfg["engine"]["_target_"] = "add_gym_b200.engine." + engine_cls
        torch.manual_seed(0)
        a = ADDAgent(cfg, distributed=world > 1, device=dev)
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        a._reset_tracker()
        return a

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(agent, K, W, e2e):
        """returns (seconds total excl. physics, physics seconds, per-stage seconds, launches, last info)"""
        stage = {"rollout": 0.0, "build_train_data": 0.0, "update": 0.0, "normalizers": 0.0}
        ev = lambda: torch.cuda.Event(enable_timing=True)
        info_host = None
        for _ in range(max(W, 2)):      # at least two: the second rollout captures the CUDA graphs of the rollout step
            info = agent._train_iter()
            if e2e:
                info_host = {k: float(v) for k, v in info.items()}
        barrier()
        agent.engine_time_events = []
        _lib.launch_count(reset=True)
        marks = []
        t_begin, t_end = ev(), ev()
        t_begin.record()
        for _ in range(K):
            m = [ev() for _ in range(5)]
            m[0].record()
            agent.set_mode(agent._mode)
            agent._rollout_train(agent._steps_per_iter)
            m[1].record()
            data_info = agent._build_train_data()
            m[2].record()
            train_info = agent._update_model()
            m[3].record()
            if agent._need_normalizer_update():
                agent._update_normalizers()
            m[4].record()
            marks.append(m)
            if e2e:   # the user-visible result of the iteration, read back to the host
                info_host = {k: float(v) for k, v in {**train_info, **data_info}.items()}
                bad = [k for k, v in info_host.items() if v != v or v in (float("inf"), float("-inf"))]
                if bad:
                    raise SystemExit("bench.py: non-finite diagnostics %s -- the timed path is numerically broken" % bad)
        t_end.record()
        barrier()
        if not e2e:   # outside the timed region: the last iteration's diagnostics and the parameters must be finite
            last = {k: float(v) for k, v in {**train_info, **data_info}.items()}
            bad = [k for k, v in last.items() if v != v or v in (float("inf"), float("-inf"))]
            if bad or not bool(torch.isfinite(agent._model.flat).all()):
                raise SystemExit("bench.py: non-finite diagnostics %s / parameters -- the timed path is numerically broken" % bad)
        total = t_begin.elapsed_time(t_end) * 1e-3
        phys = sum(a.elapsed_time(b) for a, b in agent.engine_time_events) * 1e-3
        agent.engine_time_events = None
        for m in marks:
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

    K, W = args.steps, max(args.warmup, 0)
    T = 32
    # ---- device-resident arm -------------------------------------------------------------------------------
    agent = make_agent()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    sec, phys, stage, launches, _ = timed(agent, K, W, e2e=False)
    clk = clocks.stop() if rank == 0 else None
    sec = max_over_ranks(sec)
    value = T * N * world * K / sec
    stage_max = {k: max_over_ranks(v) for k, v in stage.items()}

    # ---- dominant kernel: the dense-layer contraction of the update, timed launch by launch --------------------
    roof = dominant_kernel_roofline(agent, args, hbm_peak, tc_peak, peak_src) if rank == 0 else None
    hbm_roof = step_kernel_roofline(agent, hbm_peak) if rank == 0 else None
    agent.release_graphs()
    del agent
    gc.collect()
    torch.cuda.empty_cache()

    # ---- end to end: simulator state arrives from pinned host memory every env step --------------------------
    e2e = None
    if not args.no_e2e:
        agent = make_agent("HostBoundaryEngine")
        ent = agent._env.robot.entity
        sec_e, phys_e, _, _, info_host = timed(agent, K, W, e2e=True)
        sec_e = max_over_ranks(sec_e)
        info_bytes = 8 * len(info_host or {})
        e2e = {"value": T * N * world * K / sec_e, "unit": UNIT,
               "h2d_bytes_per_step": int(ent.h2d_bytes_per_env_step * T), "d2h_bytes_per_step": int(ent.d2h_bytes_per_env_step * T + info_bytes),
               "what": "ADDAgent._train_iter with the simulator state copied from pinned host memory before every env "
                       "step, the action copied back to pinned host memory after every actor forward and the "
                       "iteration diagnostics read back as Python floats"}
        agent.release_graphs()
        del agent
        gc.collect()
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, s, cores = cpu_iteration_rate(args.cpu_envs, 4, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle port of the reference agent (torch CPU fp32, %d threads), 1 warm-up + 4 timed full "
                         "iterations at %d envs (%.1f s each)" % (cores, args.cpu_envs, s)}

    fl = flops_per_iteration(N)
    stages = {
        "envs_per_gpu": N, "n_gpus": world, "precision": args.precision, "seconds_per_iteration": sec / K,
        "physics_seconds_per_iteration_excluded": phys / K,
        "stages_seconds_per_iteration": {k: v / K for k, v in stage_max.items()},
        "rollout_env_steps_per_s": T * N * world * K / max(stage_max["rollout"], 1e-9),
        "ppo_update_samples_per_s": 5 * T * N * world * K / max(stage_max["update"], 1e-9),
        "tflops": {"rollout_actor": fl[0] * K / max(stage_max["rollout"], 1e-9) / 1e12,
                   "build_train_data": fl[1] * K / max(stage_max["build_train_data"], 1e-9) / 1e12,
                   "update": fl[2] * K / max(stage_max["update"], 1e-9) / 1e12},
        "launches_per_iteration": launches / K,
    }
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
        "config": {"workload": ("BASELINE configs[1]" if N == 4096 else "BASELINE configs[4] shape (physics-free synthetic-state benchmark)" if N == 32768
                                else "BASELINE configs[1] at another env count") +
                               ": G1 walk1_subject1_trimmed, %d envs/GPU, %s MLPs, one iteration = "
                               "32-step rollout + build_train_data + 5x8 ADD/PPO minibatches; synthetic engine stands in "
                               "for Genesis (not installed), scene.step() timed separately and excluded" % (N, args.precision),
                   "envs_per_gpu": N, "steps_per_iter": T, "minibatch": 4 * N, "optimizer_steps": 40,
                   "parallelism": "dp%d" % world,
                   "l2": "working set per iteration (experience buffers %.0f MB + activations) exceeds the 126 MB L2" %
                         (T * N * (264 * 2 + 114 * 2 + 29 + 8) * 4 / 1e6),
                   "ppo_update_samples_per_s": stages["ppo_update_samples_per_s"],
                   "physics_ms_per_step_excluded": phys / K * 1e3},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "roofline_hbm_stage": hbm_roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _ncu_traffic(key):
    """dram bytes read + written per launch from the committed `ncu --set full` capture (profiles/r01_ncu_metrics.json)."""
    try:
        with open(os.path.join(REPO, "profiles", "r01_ncu_metrics.json")) as f:
            m = json.load(f)[key]
        return int(m["dram_bytes_read"]) + int(m["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        return None


def step_kernel_roofline(agent, hbm_peak):
    """The HBM-bound stage: the fused per-env step kernel (csrc/step.cu), 5,624 algorithmic bytes per env-step
    (SURVEY 8d).  `reps` back-to-back launches through the C-ABI (structs prepared beforehand, successive experience
    rows) between one CUDA-event pair on the launch stream -> average launch duration."""
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        L.addk_env_step(_lib.stream(), C.byref(core.task), C.byref(core.c_lib), C.byref(sim), C.byref(core.c_env),
                        C.byref(rows[i]), _lib.ptr(core.dof_err_w), None, C.c_int(N), C.c_int(i % 3), C.c_int(flags))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 5624.0 * N
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "env_step_kernel (fused obs/disc-obs/reward/done/record)", "achieved": achieved,
            "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": _ncu_traffic("env_step_kernel_32768") if N == 32768 else None, "avg_launch_ms": ms,
            "algorithmic_bytes_per_launch": nbytes, "envs": N,
            "note": "not HBM-bound yet: a per-warp latency chain (ncu at 32768 envs: issue slots 64 % busy, 46 % of the "
                    "warp slots occupied, 127 MB of DRAM traffic for 184 MB algorithmic; profiles/r01_ncu_step_fast_"
                    "metrics.txt); one launch is shorter than the DRAM pipeline fill at 4096 envs -- the fraction is "
                    "meaningful at 32768 envs (config 5: 52 %)"}


def dominant_kernel_roofline(agent, args, hbm_peak, tc_peak, peak_src):
    """Dominant kernel of the step = the dense-layer contraction (csrc/gemm*.cu); its most frequent big shape is the
    1024x1024 hidden layer over one minibatch (M = 4N rows).  Timed launch by launch with CUDA events on the launch
    stream, rotating over operand buffers larger than L2."""
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
    ready = [0]
    if h3:   # fp16 hi/lo planes + max|x| word per operand; filled by the first (untimed) calls, then reused: the
        # timed launches are the dense-layer kernel itself, the split pre-pass is part of the step numbers
        A16 = [torch.zeros(2 * M * Kd, device=m.flat.device, dtype=torch.float16) for _ in range(nbuf)]
        W16 = torch.zeros(2 * Nd * Kd, device=m.flat.device, dtype=torch.float16)
        slots = torch.zeros(2 * (nbuf + 1), device=m.flat.device, dtype=torch.int32)     # {sticky scale word, max|x|} pairs

    def launch(i):
        a = _lib.AddkGemmArgs(A=A[i % nbuf].data_ptr(), lda=Kd, B=Wt.data_ptr(), ldb=Kd, C=Cc[i % 2].data_ptr(), ldc=Nd,
                              M=M, N=Nd, K=Kd, bias=bias.data_ptr(), a_mean=None, a_std=None, relu_mask_src=None,
                              ld_mask=0, trans_a=0, trans_b=1, relu=1, split_k=1, accumulate=0, slab_stride=0,
                              A16=A16[i % nbuf].data_ptr() if (bf16 or h3) else None, B16=W16.data_ptr() if (bf16 or h3) else None,
                              C16=C16[i % 2].data_ptr() if bf16 else None)
        if h3:
            a.a16_plane, a.b16_plane = M * Kd, Nd * Kd
            a.a_amax, a.b_amax = slots[2 * (1 + i % nbuf):].data_ptr(), slots.data_ptr()
            a.a16_ready = a.b16_ready = ready[0]
        _lib.check(L.addk_gemm(_lib.stream(), C.byref(a), C.c_int(m.precision)), "addk_gemm")
    for i in range(max(3, nbuf if h3 else 0)):
        launch(i)
    ready[0] = 1
    torch.cuda.synchronize()
    reps = 10
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (a, b) in enumerate(evs):
        a.record()
        launch(i)
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / reps
    flops = 2.0 * M * Nd * Kd
    achieved = flops / (ms * 1e-3) / 1e12
    traffic = _ncu_traffic("dense_layer_16384x1024x1024_%s" % args.precision) if M == 16384 else None
    passes = {"f16x3": 3, "tf32x3": 3}.get(args.precision, 1)
    out = {"bound": "tensor", "kernel": "dense layer %dx%dx%d (%s)" % (M, Nd, Kd, args.precision), "achieved": achieved,
           "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak, "traffic": traffic, "peak_source": peak_src,
           "avg_launch_ms": ms, "algorithmic_flops_per_launch": flops}
    if passes > 1:   # fp32-parity modes issue 3 tensor-core products per algorithmic one (hi.hi + lo.hi + hi.lo)
        out["mma_passes"] = passes
        out["tensor_pipe_tflops"] = passes * achieved
        out["note"] = ("achieved / frac count the ALGORITHMIC 2MNK flops of the fp32 layer; the tensor pipe executes %d MMA "
                       "passes per layer (%s), i.e. %.0f TFLOP/s of %s work = %.0f %% of the measured bf16 peak" %
                       (passes, "fp16 hi/lo planes" if args.precision == "f16x3" else "tf32 hi/lo split", passes * achieved,
                        "kind::f16" if args.precision == "f16x3" else "kind::tf32", 100.0 * passes * achieved / tc_peak))
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
