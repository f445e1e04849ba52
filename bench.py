#!/usr/bin/env python
"""Benchmark of the rollout-and-update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one training epoch of the reference's loop (pipelines/pipeline.py:163-164):
`buffer.sample()` (policy-in-the-loop rollout of every env for the full horizon)
followed by `algorithm.learn(buffer)` (RTG + group-relative advantages +
`updates_per_iter` clipped-surrogate updates with Adam).  The default workload is
BASELINE.json configs[1]: Pendulum GRPO, 65,536 envs x 200 steps, group size 16,
MLP 3-64-64-1, per GPU (weak scaling: every rank runs the full per-GPU shape with
whole groups; the only collective is the NCCL gradient allreduce).

`value`  : valid env-steps/s of the whole job with the initial states already in HBM.
`e2e`    : the same through the reference-facing host API (RolloutManager /
           Rollout_Buffer / GRPO) with HOST initial states copied H2D from pinned
           memory every step and the episode lengths + mean return read back D2H.
Extra keys: rollout-only env-steps/s, GRPO updates/s, roofline of the dominant
kernel, the CPU baseline measured on this box, clocks under load.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: env kind, class, T, group size E, groups per GPU, hidden, cov, gamma, eps, lr, updates, restart
    "pendulum": dict(kind=1, cls="Pendulum", T=200, E=16, G=4096, hidden=[64, 64], cov=0.5, gamma=0.99, eps=0.2,
                     lr=5e-4, updates=5, desc="Pendulum GRPO, 65,536 envs x 200-step horizon, group size 16, MLP 64x64"),
    "cartpole": dict(kind=0, cls="CartPole", T=500, E=10, G=10, hidden=[128, 128, 128, 128], cov=0.5, gamma=0.5,
                     eps=0.15, lr=3e-4, updates=1, desc="CartPole GRPO (scripts/cartpole_nn_grpo.py defaults)"),
    # The quadrotor envs end an episode when the vehicle leaves its box; a freshly initialised policy
    # with the reference's exploration noise (cov 0.3-0.5 => +-55-70 % thrust jitter) crashes within a
    # few dozen steps, so throughput would measure the zero-fill path.  These two workloads therefore use
    # a hover-biased start (output layer zeroed: mean action = hover thrust) and cov 1e-4, which keeps
    # most envs alive for the horizon; `value` still counts VALID steps only.
    "quadpole2d": dict(kind=2, cls="QuadPole2D", T=500, E=16, G=16384, hidden=[128, 128], cov=1e-4, gamma=0.99,
                       eps=0.2, lr=2e-4, updates=2, hover_init=True,
                       desc="QuadPole2D GRPO, 262,144 envs x 500 steps, group 16, MLP 128x128, hover-biased init"),
    # BASELINE configs[3] at its full size when run on 8 GPUs: 8 x 524,288 = 4,194,304 envs x 1000 steps, group 64
    # (55 GB of trajectory per GPU); `python -m torch.distributed.run --nproc-per-node 8 bench.py --gpus 8 --workload quadpole_cfg4`
    "quadpole_cfg4": dict(kind=3, cls="QuadPole", T=1000, E=64, G=8192, hidden=[256, 256], cov=1e-4, gamma=0.999, eps=0.2,
                          lr=3e-4, updates=1, hover_init=True,
                          desc="3D QuadPole GRPO, 524,288 envs per GPU x 1000 steps (4,194,304 envs on 8 GPUs), group 64, "
                               "MLP 256x256, hover-biased init"),
    # BASELINE configs[2] names PPO as well: the shipped quadpole2d_pipeline_ppo.py setting (full batch, GAE off =
    # Monte-Carlo returns, c1 0.5, kl 0.5) with 2 instead of 24 updates per epoch to keep the run short
    "quadpole2d_ppo": dict(kind=2, cls="QuadPole2D", T=500, E=16, G=16384, hidden=[128, 128], cov=1e-4, gamma=0.99,
                           eps=0.2, lr=2e-4, updates=2, hover_init=True, algo="ppo",
                           desc="QuadPole2D PPO (actor + critic 128x128, full batch), 262,144 envs x 500 steps, "
                                "hover-biased init"),
    "quadpole": dict(kind=3, cls="QuadPole", T=1000, E=64, G=1024, hidden=[256, 256], cov=1e-4, gamma=0.999, eps=0.2,
                     lr=3e-4, updates=1, hover_init=True,
                     desc="3D QuadPole GRPO, 65,536 envs x 1000 steps, group 64, MLP 256x256, hover-biased init"),
}
OBS = {0: 5, 1: 3, 2: 10, 3: 20}
ACT = {0: 1, 1: 1, 2: 2, 3: 4}


def mlp_macs(dims):
    return sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))


def k1_flops_of(P, valid_steps, ms):
    """achieved TFLOP/s of the rollout kernel: 2*P FLOP per valid env-step."""
    return 2.0 * P * valid_steps / (ms * 1e-3) / 1e12


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML
    (a polling `nvidia-smi -lms` subprocess was measured to add ~5 ms of launch stalls per step)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                             pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
                    except Exception:
                        pass
                    time.sleep(0.02)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.thread = None

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = [c for c, _ in self.samples]
        reasons = set()
        for _, mask in self.samples:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU arm: the reference's way (oracle/cpu_port.py), bounded sample of the same workload
# --------------------------------------------------------------------------------------
def cpu_leg(w, seed=0, budget_workers=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_port
    import restate as R
    import torch
    kind, T = w["kind"], w["T"]
    rng = np.random.default_rng(seed)
    dims = [OBS[kind]] + w["hidden"] + [ACT[kind]]
    torch.manual_seed(seed)
    Ws, bs = [], []
    for i in range(len(dims) - 1):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        Ws.append(lin.weight.detach().numpy().copy()); bs.append(lin.bias.detach().numpy().copy())
    cores = os.cpu_count() or 1
    G = budget_workers or cores                  # one group per host core
    E = max(2, min(w["E"], int(3000 // T) or 2))  # ~3k env-steps per worker: a 10-30 s sample
    cov = [w["cov"]] * ACT[kind]
    (obs, act, rew, lens, mask), t_roll, procs = cpu_port.rollout_mp(kind, T, R.DEFAULT_DT[kind], Ws, bs, cov, G, E,
                                                                     True, seed)
    torch.set_num_threads(cores)
    t_learn = cpu_port.grpo_learn(obs, act, rew, mask, Ws, bs, cov, w["gamma"], w["eps"], w["updates"], w["lr"])
    steps = int(lens.sum())
    return {
        "value": steps / (t_roll + t_learn), "unit": "env-steps/s", "cores": procs, "kind": "port",
        "sample": f"{G} groups x {E} episodes x <={T} steps ({steps} valid env-steps) rollout on {procs} worker "
                  f"processes (OMP_NUM_THREADS=1) + GRPO.learn x{w['updates']} updates on torch-CPU ({cores} threads)",
        "rollout_env_steps_per_s": steps / t_roll, "rollout_s": t_roll, "learn_s": t_learn,
        "updates_per_s": w["updates"] / t_learn,
    }


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host thread (torch reads
    # the variable when it is first imported, which happens inside cpu_leg)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(k, None)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_leg(w, seed=i)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    last["value"] = v
    line = {
        "impl": "reference", "metric": "policy-in-loop env-steps/sec (rollout + GRPO update epoch)", "value": v,
        "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": w["desc"], "cpu_sample": last["sample"]},
        "cpu_baseline": last, "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import trajopt_grpo_b200 as tg
    from trajopt_grpo_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    kind, T, E, G = w["kind"], w["T"], w["E"], w["G"]
    N = G * E
    O, A = OBS[kind], ACT[kind]
    dims = [O] + w["hidden"] + [A]
    P = mlp_macs(dims)

    torch.manual_seed(1234)                      # identical initial weights on every rank
    ppo = w.get("algo") == "ppo"
    policy = (tg.GaussianActorCritic_NeuralNetwork if ppo else tg.GaussianActor_NeuralNetwork)(O, A, w["hidden"], "ReLU",
                                                                                              w["cov"])
    if w.get("hover_init"):
        with torch.no_grad():
            last = policy.actor.network[-1]
            last.weight.zero_(); last.bias.zero_()
    opt = torch.optim.Adam(policy.parameters(), lr=w["lr"] * (1e-3 if w.get("hover_init") else 1.0))
    if ppo:
        algo = tg.PPO(w["eps"], policy, opt, None, w["updates"], c1=0.5, kl_coeff=0.5, gamma=w["gamma"], lam=0.95,
                      entropy=0.01, batch_size=None, monte_carlo=True)
    else:
        algo = tg.GRPO(w["eps"], 0.01, w["gamma"], policy, opt, None, updates_per_iter=w["updates"])
    env_cls = getattr(tg, w["cls"])
    mgr = tg.RolloutManager(lambda: env_cls(max_steps=T), policy, restart=True, num_workers=G * world,
                            num_episodes_per_worker=E, use_multiprocessing=False, seed=7, rank=rank, world_size=world)
    buf = tg.Rollout_Buffer(mgr)
    env = mgr.env
    rng = np.random.default_rng(100 + rank)
    total = args.warmup + args.steps

    def host_init():
        s0 = np.repeat(env.sample_initial_states(G, rng), E, axis=0)
        return torch.from_numpy(np.ascontiguousarray(s0.T)).to(torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    # ---------------- device-resident arm: inputs already in HBM ----------------
    inits = [host_init().to(dev) for _ in range(total)]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase = {"rollout": [], "adv": [], "update": []}
    lens_sum = torch.zeros((), dtype=torch.int64, device=dev)
    host_marks = []

    def one_step_device(i, timed):
        e = [ev() for _ in range(4)]
        e[0].record()
        r = mgr.rollout_device(init_state=inits[i])
        e[1].record()
        buf.device_rollout = r
        # GRPO.learn, with event marks between its phases
        e[2].record()
        algo.learn(buf)
        e[3].record()
        # also during warm-up: the first call of a torch op loads its CUDA module lazily, which stalls the
        # host for tens of ms -- that must not land inside the timed region
        lens_sum.add_(r.len.sum())
        if timed:
            host_marks.append(time.perf_counter())
            phase["rollout"].append((e[0], e[1]))
            phase["update"].append((e[2], e[3]))

    for i in range(args.warmup):
        one_step_device(i, False)
    lens_sum.zero_()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # NVML initialisation takes tens of ms on a cold driver: BEFORE the barrier, so that the
    barrier()                   # other ranks do not start their timed region (and wait in the first allreduce) meanwhile
    launches0 = engine.COUNTERS["launches"]
    t0, t1 = ev(), ev()
    host_t0 = time.perf_counter()
    t0.record()
    for i in range(args.warmup, total):
        one_step_device(i, True)
    t1.record()
    barrier()
    launches = engine.COUNTERS["launches"] - launches0
    ms_total = max_over_ranks(t0.elapsed_time(t1))
    valid_steps = sum_over_ranks(float(lens_sum.item()))
    value = valid_steps / (ms_total * 1e-3)
    roll_ms = float(np.mean([a.elapsed_time(b) for a, b in phase["rollout"]]))
    upd_ms = float(np.mean([a.elapsed_time(b) for a, b in phase["update"]]))
    valid_per_step_rank = float(lens_sum.item()) / args.steps

    # ---------------- dominant-kernel timing: K3 alone, CUDA events on its stream -------------
    r = buf.device_rollout
    adv, _ = engine.advantage(0, r.G, r.E, r.T, w["gamma"], 0.0, r.rew, r.len)
    flat = policy.actor.flat_params() if ppo else policy.flat_parameters()
    k3 = []
    for i in range(3 + 5):
        a, b = ev(), ev()
        a.record()
        engine.policy_grad(dims, "ReLU", flat, policy.cov_diag, r.obs, r.act, adv, r.logp, r.len, w["eps"], 1.0 / G)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            k3.append(a.elapsed_time(b))
    k3_ms = float(np.mean(k3))
    k1 = []
    for i in range(2 + 3):
        a, b = ev(), ev()
        a.record()
        mgr.rollout_device(init_state=inits[0])
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            k1.append(a.elapsed_time(b))
    k1_ms = float(np.mean(k1))
    k2 = []
    for i in range(2 + 3):
        a, b = ev(), ev()
        a.record()
        engine.advantage(0, r.G, r.E, r.T, w["gamma"], 0.0, r.rew, r.len)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            k2.append(a.elapsed_time(b))
    k2_ms = float(np.mean(k2))
    clock_info = clocks.stop() if rank == 0 else None

    # env-steps a rollout tile actually executes: a tile of 128 consecutive envs runs until its longest episode ends
    ln_t = r.len.to(torch.int64)
    pad_n = (-ln_t.numel()) % 128
    tile_max = torch.nn.functional.pad(ln_t, (0, pad_n)).view(-1, 128).max(dim=1).values
    k1_exec_steps = float(tile_max.sum().item()) * 128.0
    if args.device_only:                          # short run for ncu: no e2e / CPU legs
        if rank == 0 and os.environ.get("TG_TIMELINE"):
            print("gpu e0/e3 ms since t0:", [(round(t0.elapsed_time(a), 2), round(t0.elapsed_time(d), 2))
                                             for (a, _), (_, d) in zip(phase["rollout"], phase["update"])],
                  "end", round(t0.elapsed_time(t1), 2), file=sys.stderr)
            print("host enqueue done ms since t0:", [round((m - host_t0) * 1e3, 2) for m in host_marks], file=sys.stderr)
        if rank == 0:
            print(json.dumps({"device_only": True, "value": value, "ms_per_step": ms_total / args.steps,
                              "k1_ms": k1_ms, "k2_ms": k2_ms, "k3_ms": k3_ms, "gpu_launches": launches,
                              "valid_frac": valid_per_step_rank / (N * T), "k1_executed_frac": k1_exec_steps / (N * T),
                              "k1_executed_tflops": 2.0 * P * k1_exec_steps / (k1_ms * 1e-3) / 1e12,
                              "rollout_env_steps_per_s": valid_per_step_rank / (k1_ms * 1e-3),
                              "k1_tflops": k1_flops_of(P, valid_per_step_rank, k1_ms),
                              "k3_tflops": 6.0 * P * valid_per_step_rank / (k3_ms * 1e-3) / 1e12}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- end-to-end arm: host buffers in, host scalars out ----------------
    host_inits = [host_init() for _ in range(total)]
    h2d = host_inits[0].numel() * 4
    d2h = N * 4 + 4
    e2e_steps = torch.zeros((), dtype=torch.int64, device=dev)
    lens_host = torch.empty(N, dtype=torch.int32).pin_memory()

    def one_step_e2e(i):
        x = host_inits[i].to(dev, non_blocking=True)                 # H2D of this step's inputs
        r = mgr.rollout_device(init_state=x)
        buf.device_rollout = r
        algo.learn(buf)
        lens_host.copy_(r.len, non_blocking=True)                    # D2H of the step's results
        mean_ret = float(r.ret.mean().item())                        # (sync) what Rollout_Buffer.store reports
        buf.avg_reward.append(mean_ret)
        return int(lens_host.sum())

    for i in range(args.warmup):
        one_step_e2e(i)
    barrier()
    a, b = ev(), ev()
    wall0 = time.perf_counter()
    a.record()
    n_e2e = 0
    for i in range(args.warmup, total):
        n_e2e += one_step_e2e(i)
    b.record()
    barrier()
    wall = time.perf_counter() - wall0
    e2e_ms = max_over_ranks(max(a.elapsed_time(b), wall * 1e3))
    e2e_value = sum_over_ranks(float(n_e2e)) / (e2e_ms * 1e-3)

    # ---------------- roofline of the dominant kernel (K3: tg_policy_grad) ----------------
    fp32_peak = engine.fp32_peak_tflops(dev)
    # algorithmic FLOPs of one K3 launch = 6*P FLOP per valid step (fwd 2P + bwd 4P), SURVEY 8d
    k3_flops = 6.0 * P * valid_per_step_rank
    k1_flops = 2.0 * P * valid_per_step_rank
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    tc = w["hidden"] in ([64, 64], [128, 128], [256, 256])     # shapes with a tcgen05 update path (fused 64, streamed 128/256)
    k3_name = ("update_tc_kernel (tg_policy_grad: fused fwd + clipped surrogate + MLP backward, tcgen05 3xTF32)"
               if w["hidden"] == [64, 64] else
               "update_tcw_fwdbwd_kernel + update_tcw_wgrad_kernel (tg_policy_grad: streamed forward/backward + split-K "
               "weight gradients, tcgen05 3xTF32)")
    k3_tf = k3_flops / (k3_ms * 1e-3) / 1e12
    k1_tf = k1_flops / (k1_ms * 1e-3) / 1e12
    if tc:
        # K3/K1 hidden GEMMs run on the tensor cores as 3xTF32: every algorithmic MAC costs 3 tf32 MMA-MACs and
        # tf32 peaks at half the bf16 rate, so an fp32-faithful kernel tops out at bf16_peak / 6.
        roofline = {
            "kernel": k3_name,
            "bound": "tensor", "achieved": k3_tf, "peak": bf16_peak, "unit": "TFLOP/s", "frac": k3_tf / bf16_peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else
                           "fallback of /opt/skills/guides/B200_PROFILING.md (MEASURED_PEAKS.json absent)",
            "traffic": 322.1e6 if w is WORKLOADS["pendulum"] else None,
            "traffic_source": "profiles/r1f_ncu_details_pendulum_kernels.csv: dram read 315.4 MB + write 6.7 MB "
                              "per launch (algorithmic: 315 MB of obs+act+adv+old logp)",
            "achieved_note": "algorithmic FLOPs (6*P per valid step, SURVEY 8d) / CUDA-event time of the launch",
            "frac_of_3xtf32_ceiling": k3_tf / (bf16_peak / 6.0),
            "frac_of_measured_fp32_fma_peak": k3_tf / fp32_peak, "fp32_fma_peak_tflops": fp32_peak,
            "ms_per_launch": k3_ms, "flops_per_launch": k3_flops,
        }
    else:
        roofline = {
            "kernel": "update_kernel (tg_policy_grad: fused fwd + clipped surrogate + MLP backward, FP32 pipe)",
            "bound": "fp32-fma", "achieved": k3_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": k3_tf / fp32_peak,
            "traffic": None,
            "peak_source": "tg_fp32_peak FFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
            "frac_of_measured_bf16_tensor_peak": k3_tf / bf16_peak,
            "ms_per_launch": k3_ms, "flops_per_launch": k3_flops,
        }
    hid = w["hidden"]
    k1_tc = len(hid) >= 2 and len(set(hid)) == 1 and hid[0] in (64, 128) or hid == [256, 256]   # tg_rollout's routing
    k1_name = {64: "rollout_tc_kernel", 128: "rollout_tc2_kernel", 256: "rollout_tc256_kernel"}.get(hid[0]) if k1_tc \
        else "rollout_kernel"
    roofline["others"] = {
        k1_name: {
            "bound": "tensor" if k1_tc else "fp32-fma", "ms": k1_ms, "achieved_tflops": k1_tf,
            "executed_tflops": 2.0 * P * k1_exec_steps / (k1_ms * 1e-3) / 1e12,
            "frac": k1_tf / (bf16_peak if k1_tc else fp32_peak),
            "frac_of_measured_fp32_fma_peak": k1_tf / fp32_peak,
            "traj_write_gbs": 4.0 * (O + A + 2) * N * T / (k1_ms * 1e-3) / 1e9},
        "adv_grpo_kernel": {"bound": "hbm", "ms": k2_ms, "achieved_gbs": 8.0 * N * T / (k2_ms * 1e-3) / 1e9,
                            "peak_gbs": hbm_peak, "frac": 8.0 * N * T / (k2_ms * 1e-3) / 1e9 / hbm_peak},
    }
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_leg(w) if (world == 1 and not args.no_cpu) else None
    line = {
        "metric": "policy-in-loop env-steps/sec (rollout + GRPO update epoch)", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "envs_per_gpu": N, "horizon": T, "group_size": E, "mlp": dims,
                   "updates_per_iter": w["updates"], "precision": "fp32 state; hidden GEMMs 3xTF32 on tcgen05 (fp32-faithful), rest fp32" if tc else
                                "fp32 state + fp32 MLP (FP32 pipe)",
                   "l2": "per-step working set %.0f MB > 126 MB L2 (inputs larger than L2, no flush)" %
                         (4.0 * (O + A + 3) * N * T / 1e6),
                   "parallelism": f"dp{world} (whole GRPO groups per GPU, NCCL grad allreduce)"},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "rollout_env_steps_per_s": valid_per_step_rank * world / (roll_ms * 1e-3),
        "grpo_updates_per_s": w["updates"] / (upd_ms * 1e-3),
        "phase_ms": {"rollout": roll_ms, "learn": upd_ms,
                     "per_step_total": [round(a.elapsed_time(d), 3) for (a, _), (_, d) in
                                        zip(phase["rollout"], phase["update"])]},
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pendulum", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--device-only", action="store_true", help="device-resident arm only (profiling runs)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
