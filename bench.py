#!/usr/bin/env python
"""bench.py -- env steps/sec of the batched job-shop environment on B200.

One "step" = one pass of the hot path over the whole batch: the device masked-uniform
policy kernel picks an action per env, the fused step kernel applies it (time
advance, legal-action heuristics, observation, mask, reward, done all written to
HBM).  Workload = BASELINE.json configs[2]: N = 65 536 concurrent ta80 (100x20)
envs per GPU, auto-reset, weak scaling across GPUs (no per-step communication; one
NCCL all-gather of the per-shard episode statistics at the end).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "env steps/sec (N=65536 ta80)"
UNIT = "env_steps/s"


def b_alg(J, M):
    """Algorithmic bytes per env-step."""
    return 72 * J + 8 * M + 27


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def usable_cpus():
    """Host parallelism actually available to this process: min(online CPUs, affinity mask, cgroup quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(instance, budget_s, threads=None):
    """The reference's CPU path restated in C (oracle/jss_oracle.c, kind "port": the reference is
    pure Python and cannot travel to the GPU box), masked-uniform policy, auto-reset, one env per
    host thread, all host threads.  Returns (steps/s, threads, sample description)."""
    import numpy as np
    from jssenv_b200.instances import load_instance
    from oracle.jss_oracle import OracleEnv
    P = threads or usable_cpus()
    m, d = load_instance(instance)
    envs = [OracleEnv(m, d) for _ in range(P)]
    envs[0].run_random(1, 0, 20000)                   # warm-up
    res = [None] * P

    def work(i):
        res[i] = envs[i].run_random_timed(1, i, budget_s)   # ctypes releases the GIL; bounded by wall time

    ths = [threading.Thread(target=work, args=(i,)) for i in range(P)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    return steps / dt, P, (f"{P} host threads (os.cpu_count()={os.cpu_count()}, usable={usable_cpus()}) x {budget_s:.0f} s of "
                           f"{instance} steps (masked-random, auto-reset): {steps} steps")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (C oracle port, the Python reference cannot travel to the
    GPU box) on all usable host threads; each "step" is one bounded wall-time window of the workload."""
    if rank != 0:
        return
    J, M = 100, 20
    windows = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    vals, sample, P = [], "", 1
    for k in range(warm + windows):
        v, P, sample = cpu_reference_leg("ta80", 2.0 if k < warm else args.cpu_seconds)
        if k >= warm:
            vals.append(v)
    value = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "ta80 (100x20) masked-random, auto-reset, CPU oracle port of the reference step()",
                   "envs": P, "bytes_per_env_step": b_alg(J, M), "timed_windows": windows,
                   "window_seconds": args.cpu_seconds,
                   "python_reference_steps_per_s_per_core": 4378},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--instance", default="ta80")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from jssenv_b200 import JssVecEnv
    from jssenv_b200.distributed import all_gather_stats

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective; keep stdout = ONE JSON line
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    N = args.envs
    env = JssVecEnv(N, {"instance_path": args.instance}, device=local_rank, auto_reset=True,
                    env_id_base=rank * N, seed=1234)
    J, M = env.jobs, env.machines
    K, W = args.steps, max(3, args.warmup)

    env.reset()
    acts = env.policy("RANDOM").clone()
    for _ in range(W):
        *_, acts = env.step_sample(acts, "RANDOM")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for k in range(K):
        *_, acts = env.step_sample(acts, "RANDOM")     # one launch: apply actions, sample the next ones
    ev1.record()
    torch.cuda.synchronize()
    launches = env.launch_count - l0
    elapsed_ms = ev0.elapsed_time(ev1)
    # the timed region is K launches of ONE kernel (the fused step), bracketed by CUDA events on the
    # launching stream: its average launch duration is elapsed / K (launch gaps, if any, count against us)
    step_kernel_ms = elapsed_ms / K
    clocks = sampler.stop()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    stats = all_gather_stats(env.stats())
    value = world * N * K / (elapsed_ms * 1e-3)

    # ---- e2e: the same transitions through the host-buffer API (H2D actions, D2H obs/mask/reward/done)
    e2e = None
    if not args.no_e2e:
        # pipelined host-buffer API: begin(step k) -> wait mask k -> host policy -> begin(step k+1) while the
        # 2.9 KB/env observation of step k is still crossing PCIe into its own pinned buffer -> consume obs k
        env.reset()
        mask = np.ascontiguousarray(env.action_mask.cpu().numpy())
        env.host_step_begin(env.host_masked_random(mask, 0))
        checksum = 0.0

        def e2e_step(k):
            m, rew, dn = env.host_wait_mask()
            env.host_step_begin(env.host_masked_random(m, k))
            return env.host_wait_obs(previous=True)      # the step's observation, on the host

        for k in range(1, 4):
            e2e_step(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            obs_host = e2e_step(10 + k)
            checksum += float(obs_host[0, 0, 0])          # touch the result
        env.host_wait_obs()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * args.e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": 4 * N,
               "d2h_bytes_per_step": N * int(env._b.mask_stride) + N * J * 7 * 4 + 16 * N,
               "steps": args.e2e_steps,
               "note": "jss_host_step_begin / jss_host_wait (pinned host buffers): H2D actions, step kernel, D2H mask + "
                       "scalars + real_obs every step; host masked-random policy from the host mask; PCIe-bound"}

    if rank == 0:
        peak, peak_src = load_peaks()
        balg = b_alg(J, M)
        achieved = balg * N / (step_kernel_ms * 1e-3) / 1e9
        traffic = None                                   # dram__bytes_read+write per launch from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj["n_envs"] == N and tj["instance"] == args.instance:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"{args.instance} ({J}x{M}) N={N} per GPU, one fused launch per step (apply actions + masked-random "
                                   "sampling of the next ones), auto-reset", "envs_per_gpu": N, "parallelism": f"env-shard x{world}",
                       "l2": "per-step working set (state 140 MB + obs 190 MB at N=65536) exceeds the 126 MB L2",
                       "bytes_per_env_step": balg},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": balg * N, "kernel": "jss_step_kernel<4, sample>", "kernel_ms": step_kernel_ms,
                         "peak_source": peak_src},
            "episode_stats": stats,
        }
        if e2e:
            out["e2e"] = e2e
        if not args.no_cpu and world == 1:               # the CPU baseline is reported at N = 1 only
            v, P, sample = cpu_reference_leg(args.instance, args.cpu_seconds)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": P, "kind": "port", "sample": sample,
                                   "python_reference_steps_per_s_per_core": 4378}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
