"""Per-shape cost probe (GPU): fused step+sample time per 65 536-env launch for one instance of every bundled
shape (feeds the cost model of the mixed-batch scheduler), and the K-step fused rollout on ta01 N=4096 (cfg2)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from jssenv_b200 import JssVecEnv

def timed(fn, k, warm=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

out = {}
n = int(os.environ.get("PROBE_N", 65536))
names = os.environ.get("PROBE_NAMES", "ta01,ta11,ta21,ta31,ta41,ta51,ta61,ta71").split(",")
for name in names:
    env = JssVecEnv(n, {"instance_path": name}, auto_reset=True, seed=1)
    env.reset(); acts = env.policy("RANDOM").clone()
    for _ in range(int(0.4 * env.jobs * env.machines)):          # into the middle of the episode
        *_, acts = env.step_sample(acts, "RANDOM")
    def f():
        global acts
        *_, acts = env.step_sample(acts, "RANDOM")
    ms = timed(f, 400)
    J, M = env.jobs, env.machines
    out[f"{name}_{J}x{M}"] = {"us_per_launch": ms * 1e3, "ns_per_env_step": ms * 1e6 / n, "B_alg": 72 * J + 8 * M + 27,
                             "GBps": (72 * J + 8 * M + 27) * n / ms / 1e6}
    env.close(); del env
if os.environ.get("PROBE_NAMES"):
    print(json.dumps(out, indent=1)); sys.exit(0)
# cfg2: fused K-step rollout with observations written every step
env = JssVecEnv(4096, {"instance_path": "ta01"}, auto_reset=True, seed=1)
env.reset()
for K in (64, 512):
    env.rollout("RANDOM", K, write_obs=True); torch.cuda.synchronize()
    t0 = time.perf_counter(); env.rollout("RANDOM", K, write_obs=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[f"cfg2_rollout{K}_obs"] = {"us_per_step": dt / K * 1e6, "env_steps_per_s": 4096 * K / dt}
acts = env.policy("RANDOM").clone()
def f2():
    global acts
    *_, acts = env.step_sample(acts, "RANDOM")
ms = timed(f2, 2000)
out["cfg2_step_sample"] = {"us_per_step": ms * 1e3, "env_steps_per_s": 4096 / ms * 1e3}
print(json.dumps(out, indent=1))
