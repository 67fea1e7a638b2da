"""Secondary throughput numbers for the other BASELINE.json configs (not bench lines; recorded in profiles/):
  cfg2: ta01 N=4096 masked-random (fused step+sample launches)
  cfg5: mixed ta01..ta80 N=65536, on-device FIFO / MWR (fused rollouts and fused step+sample)
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from jssenv_b200 import JssVecEnv

def timed(fn, k, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

out = {}
env = JssVecEnv(4096, {"instance_path": "ta01"}, auto_reset=True, seed=1)
env.reset(); acts = env.policy("RANDOM").clone()
def s2():
    global acts
    *_, acts = env.step_sample(acts, "RANDOM")
ms = timed(s2, 3000)
out["cfg2_ta01_N4096_random"] = {"ms_per_step": ms, "env_steps_per_s": 4096 / ms * 1e3, "bytes_per_env_step": 1227,
                                 "algorithmic_GBps": 1227 * 4096 / ms / 1e6}
del env
names = ["ta%02d" % (k + 1) for k in range(80)]
n = 65536
env = JssVecEnv(n, {"instance_paths": names, "env_to_instance": np.arange(n) % 80}, auto_reset=True, seed=2)
for rule in ("RANDOM", "FIFO", "MWR"):
    env.reset(); acts = env.policy(rule).clone()
    def s5():
        global acts
        *_, acts = env.step_sample(acts, rule)
    ms = timed(s5, 1500)
    out[f"cfg5_mixed_ta01-80_N65536_{rule}_step_sample"] = {"ms_per_step": ms, "env_steps_per_s": n / ms * 1e3}
    env.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    env.rollout(rule, 500, write_obs=False); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[f"cfg5_mixed_ta01-80_N65536_{rule}_rollout500_noobs"] = {"ms_per_step": dt / 500 * 1e3, "env_steps_per_s": n * 500 / dt}
print(json.dumps(out, indent=1))
