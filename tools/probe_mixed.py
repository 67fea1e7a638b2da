"""cfg5 probe (GPU): fused step + FIFO rule on the mixed ta01..ta80 batch, N = 65 536, for the cost-model
parameters given in JSS_COST_A / JSS_COST_B (see jss_assign).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from jssenv_b200 import JssVecEnv
names = ["ta%02d" % (k + 1) for k in range(80)]
n = 65536
env = JssVecEnv(n, {"instance_paths": names, "env_to_instance": np.arange(n) % 80}, auto_reset=True, seed=2)
env.reset()
rule = os.environ.get("PROBE_RULE", "FIFO")
acts = env.policy(rule).clone()
for _ in range(600):
    *_, acts = env.step_sample(acts, rule)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 600
for _ in range(K):
    *_, acts = env.step_sample(acts, rule)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"cost_a": os.environ.get("JSS_COST_A"), "cost_b": os.environ.get("JSS_COST_B"), "rule": rule,
                  "us_per_step": ms * 1e3, "env_steps_per_s": n / ms * 1e3, "errors": env.stats()["envs_error"]}))
