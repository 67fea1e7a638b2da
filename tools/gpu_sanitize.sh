#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck on a small mixed batch (all KJ classes, all kernels)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from jssenv_b200 import JssVecEnv
names = ["ta01", "ta31", "ta51", "ta80"]
env = JssVecEnv(64, {"instance_paths": names, "env_to_instance": np.arange(64) % 4}, seed=5, auto_reset=True, record_solution=True)
env.reset()
acts = env.policy("RANDOM").clone()
for k in range(300):
    *_, acts = env.step_sample(acts, "RANDOM")
for rule in ("SPT", "FIFO", "MWR", "CR"):
    for k in range(50):
        env.step(env.policy(rule))
env.rollout("LWR", 100)
snap = {k: v.clone() for k, v in env.export_state().items()}
env.import_state(snap)
mask = np.ascontiguousarray(env.action_mask.cpu().numpy())
env.host_step_begin(env.host_masked_random(mask, 0)); env.host_wait_mask(); env.host_wait_obs()
env.step_host(env.host_masked_random(mask, 1))
print(env.stats())
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
done
