#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck on small batches: mixed lane classes (one persistent launch),
# uniform batch, tiny uniform batch (J <= 4: the no-op horizon table must fit the per-warp scratch), every kernel family
mkdir -p gpurun_out
T=${TAG:-r02}
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from jssenv_b200 import JssVecEnv, JssEnv
names = ["ta01", "ta31", "ta51", "ta80"]
env = JssVecEnv(64, {"instance_paths": names, "env_to_instance": np.arange(64) % 4}, seed=5, auto_reset=True, record_solution=True)
env.reset()
acts = env.policy("RANDOM").clone()
for k in range(300):
    *_, acts = env.step_sample(acts, "RANDOM")
for rule in ("SPT", "FIFO", "MWR", "CR"):
    for k in range(50):
        env.step(env.policy(rule))
    acts = env.policy(rule).clone()                     # fresh actions for the current state
    for k in range(20):
        *_, acts = env.step_sample(acts, rule)
env.rollout("LWR", 100)
tr = env.rollout_record("FIFO", 40)
snap = {k: v.clone() for k, v in env.export_state().items()}
env.import_state(snap)
mask = np.ascontiguousarray(env.action_mask.cpu().numpy())
env.host_step_begin(env.host_masked_random(mask, 0), packed=True)
mask, _, _ = env.host_wait_mask()                       # the mask AFTER the step: the next actions must be legal for it
env.host_step_begin(env.host_masked_random(mask, 1), packed=False)
env.host_wait_obs(previous=True)
mask, _, _ = env.host_wait_mask()
env.host_wait_obs()
env.step_host(env.host_masked_random(mask, 2))
st = env.stats(); print(st); assert st["envs_error"] == 0, st
# uniform batch (kernel variant with the instance scalars in the kernel parameters) and tiny uniform batch
rng = np.random.default_rng(0)
big = (np.stack([rng.permutation(6) for _ in range(200)]).astype(np.int32), rng.integers(1, 40, size=(200, 6)).astype(np.int32))
for inst, n in (("ta80", 24), ((np.array([[0, 1], [1, 0], [0, 1]], np.int32), np.array([[3, 2], [2, 4], [1, 1]], np.int32)), 16),
                (big, 10)):       # 200 jobs: the 8-jobs-per-lane class
    e = JssVecEnv(n, {"instance_path": inst}, seed=3, auto_reset=True)
    e.reset(); a = e.policy("RANDOM").clone()
    for k in range(150):
        *_, a = e.step_sample(a, "RANDOM")
    e.rollout("SPT", 60); e.rollout_record("RANDOM", 30)
    st = e.stats(); print(st); assert st["envs_error"] == 0, st
# single-env facade: host-mapped outputs, fused step + decode launch
f = JssEnv({"instance_path": "ta01"}); obs = f.reset(); done = False
while not done:
    obs, r, done, _, _ = f.step(int(np.flatnonzero(obs["action_mask"])[0]))
print("facade makespan", f.current_time_step)
PY
for tool in ${SAN_ONLY:-memcheck racecheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool python /tmp/san.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|envs_error|facade" gpurun_out/${T}_sanitizer_$tool.log | tail -6
done
