"""Where does an e2e (host-buffer) step go?  Raw pinned D2H bandwidth, host policy, jss_step_host."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from jssenv_b200 import JssVecEnv
N = 65536
env = JssVecEnv(N, {"instance_path": "ta80"}, device=0, auto_reset=True, seed=1)
env.reset()
host = torch.empty((N, 100, 7), dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize()
for nm, src in (("obs 183MB", env.real_obs),):
    t0 = time.perf_counter()
    for _ in range(10):
        host.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"D2H {nm}: {dt*1e3:.2f} ms = {src.numel()*4/dt/1e9:.1f} GB/s")
mask = np.ascontiguousarray(env.action_mask.cpu().numpy())
obs, *_ = env.step_host(env.host_masked_random(mask, 0))
t0 = time.perf_counter()
for k in range(20):
    a = env.host_masked_random(obs["action_mask"], k)
dt = (time.perf_counter() - t0) / 20
print(f"host policy: {dt*1e3:.2f} ms")
a = env.host_masked_random(obs["action_mask"], 99)
t0 = time.perf_counter()
for k in range(20):
    env.step_host(a * 0 - 1)          # SKIP actions: isolates copies + launch
dt = (time.perf_counter() - t0) / 20
print(f"jss_step_host (skip actions): {dt*1e3:.2f} ms")
t0 = time.perf_counter()
for k in range(20):
    env.step_host(a * 0 - 1, want_obs=False)
dt = (time.perf_counter() - t0) / 20
print(f"jss_step_host without obs copy: {dt*1e3:.2f} ms")
print("usable cpus", len(os.sched_getaffinity(0)), open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "")
