"""Host-helper probe (run on the GPU box): jss_host_masked_random and jss_host_expand_obs on a 65 536-env ta80 batch,
pool pinned / unpinned, plus PCIe D2H time of the packed rows -- explains where an e2e step spends its host time."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from jssenv_b200 import JssVecEnv

n = 65536
bound = JssVecEnv.host_configure(0, 0 if os.environ.get("PROBE_BIND", "1") == "1" else -1)
env = JssVecEnv(n, {"instance_path": "ta80"}, auto_reset=True, seed=1)
env.reset(); acts = env.policy("RANDOM").clone()
for _ in range(500):
    *_, acts = env.step_sample(acts, "RANDOM")
torch.cuda.synchronize()
L, h = env._L, env._h
out = {"threads": int(L.jss_host_threads()), "bound_cpus": bound, "pin": os.environ.get("JSS_HOST_PIN", "1")}
mask = np.ascontiguousarray(env.action_mask.cpu().numpy())

def t(fn, reps=20):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3

out["masked_random_ms"] = t(lambda: env.host_masked_random(mask, 3))
env.host_step_begin(env.host_masked_random(mask, 0), packed=True)
env.host_wait_mask(); env.host_wait_obs()
b = env._pipe[env._pipe_slot]
wire, sc, obs = b["wire"], b["scalars"], b["obs"]
out["expand_ms"] = t(lambda: L.jss_host_expand_obs(h, ctypes.c_void_p(wire.data_ptr()), ctypes.c_void_p(sc.data_ptr()), ctypes.c_void_p(obs.data_ptr())))
for lvl, name in ((1, "expand_avx2_ms"), (0, "expand_scalar_ms")):
    L.jss_host_set_simd(lvl)
    out[name] = t(lambda: L.jss_host_expand_obs(h, ctypes.c_void_p(wire.data_ptr()), ctypes.c_void_p(sc.data_ptr()), ctypes.c_void_p(obs.data_ptr())), 5)
L.jss_host_set_simd(2)
# expansion while the copy engine streams into host memory (what happens inside an e2e step)
big_dev = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
big_host = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
side = torch.cuda.Stream()
def expand_under_dma():
    with torch.cuda.stream(side):
        for _ in range(2):
            big_host.copy_(big_dev, non_blocking=True)
    L.jss_host_expand_obs(h, ctypes.c_void_p(wire.data_ptr()), ctypes.c_void_p(sc.data_ptr()), ctypes.c_void_p(obs.data_ptr()))
out["expand_under_dma_ms"] = t(expand_under_dma, 10)
torch.cuda.synchronize()
dev = torch.empty(wire.shape, dtype=torch.uint8, device="cuda")
def d2h():
    wire.copy_(dev, non_blocking=True); torch.cuda.synchronize()
out["wire_d2h_ms"] = t(d2h)
out["wire_MB"] = wire.numel() / 1e6
a_np = env.host_masked_random(mask, 5)
def begin_only():
    env.host_step_begin(a_np, packed=True); env.host_wait_obs()
out["begin_plus_wait_all_ms"] = t(begin_only, 10)
print(json.dumps(out))
