"""Batched counterpart of the reference's examples/dispatching_rules_example.py: run every dispatching
rule on a whole batch of environments at once (fused rollouts on the GPU) and print makespan statistics.

    python examples/compare_rules_batched.py [instance] [num_envs]
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from jssenv_b200 import JssVecEnv  # noqa: E402
from jssenv_b200.dispatching import DISPATCHING_RULES  # noqa: E402


def main():
    instance = sys.argv[1] if len(sys.argv) > 1 else "ta80"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    env = JssVecEnv(n, {"instance_path": instance}, device=0, seed=0)
    print(f"{instance}: {env.jobs} jobs x {env.machines} machines, {n} episodes per rule (10 % exploration no-ops)")
    print(f"{'rule':<6}{'mean makespan':>15}{'best':>8}{'worst':>8}{'episodes/s':>14}")
    for name, rule in DISPATCHING_RULES.items():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ret, mk = rule.run_batch(env)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{name:<6}{mk.double().mean().item():>15.1f}{int(mk.min()):>8}{int(mk.max()):>8}{n / dt:>14.0f}")


if __name__ == "__main__":
    main()
