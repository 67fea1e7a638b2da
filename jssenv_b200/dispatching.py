"""Dispatching rules -- mirror of ``JSSEnv/dispatching.py`` (same class names, same
``DISPATCHING_RULES`` / ``get_rule`` / ``compare_rules`` API), evaluated on the GPU.

The arg-min / arg-max of every rule (SPT dispatching.py:92-116, FIFO 133-156, MWR
173-199, LWR 216-242, MOR 259-283, LOR 300-324, CR 365-408) is computed by the
``jss_policy`` kernel; like the reference, ties go to the lowest job index and the
"wait 10 % of the time" coin is drawn from ``np.random.random()`` ONLY when the
no-op is legal, so a seeded episode reproduces the reference's decisions exactly.
``run_batch`` / ``compare_rules_batched`` are the vectorised forms: whole batches of
episodes in one fused rollout launch, coin from the device counter RNG.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np


class DispatchingRule:
    """Base class for all dispatching rules (dispatching.py:21-75)."""

    def __init__(self, name: str, description: str):
        self.name = name
        self.description = description

    def _prepare(self, env) -> None:
        """Rule parameters that live in the native handle (only CR has one)."""

    def __call__(self, env) -> int:
        self._prepare(env)
        action, noop_legal = env.rule_action(self.name)
        if action == env.jobs:                 # only the no-op is legal (e.g. dispatching.py:96-97): no draw
            return action
        if noop_legal and np.random.random() < 0.1:   # e.g. dispatching.py:113-114
            return env.jobs
        return action

    def get_name(self) -> str:
        return self.name

    def get_description(self) -> str:
        return self.description

    def run_episode(self, env) -> Tuple[float, int]:
        """dispatching.py:55-75"""
        env.reset()
        done = False
        total_reward = 0.0
        while not done:
            action = self(env)
            _, reward, done, _, _ = env.step(action)
            total_reward += reward
        return total_reward, env.current_time_step

    def run_batch(self, vec_env, max_steps: Optional[int] = None, chunk: int = 256):
        """Whole-batch episodes on device: returns (raw returns int32[N], makespans int32[N]).
        Envs are reset first; each env stops at its own `done` (frozen afterwards)."""
        self._prepare(vec_env)
        vec_env.reset()
        if max_steps is None:
            # a transition either allocates an op (J*M of them) or advances time (at most one
            # event per allocated op), so 2*J*M+1 transitions always finish an episode
            max_steps = 2 * vec_env.jobs * vec_env.machines + 1
        done_steps = 0
        while done_steps < max_steps:
            n = min(chunk, max_steps - done_steps)
            vec_env.rollout(self.name, n, write_obs=False)
            done_steps += n
            if bool(vec_env.done.all()):
                break
        return vec_env.last_return.clone(), vec_env.last_makespan.clone()


def _mk(cls_name, name, description):
    def __init__(self):
        DispatchingRule.__init__(self, name, description)
    return type(cls_name, (DispatchingRule,), {"__init__": __init__, "__doc__": description})


ShortestProcessingTime = _mk("ShortestProcessingTime", "SPT",
                             "Shortest Processing Time: Schedule the job with the shortest processing time next")
FirstInFirstOut = _mk("FirstInFirstOut", "FIFO",
                      "First In First Out: Schedule the job that has been waiting the longest")
MostWorkRemaining = _mk("MostWorkRemaining", "MWR",
                        "Most Work Remaining: Schedule the job with the most processing time remaining")
LeastWorkRemaining = _mk("LeastWorkRemaining", "LWR",
                         "Least Work Remaining: Schedule the job with the least processing time remaining")
MostOperationsRemaining = _mk("MostOperationsRemaining", "MOR",
                              "Most Operations Remaining: Schedule the job with the most operations remaining")
LeastOperationsRemaining = _mk("LeastOperationsRemaining", "LOR",
                               "Least Operations Remaining: Schedule the job with the fewest operations remaining")


class CriticalRatio(DispatchingRule):
    """dispatching.py:327-408; ``due_date_factor`` (:337-349) is a launch parameter of the device rule."""

    def __init__(self, due_date_factor: float = 1.5):
        super().__init__("CR", "Critical Ratio: Schedule based on the ratio of time to due date versus remaining work")
        self.due_date_factor = float(due_date_factor)

    def _prepare(self, env) -> None:
        env.set_cr_due_date_factor(self.due_date_factor)


DISPATCHING_RULES = {
    "SPT": ShortestProcessingTime(), "FIFO": FirstInFirstOut(), "MWR": MostWorkRemaining(),
    "LWR": LeastWorkRemaining(), "MOR": MostOperationsRemaining(), "LOR": LeastOperationsRemaining(),
    "CR": CriticalRatio(),
}


def get_rule(rule_name: str) -> DispatchingRule:
    """dispatching.py:423-439"""
    if rule_name not in DISPATCHING_RULES:
        raise ValueError(f"Rule '{rule_name}' not found. Available rules: {list(DISPATCHING_RULES.keys())}")
    return DISPATCHING_RULES[rule_name]


def compare_rules(env, rules: Optional[List[str]] = None, num_episodes: int = 10) -> Dict[str, Dict[str, float]]:
    """dispatching.py:442-475 (single env, host coin)."""
    if rules is None:
        rules = list(DISPATCHING_RULES.keys())
    results = {}
    for rule_name in rules:
        rule = get_rule(rule_name)
        total_reward, total_makespan = 0.0, 0.0
        for _ in range(num_episodes):
            reward, makespan = rule.run_episode(env)
            total_reward += reward
            total_makespan += makespan
        results[rule_name] = {"avg_reward": total_reward / num_episodes, "avg_makespan": total_makespan / num_episodes}
    return results


def compare_rules_batched(vec_env, rules: Optional[List[str]] = None) -> Dict[str, Dict[str, float]]:
    """compare_rules over a whole batch: one episode per env and rule, fused rollouts on device."""
    if rules is None:
        rules = list(DISPATCHING_RULES.keys())
    mto = vec_env.instance_scalars[vec_env.env_to_instance, 0].astype(np.float64)
    results = {}
    for rule_name in rules:
        ret, mk = get_rule(rule_name).run_batch(vec_env)
        ret = ret.cpu().numpy().astype(np.float64) / mto
        results[rule_name] = {"avg_reward": float(ret.mean()), "avg_makespan": float(mk.double().mean())}
    return results
