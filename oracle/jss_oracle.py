"""ctypes front-end of oracle/jss_oracle.c -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs; never
from the product package ``jssenv_b200``.  Attribute names follow the reference
(JSSEnv/envs/jss_env.py) so tests read like the reference's own tests.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libjss_oracle.so")
_lib = None

RULES = {"SPT": 0, "FIFO": 1, "MWR": 2, "LWR": 3, "MOR": 4, "LOR": 5, "CR": 6}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jss_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libjss_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)
        L.jsso_create.restype = vp
        L.jsso_create.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p]
        L.jsso_destroy.argtypes = [vp]
        L.jsso_reset.argtypes = [vp]
        L.jsso_step.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)]
        L.jsso_increase_time_step.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
        L.jsso_rule_action.argtypes = [vp, ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_int)]
        L.jsso_set_cr_due_date_factor.argtypes = [vp, ctypes.c_double]
        L.jsso_set_cr_due_date_factor.restype = None
        L.jsso_scalar.restype = ctypes.c_int64
        L.jsso_scalar.argtypes = [vp, ctypes.c_int]
        L.jsso_array.restype = vp
        L.jsso_array.argtypes = [vp, ctypes.c_int]
        L.jsso_masked_random_action.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
        L.jsso_run_random.restype = ctypes.c_int64
        L.jsso_run_random.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        L.jsso_run_random_timed.restype = ctypes.c_int64
        L.jsso_run_random_timed.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double,
                                            ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        _lib = L
    return _lib


class OracleError(IndexError):
    """The reference raises IndexError in these situations (jss_env.py:444, 517)."""


class OracleEnv:
    """Single env with the reference's attribute surface, backed by the C oracle."""

    def __init__(self, machine: np.ndarray, duration: np.ndarray):
        L = lib()
        machine = np.ascontiguousarray(machine, dtype=np.int32)
        duration = np.ascontiguousarray(duration, dtype=np.int32)
        assert machine.shape == duration.shape and machine.ndim == 2
        self.jobs, self.machines = machine.shape
        self._L = L
        self._h = L.jsso_create(self.jobs, self.machines,
                                machine.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                duration.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        if not self._h:
            raise ValueError("invalid instance (needs jobs > 0 and machines > 1)")
        self.max_time_op = L.jsso_scalar(self._h, 0)
        self.max_time_jobs = L.jsso_scalar(self._h, 1)
        self.sum_op = L.jsso_scalar(self._h, 2)
        J, M = self.jobs, self.machines
        self._views = {}
        for name, idx, dt, shape in [
            ("legal_actions", 0, np.bool_, (J + 1,)), ("state", 1, np.float64, (J, 7)),
            ("time_until_available_machine", 2, np.int64, (M,)),
            ("time_until_finish_current_op_jobs", 3, np.int64, (J,)),
            ("todo_time_step_job", 4, np.int64, (J,)),
            ("total_perform_op_time_jobs", 5, np.int64, (J,)),
            ("needed_machine_jobs", 6, np.int64, (J,)),
            ("total_idle_time_jobs", 7, np.int64, (J,)),
            ("idle_time_jobs_last_op", 8, np.int64, (J,)),
            ("illegal_actions", 9, np.bool_, (M, J)),
            ("action_illegal_no_op", 10, np.bool_, (J,)),
            ("machine_legal", 11, np.bool_, (M,)),
            ("solution", 12, np.int64, (J, M)),
            ("jobs_length", 14, np.int64, (J,)),
        ]:
            ptr = L.jsso_array(self._h, idx)
            n = int(np.prod(shape))
            buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
            setattr(self, name, np.frombuffer(buf, dtype=dt).reshape(shape))
        self.instance_matrix = np.stack([machine, duration], axis=-1).astype(np.int64)

    def __del__(self):
        try:
            if self._h:
                self._L.jsso_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- scalars ----------------------------------------------------------
    @property
    def current_time_step(self):
        return int(self._L.jsso_scalar(self._h, 5))

    @property
    def last_time_step(self):
        v = int(self._L.jsso_scalar(self._h, 6))
        return float("inf") if v < 0 else v

    @property
    def nb_legal_actions(self):
        return int(self._L.jsso_scalar(self._h, 3))

    @property
    def nb_machine_legal(self):
        return int(self._L.jsso_scalar(self._h, 4))

    @property
    def next_time_step(self):
        n = int(self._L.jsso_scalar(self._h, 7))
        ptr = self._L.jsso_array(self._h, 13)
        return list(np.ctypeslib.as_array((ctypes.c_int64 * max(n, 1)).from_address(ptr))[:n])

    # -- reference API ------------------------------------------------------
    def _obs(self):
        return {"real_obs": self.state, "action_mask": self.legal_actions}

    def get_legal_actions(self):
        return self.legal_actions

    def reset(self):
        self._L.jsso_reset(self._h)
        return self._obs()

    def step(self, action: int):
        r, raw, done = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int()
        rc = self._L.jsso_step(self._h, int(action), ctypes.byref(r), ctypes.byref(raw), ctypes.byref(done))
        if rc != 0:
            raise OracleError("oracle step failed rc=%d" % rc)
        self.last_raw_reward = int(raw.value)
        return self._obs(), float(r.value), bool(done.value), False, {}

    def increase_time_step(self):
        hole = ctypes.c_int64()
        rc = self._L.jsso_increase_time_step(self._h, ctypes.byref(hole))
        if rc != 0:
            raise OracleError("pop from empty list")
        return int(hole.value)

    def set_cr_due_date_factor(self, factor: float):
        """CriticalRatio(due_date_factor=...) of the reference (dispatching.py:337-349); default 1.5."""
        self._L.jsso_set_cr_due_date_factor(self._h, float(factor))

    def rule_action(self, rule: str, u: float):
        """Returns (action, consumed): `consumed` says whether the reference would
        have drawn np.random.random() (only when the no-op is legal)."""
        c = ctypes.c_int()
        a = self._L.jsso_rule_action(self._h, RULES[rule], float(u), ctypes.byref(c))
        return int(a), bool(c.value)

    def masked_random_action(self, seed: int, env: int, ctr: int) -> int:
        return int(self._L.jsso_masked_random_action(self._h, seed, env, ctr))

    def run_random(self, seed: int, env: int, n_steps: int):
        ep, ms = ctypes.c_int64(0), ctypes.c_int64(0)
        n = self._L.jsso_run_random(self._h, seed, env, n_steps, ctypes.byref(ep), ctypes.byref(ms))
        return int(n), int(ep.value), int(ms.value)

    def run_random_timed(self, seed: int, env: int, seconds: float):
        ep, ms = ctypes.c_int64(0), ctypes.c_int64(0)
        n = self._L.jsso_run_random_timed(self._h, seed, env, float(seconds), ctypes.byref(ep), ctypes.byref(ms))
        return int(n), int(ep.value), int(ms.value)
