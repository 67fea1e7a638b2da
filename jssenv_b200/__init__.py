"""jssenv_b200 -- B200-native batched job-shop scheduling environment.

Drop-in for the hot path of prosysscience/JSSEnv (``gym.make('jss-v1')``): the
``JssEnv`` facade keeps the reference's single-env object interface, ``JssVecEnv``
is the batched form; both run the hand-written sm_100a kernels in ``csrc/`` through
the C-ABI of ``include/jss_b200.h``.  Importing the package needs neither a GPU nor
the built library; constructing an env does (there is no CPU fallback).
"""
__version__ = "0.1.0"

from .env import JssEnv  # noqa: F401
from .vec_env import JssVecEnv  # noqa: F401
from . import dispatching  # noqa: F401
from .gym_vector import JssGymVectorEnv  # noqa: F401
from .instances import bundled_names, load_instance, parse_taillard, write_taillard  # noqa: F401

try:  # same id / entry-point style as JSSEnv/__init__.py:6-9 (gymnasium is optional here)
    from gymnasium.envs.registration import register as _register

    _register(id="jss-v1", entry_point="jssenv_b200.env:JssEnv")
except Exception:  # pragma: no cover - gymnasium absent in the build image
    pass
