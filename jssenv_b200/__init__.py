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
from .gym_vector import JssGymVectorEnv, create_env  # noqa: F401
from .instances import bundled_names, load_instance, parse_taillard, write_taillard  # noqa: F401



def register_gymnasium() -> bool:
    """Register the id the reference registers (JSSEnv/__init__.py:6-9: ``jss-v1``), pointing at the drop-in
    facade, so ``gym.make('jss-v1', env_config={...})`` (README.md:46) keeps working.  Returns False when
    gymnasium is not importable (it is optional: neither the build image nor the GPU box has it)."""
    try:
        from gymnasium.envs.registration import register as _register
    except Exception:
        return False
    _register(id="jss-v1", entry_point="jssenv_b200.env:JssEnv")
    return True


register_gymnasium()
