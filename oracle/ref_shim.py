"""Loader for the UNMODIFIED reference (prosysscience/JSSEnv) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` and
``tools/pack_instances.py`` to (a) validate the C restatement in
``oracle/jss_oracle.c`` and (b) generate the committed fixtures under
``tests/golden/``.  ``/root/reference`` does not exist on the GPU box; there the loader falls back to
``oracle/_ref`` (the unmodified package installed by ``oracle/install_ref.py``, git-ignored,
shipped with the snapshot), which is what ``bench.py``'s CPU legs time as the reference's own
NumPy ``step()``.  Never imported by the product package.

The reference needs ``gymnasium`` and ``plotly`` only for ``gym.Env`` /
``gym.spaces`` (jss_env.py:8,14,97,112-119) and for ``render`` (jss_env.py:10-11,
683); neither is installed here, so minimal stand-ins are injected into
``sys.modules`` before the import.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_INSTALLED = os.path.join(_HERE, "_ref")          # pip --target install made by oracle/install_ref.py (travels to the GPU box)


def _pick_root():
    env = os.environ.get("JSS_REFERENCE_ROOT")
    for cand in (env, "/root/reference", _INSTALLED):
        if cand and os.path.isfile(os.path.join(cand, "JSSEnv", "envs", "jss_env.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "JSSEnv", "envs", "jss_env.py"))


def _install_stubs():
    if "gymnasium" in sys.modules:
        return
    gym = types.ModuleType("gymnasium")

    class Env:  # gym.Env stand-in: the reference only subclasses it
        pass

    class _Space:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k
            self.shape = k.get("shape")
            self.n = a[0] if a else None

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Discrete = type("Discrete", (_Space,), {})
    spaces.Box = type("Box", (_Space,), {})
    spaces.Dict = type("Dict", (_Space,), {})
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda *a, **k: None
    envs.registration = reg
    gym.Env, gym.spaces, gym.envs = Env, spaces, envs
    sys.modules.update({
        "gymnasium": gym, "gymnasium.spaces": spaces,
        "gymnasium.envs": envs, "gymnasium.envs.registration": reg,
    })
    plotly = types.ModuleType("plotly")
    ff = types.ModuleType("plotly.figure_factory")
    go = types.ModuleType("plotly.graph_objects")
    go.Figure = object
    plotly.figure_factory, plotly.graph_objects = ff, go
    sys.modules.update({"plotly": plotly, "plotly.figure_factory": ff,
                        "plotly.graph_objects": go})


def load_reference():
    """Returns (JssEnv class, dispatching module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from JSSEnv.envs.jss_env import JssEnv  # noqa
    import JSSEnv.dispatching as dispatching  # noqa
    return JssEnv, dispatching


def reference_instance_path(name: str) -> str:
    return os.path.join(REFERENCE_ROOT, "JSSEnv", "envs", "instances", name)
