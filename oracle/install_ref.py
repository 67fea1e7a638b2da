"""Recipe: install the UNMODIFIED reference package (prosysscience/JSSEnv, pure Python) into oracle/_ref/.

TEST / BENCH INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the history) but it is
NOT gpurun-ignored, so the installed package travels to the GPU box with the snapshot; there ``bench.py`` times the
reference's own NumPy ``step()`` (JSSEnv/envs/jss_env.py:403-481) on the box's host cores next to the GPU numbers
 and the tests can cross-check the C oracle against it.  ``/root/reference`` itself does not
exist on the GPU box.  The install is the one offline pip install the task allows:

    pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>

(from a copy under /tmp because /root/reference is read-only and setuptools writes build/ and *.egg-info into the
source tree; --no-deps because gymnasium / plotly / imageio are not in the wheelhouse -- oracle/ref_shim.py
injects the stand-ins the import needs).  Run by ``__graft_entry__.build()`` whenever /root/reference is present.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("JSS_REFERENCE_ROOT", "/root/reference")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "JSSEnv", "envs", "jss_env.py"))


def install(force: bool = False) -> str:
    """Returns the target directory (oracle/_ref); '' if the reference tree is not present here."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "JSSEnv")):
        return ""
    tmp = tempfile.mkdtemp(prefix="jss_ref_src_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "*.gif"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert installed(), "pip install of the reference did not produce oracle/_ref/JSSEnv"
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv) or "reference tree not present; nothing installed")
