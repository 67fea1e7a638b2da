"""Pack the public Taillard (ta01..ta80) and Demirkol (dmu16..dmu20) benchmark
instances that the reference bundles (JSSEnv/envs/instances/*, parsed at
jss_env.py:72-88) into ONE binary bundle, ``jssenv_b200/data/instances.npz``.

Run in the build container only (needs /root/reference).  The bundle stores, per
instance ``<name>``, an int16 array ``<name>`` of shape (J, M, 2) holding
(machine, duration) — input DATA, not source code.  ``jssenv_b200.instances``
reads it back and can re-emit the standard text format (line 1 ``J M``, then J
lines of M ``machine duration`` pairs) for callers that pass ``instance_path``.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle.ref_shim import REFERENCE_ROOT  # noqa: E402


def parse_taillard_text(path):
    with open(path) as f:
        rows = [list(map(int, ln.split())) for ln in f if ln.strip()]
    J, M = rows[0]
    arr = np.zeros((J, M, 2), dtype=np.int16)
    for j in range(J):
        r = rows[1 + j]
        assert len(r) == 2 * M
        arr[j, :, 0] = r[0::2]
        arr[j, :, 1] = r[1::2]
    return arr


def main():
    src = os.path.join(REFERENCE_ROOT, "JSSEnv", "envs", "instances")
    out = {}
    for name in sorted(os.listdir(src)):
        out[name] = parse_taillard_text(os.path.join(src, name))
    dst = os.path.join(os.path.dirname(__file__), "..", "jssenv_b200", "data", "instances.npz")
    np.savez_compressed(dst, **out)
    print("packed", len(out), "instances ->", os.path.abspath(dst), os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
