#!/bin/bash
# Round-2 GPU session 4: tests, bench (hybrid e2e, mixed kernel v3), probes, ncu of the final kernels
mkdir -p gpurun_out
T=${TAG:-r02d}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python tools/probe_shapes.py > gpurun_out/${T}_probe_shapes.json 2> gpurun_out/${T}_probe.err; echo "probe rc=$?"
python -c "
import json
d = json.load(open('gpurun_out/${T}_probe_shapes.json'))
print({k: round(x.get('us_per_launch', x.get('us_per_step', 0)), 1) for k, x in d.items()})"
for r in FIFO RANDOM; do PROBE_RULE=$r timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_mixed.jsonl; done
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step_kernel \
    -s 3000 -c 2 -f -o gpurun_out/${T}_prof_step python bench.py --steps 10 --warmup 800 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_step.log 2>&1; echo "ncu step rc=$?"
cat > /tmp/mixed.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from jssenv_b200 import JssVecEnv
names = ["ta%02d" % (k + 1) for k in range(80)]
n = 65536
env = JssVecEnv(n, {"instance_paths": names, "env_to_instance": np.arange(n) % 80}, auto_reset=True, seed=2)
env.reset(); acts = env.policy("FIFO").clone()
for k in range(400):
    *_, acts = env.step_sample(acts, "FIFO")
torch.cuda.synchronize()
PY
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step \
    -s 300 -c 2 -f -o gpurun_out/${T}_prof_mixed python /tmp/mixed.py > gpurun_out/${T}_ncu_mixed.log 2>&1; echo "ncu mixed rc=$?"
ls -la gpurun_out | tail -8
# A/B: dynamic tail off, opaque warp-base variant (ta80-shaped batch only)
JSS_TAIL=0 PROBE_NAMES=ta71,ta80 timeout 300 python tools/probe_shapes.py > gpurun_out/${T}_probe_notail.json 2>> gpurun_out/${T}_probe.err
JSS_B200_LIB=$PWD/jssenv_b200/variants/libjss_b200_opaque.so PROBE_NAMES=ta71,ta80 timeout 300 python tools/probe_shapes.py > gpurun_out/${T}_probe_opaque.json 2>> gpurun_out/${T}_probe.err
python -c "
import json
for v in ('notail','opaque'):
    d = json.load(open('gpurun_out/${T}_probe_%s.json' % v)); print(v, {k: round(x['us_per_launch'], 1) for k, x in d.items()})"
