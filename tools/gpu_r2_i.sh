#!/bin/bash
# Round-2 GPU session 9 (after the warp-uniformity hints): full validation (smoke, parity tests, both bench arms, probes,
# sanitizer) + ncu --set full of the mixed-batch kernel and of the KJ = 1 / KJ = 2 uniform kernels + the bench launch list
mkdir -p gpurun_out
T=${TAG:-r02i}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?"
timeout 600 python tools/probe_shapes.py > gpurun_out/${T}_probe_shapes.json 2> gpurun_out/${T}_probe.err; echo "probe rc=$?"
python -c "
import json
d = json.load(open('gpurun_out/${T}_probe_shapes.json'))
print({k: round(x.get('us_per_launch', x.get('us_per_step', 0)), 2) for k, x in d.items()})"
for r in FIFO MWR RANDOM; do PROBE_RULE=$r timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_mixed.jsonl; done
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
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step \
    -s 300 -c 2 -f -o gpurun_out/${T}_prof_mixed python /tmp/mixed.py > gpurun_out/${T}_ncu_mixed.log 2>&1; echo "ncu mixed rc=$?"
for inst in ta21 ta51; do
cat > /tmp/small.py <<PY
import sys; sys.path.insert(0, '.')
import torch
from jssenv_b200 import JssVecEnv
env = JssVecEnv(65536, {"instance_path": "$inst"}, auto_reset=True, seed=2)
env.reset(); acts = env.policy("RANDOM").clone()
for k in range(260):
    *_, acts = env.step_sample(acts, "RANDOM")
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step_kernel \
    -s 200 -c 2 -f -o gpurun_out/${T}_prof_$inst python /tmp/small.py > gpurun_out/${T}_ncu_$inst.log 2>&1; echo "ncu $inst rc=$?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jss_ -s 4480 -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 300 --warmup 20 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
grep -c jss_ gpurun_out/${T}_launches.csv
TAG=$T bash tools/gpu_sanitize.sh
