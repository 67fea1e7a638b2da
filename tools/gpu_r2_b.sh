#!/bin/bash
# Round-2 GPU session 2: parity tests, bench (packed e2e, static mixed kernel), cost-model sweep for cfg5, ncu of the mixed kernel
mkdir -p gpurun_out
T=${TAG:-r02b}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
for ab in "0.84 0.0057" "0.5 0.009" "0.3 0.011" "1.0 0.004" "0.0 0.0136"; do
  set -- $ab
  JSS_COST_A=$1 JSS_COST_B=$2 timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_mixed.jsonl
done
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
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jss_ -s 4480 -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 300 --warmup 20 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
grep -c jss_ gpurun_out/${T}_launches.csv
ls -la gpurun_out | tail -12
