#!/bin/bash
# Round-2 GPU session 3: tests, bench, per-shape probe (default + occupancy variants), mixed-kernel mapping sweep,
# host-helper probe, ncu captures (cfg3 kernel + mixed kernel), sanitizer
mkdir -p gpurun_out
T=${TAG:-r02c}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python tools/probe_shapes.py > gpurun_out/${T}_probe_shapes.json 2> gpurun_out/${T}_probe.err; echo "probe rc=$?"
for v in small4 small5; do
  JSS_B200_LIB=$PWD/jssenv_b200/variants/libjss_b200_$v.so timeout 600 python tools/probe_shapes.py > gpurun_out/${T}_probe_shapes_$v.json 2>> gpurun_out/${T}_probe.err; echo "probe $v rc=$?"
done
python - <<'PY'
import json
for v in ("", "_small4", "_small5"):
    d = json.load(open(f"gpurun_out/r02c_probe_shapes{v}.json"))
    print(v or "default", {k: round(x.get("us_per_launch", x.get("us_per_step", 0)), 1) for k, x in d.items()})
PY
for cfg in "1 0.84 0.0057" "0 0.84 0.0057" "1 0.6 0.008" "1 1.0 0.004"; do
  set -- $cfg
  JSS_MIXED_MAP=$1 JSS_COST_A=$2 JSS_COST_B=$3 timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | sed "s/^/map=$1 /" | tee -a gpurun_out/${T}_probe_mixed.jsonl
done
PROBE_RULE=RANDOM timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_mixed.jsonl
for pb in "1 1" "0 1" "1 0"; do
  set -- $pb
  JSS_HOST_PIN=$1 PROBE_BIND=$2 timeout 300 python tools/probe_host.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_host.jsonl
done
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
TAG=$T bash tools/gpu_sanitize.sh
ls -la gpurun_out | tail -14
