#!/bin/bash
# Round-2 GPU session 5: tests + bench on the current build, kernel A/B (f32x2 fix, smem layout), e2e experiments, 2nd ncu
mkdir -p gpurun_out
T=${TAG:-r02e}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?"
timeout 600 python tools/probe_shapes.py > gpurun_out/${T}_probe_shapes.json 2> gpurun_out/${T}_probe.err; echo "probe rc=$?"
python -c "
import json
d = json.load(open('gpurun_out/${T}_probe_shapes.json'))
print({k: round(x.get('us_per_launch', x.get('us_per_step', 0)), 1) for k, x in d.items()})"
for r in FIFO RANDOM; do PROBE_RULE=$r timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_mixed.jsonl; done
for nt in 1 0; do JSS_HOST_NT=$nt timeout 300 python tools/probe_host.py 2>&1 | tail -1 | sed "s/^/nt=$nt /" | tee -a gpurun_out/${T}_probe_host.jsonl; done
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step_kernel \
    -s 3000 -c 2 -f -o gpurun_out/${T}_prof_step python bench.py --steps 10 --warmup 800 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jss_ -s 4480 -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 300 --warmup 20 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
cat > /tmp/rec.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from jssenv_b200 import JssVecEnv
env = JssVecEnv(4096, {"instance_path": "ta01"}, auto_reset=True, seed=1)
env.reset()
tr = env.rollout_record("RANDOM", 64)
tr = env.rollout_record("RANDOM", 64, out=tr)
env.rollout("RANDOM", 64, write_obs=True)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_env_kernel \
    -s 3 -c 3 -f -o gpurun_out/${T}_prof_rollout python /tmp/rec.py > gpurun_out/${T}_ncu_rollout.log 2>&1; echo "ncu rollout rc=$?"
ls -la gpurun_out | tail -10
