#!/bin/bash
# Round-2 GPU session 1: parity tests, full bench line (all configs + CPU legs), reference arm, ncu launch list,
# ncu --set full capture of the mixed-batch step, per-shape cost probe.
mkdir -p gpurun_out
T=${TAG:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -30 >> gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?"
cat gpurun_out/${T}_bench_reference.json
timeout 600 python tools/probe_shapes.py > gpurun_out/${T}_probe_shapes.json 2> gpurun_out/${T}_probe.err; echo "probe rc=$?"; cat gpurun_out/${T}_probe_shapes.json
# launch list of the bench command (kernel launches of the timed region included: pre-roll = 2 x 2236 launches first)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jss_ -s 4480 -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 300 --warmup 20 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
grep -c jss_ gpurun_out/${T}_launches.csv
# one full capture of the mixed-batch step (cfg5) mid-episode: all three lane classes
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
ls -la gpurun_out | tail -20
