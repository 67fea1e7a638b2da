#!/bin/bash
# Round-2 GPU session 6: ncu --set full of the KJ = 1 and KJ = 2 uniform step kernels (where do ~600 instructions per
# 15..30-job env-step go?)
mkdir -p gpurun_out
T=${TAG:-r02f}
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
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step_kernel \
    -s 200 -c 2 -f -o gpurun_out/${T}_prof_$inst python /tmp/small.py > gpurun_out/${T}_ncu_$inst.log 2>&1; echo "ncu $inst rc=$?"
done
ls -la gpurun_out | tail -5
