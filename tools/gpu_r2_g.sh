#!/bin/bash
# Round-2 GPU session 7: warp-uniformity hints (jss_uniform) -- parity of the new default build, A/B against the build
# without hints (= the previous kernels) on every shape / the mixed batch / the fused rollout, ncu of both on ta71
mkdir -p gpurun_out
T=${TAG:-r02g}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
for v in default nohints default nohints; do
  if [ $v = default ]; then unset JSS_B200_LIB; else export JSS_B200_LIB=$PWD/jssenv_b200/variants/libjss_b200_$v.so; fi
  timeout 300 python tools/probe_shapes.py > gpurun_out/${T}_probe_$v.json 2>> gpurun_out/${T}_probe.err
  python -c "
import json
d = json.load(open('gpurun_out/${T}_probe_$v.json')); print('$v', {k: round(x.get('us_per_launch', x.get('us_per_step', 0)), 2) for k, x in d.items()})"
  for r in FIFO MWR RANDOM; do echo -n "$v $r "; PROBE_RULE=$r timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_mixed_$v.jsonl; done
done
cat > /tmp/small.py <<PY
import sys; sys.path.insert(0, '.')
import torch
from jssenv_b200 import JssVecEnv
env = JssVecEnv(65536, {"instance_path": "ta71"}, auto_reset=True, seed=2)
env.reset(); acts = env.policy("RANDOM").clone()
for k in range(860):
    *_, acts = env.step_sample(acts, "RANDOM")
torch.cuda.synchronize()
PY
for v in default nohints; do
  if [ $v = default ]; then unset JSS_B200_LIB; else export JSS_B200_LIB=$PWD/jssenv_b200/variants/libjss_b200_$v.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step_kernel \
    -s 800 -c 2 -f -o gpurun_out/${T}_prof_ta71_$v python /tmp/small.py > gpurun_out/${T}_ncu_$v.log 2>&1; echo "ncu $v rc=$?"
done
