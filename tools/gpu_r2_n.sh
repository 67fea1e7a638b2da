#!/bin/bash
# Round-2 GPU session 14: mixed step kernel -- instance id carried with the prefetched tile descriptor, J / M refreshed
# only when the instance is re-staged: A/B against the previous build, then the GPU tests
mkdir -p gpurun_out
T=${TAG:-r02n}
for v in default prev default prev; do
  if [ $v = default ]; then unset JSS_B200_LIB; else export JSS_B200_LIB=$PWD/jssenv_b200/variants/libjss_b200_$v.so; fi
  for r in FIFO RANDOM MWR; do echo -n "$v $r "; PROBE_RULE=$r timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | cut -c1-130; done
done 2>&1 | tee gpurun_out/${T}_probe_mixed_ab.txt
unset JSS_B200_LIB
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
