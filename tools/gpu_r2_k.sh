#!/bin/bash
# Round-2 GPU session 11: how the warp polls the prefetch mbarrier (try_wait / test_wait / try_wait with a suspend hint),
# with the flags hint forced on for the uniform KJ = 4 kernel (hintall*) and in the default configuration
mkdir -p gpurun_out
T=${TAG:-r02k}
for v in ${VARIANTS:-default hintall hintall_test hintall_hint test hint default hintall_test}; do
  if [ $v = default ]; then unset JSS_B200_LIB; else export JSS_B200_LIB=$PWD/jssenv_b200/variants/libjss_b200_$v.so; fi
  PROBE_NAMES=ta21,ta51,ta71 timeout 300 python tools/probe_shapes.py > gpurun_out/${T}_probe_$v.json 2>> gpurun_out/${T}_probe.err
  python -c "
import json
d = json.load(open('gpurun_out/${T}_probe_$v.json')); print('$v', {k: round(x.get('us_per_launch', x.get('us_per_step', 0)), 2) for k, x in d.items()})"
  for r in FIFO; do echo -n "$v $r "; PROBE_RULE=$r timeout 300 python tools/probe_mixed.py 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/${T}_probe_mixed_$v.jsonl; done
done
