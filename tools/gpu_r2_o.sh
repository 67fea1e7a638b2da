#!/bin/bash
# Round-2 GPU session 15: ncu --set full of the final cfg3 kernel in the stationary episode mix (refreshes
# profiles/r02_step_kernel.txt and profiles/traffic.json for the code as committed)
mkdir -p gpurun_out
T=${TAG:-r02o}
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:jss_step_kernel \
    -s 3000 -c 2 -f -o gpurun_out/${T}_prof_step python bench.py --steps 10 --warmup 800 --no-cpu --no-e2e --configs none \
    > gpurun_out/${T}_ncu_step.log 2>&1; echo "ncu step rc=$?"
tail -3 gpurun_out/${T}_ncu_step.log
