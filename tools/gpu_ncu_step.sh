#!/bin/bash
# one ncu --set full capture of the step kernel (KJ=4) mid-episode
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:${KREGEX:-jss_step_kernel} -s ${SKIP:-700} -c 2 -f -o gpurun_out/${OUT:-prof_step} \
    python bench.py --steps 10 --warmup 800 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
grep -v "^{" gpurun_out/ncu_full.log | tail -5
ls -la gpurun_out/*.ncu-rep
