#!/bin/bash
# ncu evidence for the step kernel: launch list of a short bench run + one full capture mid-episode.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'jss_env_kernel<4, 1>' -s 700 -c 2 -f -o gpurun_out/prof_step \
    python bench.py --steps 10 --warmup 800 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
