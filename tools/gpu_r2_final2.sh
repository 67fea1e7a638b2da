#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02w}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?"
