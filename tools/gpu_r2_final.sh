#!/bin/bash
# Round-2 final GPU validation: smoke, parity tests, bench (both arms), shape probe, sanitizer
mkdir -p gpurun_out
T=${TAG:-r02z}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
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
TAG=$T bash tools/gpu_sanitize.sh
