#!/bin/bash
# gpu tests + device-resident bench (no cpu/e2e legs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 2400 --warmup 20 --no-cpu --no-e2e | tee gpurun_out/bench_quick.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"
