#!/bin/bash
# rebuild with different compile-time knobs (one variant per line of $1 or stdin) and time the bench
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
while IFS= read -r flags; do
  JSS_NVCC_EXTRA="$flags" python -m jssenv_b200.build --force > /dev/null 2>&1
  echo "== [$flags]" | tee -a gpurun_out/sweep.txt
  python bench.py --steps 2000 --warmup 20 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/sweep.txt
done < "${1:-/dev/stdin}"
python -m jssenv_b200.build --force > /dev/null 2>&1
