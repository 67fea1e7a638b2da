#!/bin/bash
# rebuild with different compile-time knobs and time the bench (device-resident numbers only)
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
for flags in ${SWEEP:-"-DJSS_MIN_CTAS=4" "-DJSS_MIN_CTAS=3"}; do
  JSS_NVCC_EXTRA="$flags" python -m jssenv_b200.build --force > /dev/null 2>&1
  echo "== $flags" | tee -a gpurun_out/sweep.txt
  python bench.py --steps 1500 --warmup 20 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])" | tee -a gpurun_out/sweep.txt
done
python -m jssenv_b200.build --force > /dev/null 2>&1
