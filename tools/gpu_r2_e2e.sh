#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02y}
timeout 900 python bench.py --steps 20 --warmup 5 --configs none --no-cpu > gpurun_out/${T}_bench_e2e.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_e2e.json')); e=d['e2e']; print(d['value'], e['value'], e['dma_fraction'], e['dma_fraction_calibration'], e['fp32_dma_variant']['value']); print(e['host_ms_per_step'])"
JSS_HOST_PIN=1 timeout 900 python bench.py --steps 20 --warmup 5 --configs none --no-cpu > gpurun_out/${T}_bench_e2e_pin.json 2>> gpurun_out/${T}_bench.err; echo "bench pin rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_e2e_pin.json')); e=d['e2e']; print(d['value'], e['value'], e['dma_fraction'], e['dma_fraction_calibration'], e['fp32_dma_variant']['value']); print(e['host_ms_per_step'])"
for pb in "0 1" "1 1"; do set -- $pb; JSS_HOST_PIN=$1 PROBE_BIND=$2 timeout 300 python tools/probe_host.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_probe_host.jsonl; done
