#!/bin/bash
python -m pytest tests -m gpu -x -q -k "host_pipeline or step_host" 2>&1 | tail -2
python bench.py --steps 600 --warmup 20 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'])"
